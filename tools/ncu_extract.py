"""Turn `ncu -i X.ncu-rep --page raw --csv` / `--page source --csv` exports into the small summaries kept under
profiles/: a transposed table of the metrics the design notes quote, and the executed-SASS mix of a kernel
(per unit of work), with FP64 instructions classified by how many distinct register operands they read
(on B200 a DFMA with three distinct register operands occupies the FP64 pipe for 3 cycles instead of 2,
tools/fp64_operands_ubench.cu).

usage: ncu_extract.py raw  <raw.csv> <out.csv>
       ncu_extract.py mix  <src.csv> <out.json> <units of work> <label>
       ncu_extract.py traffic <raw.csv> <key> <units per launch> <source label> [launch index]
           -> profiles/traffic.json[key] = DRAM bytes per unit of that launch, with the kernel's register count and the
              digest of the csrc/ sources it was built from (bench.py quotes the figure only while both still match)"""
import collections
import csv
import json
import re
import sys

KEEP = re.compile(r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
                  r"gpu__dram_throughput.*|launch__(block_size|grid_size|registers_per_thread|occupancy_limit_.*|waves_per_multiprocessor)|"
                  r"lts__t_sector_hit_rate\.pct|lts__t_sectors_srcunit_tex_op_read\.sum|l1tex__throughput\.avg\.pct_of_peak_sustained_active|"
                  r"l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                  r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__cycles_elapsed\.avg|"
                  r"smsp__sass_thread_inst_executed_op_d(add|fma|mul)_pred_on\.sum(\.per_cycle_elapsed)?|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|sm__throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def raw(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        for j, name in enumerate(hdr):
            if KEEP.match(name):
                w.writerow([name, units[j]] + [r[j] for r in launches])


def mix(path, out, units, label):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Source" in r and "Instructions Executed" in r:
            hdr, start = r, i + 1
            break
    si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
    per_op = collections.Counter()
    fp64 = collections.Counter()
    total = 0
    for r in rows[start:]:
        try:
            n = int(r[ei])
        except (ValueError, IndexError):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?(\w+)(\.[\w.]+)?\s*(.*?);?$", r[si].strip())
        if not m or n == 0:
            continue
        op = m.group(2)
        total += n
        per_op[op] += n
        if op in ("DFMA", "DMUL", "DADD"):
            srcs = [o.strip() for o in m.group(4).split(",")][1:]
            regs = {re.sub(r"[-|]|\.reuse", "", o) for o in srcs if re.match(r"-?\|?R\d+", o)}
            fp64[f"{op} with {len(regs)} distinct register operand(s)"] += n
    u = float(units)
    json.dump({"kernel": label, "warp_instructions_per_unit": total / u,
               "per_opcode": {k: v / u for k, v in per_op.most_common(30)},
               "fp64_by_register_operands": {k: v / u for k, v in sorted(fp64.items())},
               "note": "unit = one (grid point, in-cutoff atom) pair per lane, i.e. warp instructions / (pairs / 32)"
               if "build" in label else "unit as stated in the label"}, open(out, "w"), indent=1)


# the sources a kernel is compiled from (bench.py keeps the same table)
KERNEL_SOURCES = {"build": ["common.cuh", "fastmath.cuh", "grid_build.cuh", "math_coeffs.cuh"], "interp": ["common.cuh", "interp.cuh"]}


def traffic(path, key, units, source, launch=0):
    import hashlib
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rows = list(csv.reader(open(path)))
    hdr, unit_row, r = rows[0], rows[1], rows[2 + int(launch)]

    def val(name):
        j = hdr.index(name)
        v = float(r[j])
        u = unit_row[j].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)
    h = hashlib.sha1()
    d = os.path.join(root, "crystalenergygrids.jl_b200", "csrc")
    for f in KERNEL_SOURCES["interp" if key.startswith("interp") else key]:
        h.update(open(os.path.join(d, f), "rb").read())
    out = os.path.join(root, "profiles", "traffic.json")
    data = json.load(open(out)) if os.path.isfile(out) else {}
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    data[key] = {"kernel": r[hdr.index("Kernel Name")], "registers": int(float(r[hdr.index("launch__registers_per_thread")])),
                 "csrc_digest": h.hexdigest()[:12], "dram_bytes_read": rd, "dram_bytes_write": wr, "units_per_launch": float(units),
                 "dram_bytes_per_unit": (rd + wr) / float(units), "gpu_time_ms": val("gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None,
                 "source": source}
    json.dump(data, open(out, "w"), indent=1)
    print(key, data[key])


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        mix(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
