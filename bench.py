#!/usr/bin/env python
"""Benchmark of the energy-grid hot path (BASELINE.json metric: grid points/sec (LJ+Ewald FP64) at
1/2/4/8 B200; tricubic interpolations/sec).

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One JSON line on stdout (rank 0).  Under torchrun (N > 1) one rank per GPU; the grids are sharded into
contiguous x-slabs (no data-path collective) and interpolation queries by index.  Timings are
device-side (CUDA events on the launching stream), max over ranks; end-to-end figures use the host clock around
the reference-facing call.

Top-level keys: C4 (BASELINE configs[3]) grid construction.  Sibling blocks: `e2e_device` (grid pair built,
gathered over NVLink and interpolable on every GPU), `c1`, `interp` (C2), `ewald`, `c3`, `c5` (configs[2] and
configs[4]), `latency_us`, `guest`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_BUILD = ("C4: CHA_1.4_3b4eeb96 (972 framework atoms, rhombohedral 28.377 Å cell), Na+ probe: fused "
                  "VdW (Buckingham+HardSphere) + Ewald real-space grids, value + 7 derivative channels, 0.10 Å spacing "
                  "(326x306x284 = 28,330,704 points), 12 Å cutoff, x-slab sharded")
WORKLOAD_INTERP = ("C2: tricubic interpolation of 10^7 random positions (splitmix64 seed 1) on the C1 Na+ VdW and "
                   "Coulomb grids (0.15 Å, 218x204x190 nodes, 270 MB each)")
WORKLOAD_C3 = ("C3: synthetic MFI-shaped framework (288 atoms, 20.022x19.899x13.383 Å, supercell 2x2x2 = 2304 atoms), CH4, C_co2 "
               "and O_co2 Lennard-Jones probes + Ewald real-space grid fused in one pass, 0.15 Å (134x134x90 nodes x 4 grids)")
WORKLOAD_C5 = ("C5: synthetic MOF-shaped framework (5000 atoms, 50 Å cubic cell), CH4 Lennard-Jones + Ewald real-space grids "
               "at 0.10 Å (502^3 = 126.5 M points x 2 grids, 4.05 GB each), x-slab sharded, then 10^9 device-generated "
               "interpolation queries on both grids, sharded by query")
N_QUERIES = 10_000_000
N_RECIP = 1_000_000
N_SWEEP = 1_000_000_000


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return {"cpu_model": model, "cpus_online": os.cpu_count(), "cpus_in_affinity_mask": aff}


def bind_to_gpu_cpus(gpu_index: int):
    """Run this process on the CPU cores NVML reports as local to its GPU, so that the host arrays it first
    touches (and pins) sit on the NUMA node behind the GPU's PCIe root.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def slabs(n, parts):
    from ceg_b200.sharding import split_range
    return split_range(n, parts)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        out["samples"] = len(rows)
        if sm:
            top = [x for x in sm if x >= 0.5 * max(sm)]   # samples under load
            out["sm_mhz"] = float(np.median(top))
            out["sm_max_mhz"] = float(rows[0][2])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nme in enumerate(names):
            if any(r[5 + i].strip().lower().startswith("active") for r in rows):
                out["reasons"].append(nme)
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


KERNEL_SOURCES = {"build": ["common.cuh", "fastmath.cuh", "grid_build.cuh", "math_coeffs.cuh"], "interp": ["common.cuh", "interp.cuh"]}


def csrc_digest(key: str) -> str:
    """Digest of the sources a kernel is compiled from (tools/ncu_extract.py stores the same digest next to every
    DRAM-traffic figure it extracts from an ncu capture)."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "crystalenergygrids.jl_b200", "csrc")
    for f in KERNEL_SOURCES["interp" if key.startswith("interp") else key]:
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:12]


def profiled_traffic(key: str, regs_loaded: int):
    """DRAM bytes per unit of the dominant kernel from the committed ncu capture (profiles/traffic.json, written by
    tools/ncu_extract.py traffic): quoted only if the capture was taken on the kernel that is loaded now (same source
    digest, same register count) — a stale capture yields null instead of a wrong number."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None, "no ncu capture committed"
    d = json.load(open(p)).get(key)
    if not d:
        return None, "no ncu capture for this kernel"
    if int(d.get("registers", -1)) != int(regs_loaded):
        return None, f"ncu capture is of a {d.get('registers')}-register build, the loaded kernel uses {regs_loaded}: stale"
    if d.get("csrc_digest") != csrc_digest(key):
        return None, f"kernel sources changed since the ncu capture ({d.get('csrc_digest')} -> {csrc_digest(key)}): stale"
    return float(d["dram_bytes_per_unit"]), d.get("source", "profiles/")


# ---------------------------------------------------------------------------------------------------
def flops_per_point(inp, stats):
    """per in-cutoff pair 8 (distance) + 64 (Ewald real-space) for every atom + 59
    (Buckingham) / 48 (LJ) for the atoms the probe has a VdW term with."""
    n_in = stats["pairs_in_cutoff"] / max(1, stats["n_points"])
    types, counts = np.unique(inp.types, return_counts=True)
    frac = {int(t): c / len(inp.types) for t, c in zip(types, counts)}
    per_pair = 8.0 + (64.0 if inp.want_coulomb else 0.0)
    for p in range(inp.table.shape[0]):
        for t, f in frac.items():
            fl = int(inp.table[p, t]["flags"])
            if fl & 4:
                per_pair += 59.0 * f
            elif fl & 2:
                per_pair += 48.0 * f
    return n_in * per_pair, n_in


class Ctx:
    pass


def run_product(args):
    import torch
    import torch.distributed as dist
    import ceg_b200 as ceg
    from ceg_b200 import workloads as wl

    rank, local_rank, world = env_rank()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    local_cpus = bind_to_gpu_cpus(local_rank) if world > 1 else 0
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def min_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    K, W = args.steps, args.warmup
    stream = torch.cuda.current_stream().cuda_stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    launches0 = ceg._lib.kernel_launch_count()
    fp64_peak = ceg._lib.measure_fp64_peak(dev)
    fp64_peak_regops = ceg._lib.measure_fp64_peak_regops(dev)
    hbm_peak, hbm_src = measured_peaks()

    def timed_build(ps, gs, a, b, ptrs, reps, warm):
        for _ in range(warm):
            ceg._lib.grid_build_dev(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, ptrs, stream)
        barrier()
        l0 = ceg._lib.kernel_launch_count()
        ev0.record()
        for _ in range(reps):
            ceg._lib.grid_build_dev(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, ptrs, stream)
        ev1.record()
        barrier()
        mine = ev0.elapsed_time(ev1) / reps
        return mine, max_over_ranks(mine), ceg._lib.kernel_launch_count() - l0

    class Handoff:
        """Sharded build -> every rank holds interpolable handles of every grid, all on the devices: each rank builds its
        x-slab in `pieces` sub-slabs (ceg_grid_build_dev); the all-gather of a finished sub-slab (NCCL over NVLink) runs
        while the next one is being built; ceg_grid_upload_parts_dev then writes the interpolation records straight from
        the gathered parts (parse_grid's conversion on the device, the full `.grid`-order array is never assembled).
        The reference makes this hand-off through `.grid` files (raspa.jl:430-438)."""

        def __init__(self, ps, inp_, cset_, precs, pieces=2):
            self.ps, self.inp, self.cset, self.precs = ps, inp_, cset_, precs
            self.gs = ceg._lib.make_gridspec(cset_)
            nx_, self.ny, self.nz = cset_.nodes
            sl = slabs(nx_, world)
            tx = 4
            self.chunks = []        # per rank: list of (i0, i1)
            for (s0, s1) in sl:
                cuts = [min(s1, s0 + -(-((s1 - s0) * k) // (pieces * tx)) * tx) for k in range(pieces)] + [s1]   # whole tiles; the last is the smallest
                self.chunks.append([(cuts[k], cuts[k + 1]) for k in range(pieces) if cuts[k + 1] > cuts[k]])
            self.pieces = min(len(c) for c in self.chunks)
            assert all(len(c) == self.pieces for c in self.chunks)
            self.pmax = [max(c[k][1] - c[k][0] for c in self.chunks) for k in range(self.pieces)]
            ng = len(precs)
            mine = self.chunks[rank]
            self.exact = [[torch.empty((8, mine[k][1] - mine[k][0], self.ny, self.nz), dtype=torch.float32, device="cuda")
                           for _ in range(ng)] for k in range(self.pieces)]
            self.recv = None
            if world > 1:
                self.recv = [[[torch.empty((8, self.pmax[k], self.ny, self.nz), dtype=torch.float32, device="cuda")
                               for _ in range(world)] for _ in range(ng)] for k in range(self.pieces)]

        def step(self):
            works = []
            for k in range(self.pieces):
                i0, i1 = self.chunks[rank][k]
                ceg._lib.grid_build_dev(self.ps.handle, self.gs, i0, i1, lam_v, thr_v, lam_c, thr_c,
                                        [t.data_ptr() for t in self.exact[k]], stream)
                if world > 1:
                    for g, t in enumerate(self.exact[k]):
                        send = t
                        if i1 - i0 != self.pmax[k]:
                            send = self.recv[k][g][rank]
                            send[:, :i1 - i0].copy_(t)
                        works.append(dist.all_gather(self.recv[k][g], send, async_op=True))   # overlaps the next sub-slab's build
            for w in works:
                w.wait()
            out = []
            for g, prec in enumerate(self.precs):
                if world > 1:
                    ptrs_, sl_, st_ = [], [], []
                    for r in range(world):
                        for k in range(self.pieces):
                            ptrs_.append(self.recv[k][g][r].data_ptr())
                            sl_.append(self.chunks[r][k])
                            st_.append(self.pmax[k])
                else:
                    ptrs_ = [self.exact[k][g].data_ptr() for k in range(self.pieces)]
                    sl_ = list(self.chunks[0])
                    st_ = [c[1] - c[0] for c in sl_]
                out.append(ceg.EnergyGrid.from_device_parts(self.cset, self.inp.supercell, prec, ptrs_, sl_, st_, device=dev))
            return out

    # ================= C4 construction: kernel-only, outputs resident in HBM ===========================
    inp, cset1, ex = wl.load_fixture("cha_na")
    cset4 = wl.respace(cset1, 0.10)
    nx, ny, nz = cset4.nodes
    sl4 = slabs(nx, world)
    a, b = sl4[rank]
    gs4 = ceg._lib.make_gridspec(cset4)
    ps = ceg.ProbeSystem(None, None, inputs=inp, device=dev)
    d_out = [torch.empty((8, b - a, ny, nz), dtype=torch.float32, device="cuda") for _ in range(2)]
    ptrs = [t.data_ptr() for t in d_out]
    sampler = ClockSampler(dev)
    sampler.start()
    my_ms, build_ms, build_launches = timed_build(ps, gs4, a, b, ptrs, K, W)
    clocks = sampler.stop()
    stats = ceg._lib.probe_last_stats(ps.handle)
    fl_pp, n_in_pp = flops_per_point(inp, stats)
    my_tflops = fl_pp * stats["n_points"] / (my_ms * 1e-3) / 1e12
    total_points = nx * ny * nz
    build_value = total_points / (build_ms * 1e-3)
    build_regs = ceg._lib.kernel_registers(0, dev)
    tr_b, tr_b_src = profiled_traffic("build", build_regs)

    # ================= C4 end to end, (1): through the host C ABI into pinned host arrays ==============
    host = [torch.empty((8, nx, ny, nz), dtype=torch.float32).pin_memory() if world == 1 else
            torch.empty((8, nx, ny, nz), dtype=torch.float32) for _ in range(2)]
    host_np = [h.numpy() for h in host]
    if world > 1:   # pin this rank's slab only: 8 contiguous runs per grid (sharding.slab_file_runs)
        for h in host_np:
            for c in range(8):
                ceg._lib.host_register(h[c, a:b])
    h2d_bytes = (inp.positions.nbytes + inp.types.nbytes + inp.charges.nbytes + inp.table.nbytes + 200)
    d2h_bytes = 2 * 8 * (b - a) * ny * nz * 4
    phases = {"probe_create_ms": [], "image_list_host_ms": [], "kernel_span_ms": [], "copy_tail_ms": [], "build_call_ms": [],
              "probe_free_ms": []}

    def e2e_step(record=False):
        t0 = time.perf_counter()
        p2 = ceg.ProbeSystem(None, None, inputs=inp, device=dev)
        t1 = time.perf_counter()
        ceg._lib.grid_build(p2.handle, gs4, a, b, lam_v, thr_v, lam_c, thr_c, host_np, None)
        t2 = time.perf_counter()
        tm = ceg._lib.probe_last_timing(p2.handle)
        p2.close()
        t3 = time.perf_counter()
        if record:
            phases["probe_create_ms"].append((t1 - t0) * 1e3)
            phases["build_call_ms"].append((t2 - t1) * 1e3)
            phases["probe_free_ms"].append((t3 - t2) * 1e3)
            for k in ("image_list_host_ms", "kernel_span_ms", "copy_tail_ms"):
                phases[k].append(tm[k])

    n_e2e = max(2, K // 2)
    for _ in range(max(2, W // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step(True)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_e2e
    e2e_value = total_points / (e2e_ms * 1e-3)
    breakdown = {k: float(np.mean(v)) for k, v in phases.items()}
    breakdown["note"] = ("rank 0, mean per call: ceg_probe_create (uploads, streams from the library's pool) / host time "
                         "building image lists (cached on the handle: a fresh handle per call builds it once) / device span "
                         "first-to-last kernel / tail between the last kernel and the last slab copy / whole "
                         "ceg_grid_build call / ceg_probe_free")
    if world > 1:   # unpin before the arrays are released: a stale registration collides with later mappings (NCCL buffers)
        for h in host_np:
            for c in range(8):
                ceg._lib.host_unregister(h[c, a:b])
    del host, host_np

    # ================= C4 end to end, (2): grid pair resident and interpolable on EVERY GPU =============
    # The hand-off the reference makes through `.grid` files (raspa.jl:430-438), without host memory in the path.
    # Product path (ceg.ShardedGridBuild): the build kernel of every rank writes each node as a finished interpolation
    # record into the grids of ALL ranks (NVLink peer stores through CUDA IPC mappings; ceg_grid_build_records) — the
    # all-gather and the re-layout pass are fused into the kernel's epilogue.  The destination grids are allocated and
    # mapped once, like the pinned host arrays of the host-landing arm above.
    def exchange(o):
        got = [None] * world
        dist.all_gather_object(got, o)
        return got

    def sharded_build(inputs_, cset_):
        """ShardedGridBuild on every rank, or (None, why) on every rank if any of them could not set it up (e.g. CUDA IPC
        refused by the container): the bench then still reports the NCCL hand-off, with the error in the JSON line."""
        sb, err = None, None
        try:
            if os.environ.get("CEG_BENCH_NO_FUSED"):       # exercises the reporting of a refused set-up
                raise RuntimeError("disabled by CEG_BENCH_NO_FUSED")
            sb = ceg.ShardedGridBuild(inputs_, cset_, inputs_.supercell, (np.inf, 1e-6), rank, world, exchange, device=dev)
        except Exception as e:  # noqa: BLE001
            err = repr(e)
        if min_over_ranks(0.0 if sb is None else 1.0) < 1.0:
            if sb is not None:
                sb.close()
                for eg in sb.grids:
                    eg.close()
            return None, err or "ShardedGridBuild failed on another rank"
        return sb, None

    sb4, fused_error = sharded_build(inp, cset4)
    chk_pos = wl.random_positions(cset4.cell.mat, 1000, 5)

    def checksum(eg):
        v = ceg.interpolate_grid(eg, chk_pos)
        return float(np.where(v == 1e100, 0.0, v).sum())

    dr_ms, chk, chk_same = None, None, True
    if sb4 is not None:
        def device_resident_step():
            sb4.build(stream)
            torch.cuda.synchronize()
            barrier()                    # every rank's slab has landed in every replica

        for _ in range(2):
            device_resident_step()
        n_dr = max(3, K // 2)
        t0 = time.perf_counter()
        for i in range(n_dr):
            device_resident_step()
        dr_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_dr
        # proof of use: the same 1000 queries on every rank's replica
        chk = checksum(sb4.grids[0]) + checksum(sb4.grids[1])
        chk_same = (min_over_ranks(chk) == max_over_ranks(chk))

    # The library-collective baseline of the same hand-off (round-2 first version, kept as the comparison): sub-slab
    # ceg_grid_build_dev -> NCCL all-gather overlapped with the next sub-slab's build -> ceg_grid_upload_parts_dev.
    n_pieces = int(os.environ.get("CEG_BENCH_PIECES", 0)) or (1 if world == 1 else 2 if world < 4 else 4)
    handoff4 = Handoff(ps, inp, cset4, (np.inf, 1e-6), pieces=n_pieces)
    for eg in handoff4.step():
        eg.close()
    barrier()
    n_nc = 3
    t0 = time.perf_counter()
    for i in range(n_nc):
        egs = handoff4.step()
        torch.cuda.synchronize()
        if i < n_nc - 1:
            for eg in egs:
                eg.close()
    barrier()
    nccl_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_nc
    chk_nccl = checksum(egs[0]) + checksum(egs[1])
    chk_same = chk_same and (min_over_ranks(chk_nccl) == max_over_ranks(chk_nccl))
    if chk is not None:
        chk_same = chk_same and (chk_nccl == chk)      # both hand-offs produce the same grids
    for eg in egs:
        eg.close()
    del d_out, handoff4
    if sb4 is not None:
        sb4.close(barrier=barrier)
        for eg in sb4.grids:
            eg.close()
    else:
        dr_ms = nccl_ms                                # reported with "fused_error" below
    del sb4

    # ================= C1 (0.15 Å) on this GPU: the "< 1 s" headline, and the grids for C2 ============
    nx1, ny1, nz1 = cset1.nodes
    gs1 = ceg._lib.make_gridspec(cset1)
    c1_pinned = [torch.empty((8, nx1, ny1, nz1), dtype=torch.float32).pin_memory() for _ in range(2)]
    c1_host = [t.numpy() for t in c1_pinned]
    ceg._lib.grid_build(ps.handle, gs1, 0, nx1, lam_v, thr_v, lam_c, thr_c, c1_host, None)   # warm-up (buffers)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ceg._lib.grid_build(ps.handle, gs1, 0, nx1, lam_v, thr_v, lam_c, thr_c, c1_host, None)
    c1_e2e_ms = (time.perf_counter() - t0) * 1e3
    d1 = [torch.empty((8, nx1, ny1, nz1), dtype=torch.float32, device="cuda") for _ in range(2)]
    p1 = [t.data_ptr() for t in d1]
    c1_kernel_ms, _, _ = timed_build(ps, gs1, 0, nx1, p1, 6, 3)
    del d1
    ps.close()

    # ================= interpolation (C2) ============================================================
    gK = [(g.astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32) for g in c1_host]   # parse_grid scaling
    del c1_host, c1_pinned
    eg_v = ceg.EnergyGrid(cset1, (1, 1, 1), float("inf"), True, gK[0], dev)
    eg_c = ceg.EnergyGrid(cset1, (1, 1, 1), 1e-6, True, gK[1], dev)
    hv, hc = eg_v.handle, eg_c.handle
    hp = eg_v.pair_with(eg_c)
    qa, qb = slabs(N_QUERIES, world)[rank]
    nq = qb - qa
    pos_host = torch.from_numpy(wl.random_positions(cset1.cell.mat, nq, 1, start=qa)).pin_memory()
    d_pos = pos_host.cuda()
    d_e = [torch.empty(nq, dtype=torch.float64, device="cuda") for _ in range(4)]
    KI = max(K, 20)

    def timed(fn, reps):
        for _ in range(W):
            fn()
        barrier()
        l0 = ceg._lib.kernel_launch_count()
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        barrier()
        mine = ev0.elapsed_time(ev1) / reps
        return mine, max_over_ranks(mine), ceg._lib.kernel_launch_count() - l0

    def interp_separate():
        ceg._lib.interpolate_dev(hv, nq, d_pos.data_ptr(), d_e[0].data_ptr(), stream)
        ceg._lib.interpolate_dev(hc, nq, d_pos.data_ptr(), d_e[1].data_ptr(), stream)

    def interp_pair():
        ceg._lib.interpolate_pair_dev(hp, nq, d_pos.data_ptr(), d_e[2].data_ptr(), d_e[3].data_ptr(), stream)

    sep_mine, sep_ms, _ = timed(interp_separate, KI)
    sampler2 = ClockSampler(dev)
    sampler2.start()
    pair_mine, pair_ms, interp_launches = timed(interp_pair, KI)
    clocks2 = sampler2.stop()
    assert torch.equal(d_e[0], d_e[2]) and torch.equal(d_e[1], d_e[3]), "fused pair records differ from the separate lookups"
    # headline = the faster of the two device paths (the fused records double the working set of a launch: on these
    # 270 MB grids it crosses the ~500 MB footprint beyond which the B200's random-sector rate collapses)
    use_pair = pair_ms < sep_ms
    best_ms = pair_ms if use_pair else sep_ms
    interp_value = 2 * N_QUERIES / (best_ms * 1e-3)
    pair_gbs = nq * 552.0 / (pair_mine * 1e-3) / 1e9
    sep_gbs = 2 * nq * 288.0 / (sep_mine * 1e-3) / 1e9
    blocked_fraction = float((d_e[2] == 1e100).double().mean().item())
    checksum = float(torch.where(d_e[2] == 1e100, torch.zeros_like(d_e[2]), d_e[2]).sum().item())
    pair_regs = ceg._lib.kernel_registers(2, dev)
    sep_regs = ceg._lib.kernel_registers(1, dev)
    tr_i, tr_i_src = profiled_traffic("interp_pair" if use_pair else "interp", pair_regs if use_pair else sep_regs)
    # e2e: pinned host positions in, host energies out, through the reference-facing batched call
    out_host = torch.empty((2, nq), dtype=torch.float64).pin_memory()
    pos_np = pos_host.numpy()
    out_np = out_host.numpy()

    def interp_e2e():     # both grids at the same positions (grids.jl:316-318), host arrays in and out
        if use_pair:
            ceg._lib.interpolate_pair(hp, pos_np, out_np)
        else:
            ceg._lib.interpolate_multi([hv, hc], pos_np, out_np)
    interp_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        interp_e2e()
    barrier()
    interp_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 3
    assert np.array_equal(out_np[0], d_e[2].cpu().numpy())

    # ================= reciprocal-space Ewald + fused energy_points (A9/A12), this rank's share =========
    n_unit = len(ex["site_charge"])
    fw = ceg.RASPASystem(ex["unit_mat"], ex["positions"][:n_unit], ["X"] * n_unit, np.ones(n_unit), ex["charges"][:n_unit], False)
    ef = ceg.initialize_ewald(fw, device=dev)
    qna = float(ex["probe_charge"])
    n_rec = N_RECIP // world
    rec_pos = torch.from_numpy(wl.random_positions(cset1.cell.mat, n_rec, 7, start=rank * n_rec)).pin_memory()
    d_rec = rec_pos.cuda()
    d_rec_out = torch.empty(n_rec, dtype=torch.float64, device="cuda")
    qarr = np.array([qna])
    _, rec_ms, _ = timed(lambda: ceg._lib.ewald_reciprocal_dev(ef.handle, n_rec, qarr, d_rec.data_ptr(), 0.0, 0.0,
                                                               d_rec_out.data_ptr(), stream), KI)
    n_k = ef.kspace.num_kvecs
    na_sys = ceg.RASPASystem(np.diag([np.inf] * 3), np.zeros((1, 3)), ["Na"], np.ones(1), np.array([qna]), True)
    setup = ceg.CrystalEnergySetup(fw, na_sys, eg_c, np.array([qna]), [eg_v], [0], ef, None, ceg.BlockFile(cset1))
    if not use_pair:
        eg_v.unpair()      # the library evaluates through the fused records for as long as they are registered
    setup.fused_pairs = use_pair
    rec_np = rec_pos.numpy()[:, None, :]
    ceg.energy_points(setup, rec_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        ceg.energy_points(setup, rec_np)
    barrier()
    ep_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 3

    # ================= latency of single Monte-Carlo steps (montecarlo.jl:490-504 consumer), rank 0 ===========
    latency = None
    if rank == 0:
        def med_us(fn, n=300):
            for _ in range(20):
                fn()
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return float(np.median(ts) * 1e6)
        one = rec_np[:1]
        p64 = rec_np[:64]
        from ceg_b200.grids import _energy_args
        vh, ch, eh, bh, enc_, stat_ = _energy_args(setup, True)
        call1, ov1, oc1 = ceg._lib.energy_points_prepared(vh, ch, eh, bh, setup.atomsidx, setup.charges, one, enc_, stat_)
        call64, _, _ = ceg._lib.energy_points_prepared(vh, ch, eh, bh, setup.atomsidx, setup.charges, p64, enc_, stat_)
        callf, _, _ = ceg._lib.energy_points_prepared(vh, ch, None, bh, setup.atomsidx, setup.charges, one, 0.0, 0.0)
        call1()
        v_ref, c_ref = ceg.energy_points(setup, rec_np[:20000])
        assert ov1[0] == v_ref[0] and abs(oc1[0] - c_ref[0]) <= 1e-9 * abs(c_ref[0]) + 1e-6
        latency = {
            "c_abi_energy_point_1_placement": med_us(call1, 2000),
            "c_abi_energy_points_64_placements": med_us(call64, 1000),
            "c_abi_framework_interactions_1_placement": med_us(callf, 2000),
            "c_abi_note": "ceg_energy_points re-issued on prepared argument buffers (no Python marshalling): what a compiled "
                          "host — the Julia shim's ccall — pays per call; one kernel launch, mapped pinned I/O, flag wait",
            "empty_launch_roundtrip": ceg._lib.measure_launch_roundtrip(dev, 1000),
            "empty_launch_note": "ceg_measure_launch_roundtrip: one EMPTY kernel through the same stream, mapped flag and polling "
                                 "loop, timed inside the library — the floor of any single-launch evaluation on this box; "
                                 "c_abi_* minus this = argument marshalling + the kernel's work",
            "energy_point_1_placement": med_us(lambda: ceg.energy_points(setup, one)),
            "energy_points_64_placements": med_us(lambda: ceg.energy_points(setup, p64)),
            "framework_interactions_1_placement": med_us(lambda: ceg.energy_points(setup, one, reciprocal=False)),
            "interpolate_grid_1_point": med_us(lambda: ceg._lib.interpolate(hv, one[0])),
            "ewald_reciprocal_1_placement": med_us(lambda: ceg._lib.ewald_reciprocal(ef.handle, one, qarr, 0.0, 0.0)),
            "ewald_single_contribution_1_placement": med_us(
                lambda: ceg.single_contribution_ewald(ef, ef.StoreRigidChargeFramework, qarr, one)),
            "unit": "us per call through the host API (median of 300), host positions in, host energies out",
            "note": "the reference's CPU call is ~1-2 us: a serial Markov chain should keep single-point interpolate_grid on the "
                    "CPU and hand batches of trial placements to this path",
        }
    eg_v.close()
    eg_c.close()
    del d_e, d_pos, d_rec, d_rec_out

    # ================= C3 (BASELINE configs[2]): 3 LJ probes + Coulomb fused, x-slab sharded ==================
    inp3, cs3, _ = wl.synthetic_inputs("mfi", 0.15, ("C_ch4", "C_co2", "O_co2"), True)
    nx3, ny3, nz3 = cs3.nodes
    a3, b3 = slabs(nx3, world)[rank]
    gs3 = ceg._lib.make_gridspec(cs3)
    ps3 = ceg.ProbeSystem(None, None, inputs=inp3, device=dev)
    d3 = [torch.empty((8, b3 - a3, ny3, nz3), dtype=torch.float32, device="cuda") for _ in range(4)]
    c3_mine, c3_ms, c3_launches = timed_build(ps3, gs3, a3, b3, [t.data_ptr() for t in d3], max(K, 10), W)
    st3 = ceg._lib.probe_last_stats(ps3.handle)
    fl3, nin3 = flops_per_point(inp3, st3)
    c3_tflops = fl3 * st3["n_points"] / (c3_mine * 1e-3) / 1e12
    h3 = [torch.empty((8, nx3, ny3, nz3), dtype=torch.float32).pin_memory() for _ in range(4)]
    h3np = [t.numpy() for t in h3]

    def c3_e2e():
        p = ceg.ProbeSystem(None, None, inputs=inp3, device=dev)
        ceg._lib.grid_build(p.handle, gs3, a3, b3, lam_v, thr_v, lam_c, thr_c, h3np, None)
        p.close()
    c3_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        c3_e2e()
    barrier()
    c3_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 5
    pts3 = nx3 * ny3 * nz3
    ps3.close()
    del d3, h3, h3np

    # ================= C5 (BASELINE configs[4]): 5000 atoms, 502^3 x 2 grids, then the 10^9-query sweep ========
    c5 = None
    if not args.no_c5:
        inp5, cs5, _ = wl.synthetic_inputs("mof", 0.10, ("C_ch4",), True)
        nx5, ny5, nz5 = cs5.nodes
        sl5 = slabs(nx5, world)
        a5, b5 = sl5[rank]
        gs5 = ceg._lib.make_gridspec(cs5)
        ps5 = ceg.ProbeSystem(None, None, inputs=inp5, device=dev)
        d5 = [torch.empty((8, b5 - a5, ny5, nz5), dtype=torch.float32, device="cuda") for _ in range(2)]
        c5_mine, c5_ms, c5_launches = timed_build(ps5, gs5, a5, b5, [t.data_ptr() for t in d5], 2, 1)
        st5 = ceg._lib.probe_last_stats(ps5.handle)
        fl5, nin5 = flops_per_point(inp5, st5)
        c5_tflops = fl5 * st5["n_points"] / (c5_mine * 1e-3) / 1e12
        del d5
        ps5.close()
        sb5, c5_fused_error = sharded_build(inp5, cs5)
        if sb5 is not None:
            sb5.build(stream)          # warm-up: image list, peer mappings touched
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            sb5.build(stream)
            torch.cuda.synchronize()
            barrier()
            c5_total_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            grids5 = sb5.grids
            sb5.close(barrier=barrier)
        else:                          # NCCL hand-off (see e2e_device.nccl_baseline)
            ps5 = ceg.ProbeSystem(None, None, inputs=inp5, device=dev)
            ho5 = Handoff(ps5, inp5, cs5, (np.inf, 1e-6), pieces=n_pieces)
            for eg in ho5.step():
                eg.close()
            barrier()
            t0 = time.perf_counter()
            grids5 = ho5.step()
            barrier()
            c5_total_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            ps5.close()
            del ho5
        c5_handoff_ms = max(0.0, c5_total_ms - c5_ms)
        q0, q1 = slabs(N_SWEEP, world)[rank]
        for eg in grids5:
            ceg.interpolate_sweep(eg, 10 ** 6, 4)
        barrier()
        t0 = time.perf_counter()
        res5 = [ceg.interpolate_sweep(eg, q1 - q0, 4, first=q0, n_keep=1000 if rank == 0 else 0) for eg in grids5]
        barrier()
        c5_sweep_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        c5_kernel_ms = max_over_ranks(sum(r["kernel_ms"] for r in res5))
        sums = [sum_over_ranks(r["sum"]) for r in res5]
        nblk = sum_over_ranks(float(res5[0]["n_blocked"]))
        mins = [min_over_ranks(r["min"]) for r in res5]
        if rank == 0:
            pos = wl.random_positions(cs5.cell.mat, 1000, 4)
            assert np.array_equal(ceg.interpolate_grid(grids5[0], pos), res5[0]["keep"])
        for eg in grids5:
            eg.close()
        pts5 = nx5 * ny5 * nz5
        c5 = {"workload": WORKLOAD_C5,
              "build": {"value": pts5 / (c5_ms * 1e-3), "unit": "grid points/s", "ms_per_step": c5_ms, "steps": 2,
                        "roofline": {"bound": "fp64", "achieved": c5_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                                     "frac": c5_tflops / fp64_peak if fp64_peak else None,
                                     "algorithmic_flop_per_point": fl5, "pairs_in_cutoff_per_point": nin5,
                                     "kernel": "grid_build_tiled_kernel<256,1,true,true>"},
                        "gpu_launches": int(c5_launches)},
              "build_and_handoff_ms": c5_total_ms, "handoff_ms_beyond_build": c5_handoff_ms, "fused_error": c5_fused_error,
              "handoff_note": "ceg_grid_build_records: every rank's build kernel writes its x-slab as interpolation records into the grids of all ranks (NVLink peer stores); "
                              "no all-gather, no re-layout pass",
              "sweep": {"value": 2 * N_SWEEP / (c5_sweep_ms * 1e-3), "unit": "interpolations/s", "ms": c5_sweep_ms,
                        "kernel_ms_max_over_ranks": c5_kernel_ms, "queries": N_SWEEP, "grids": 2,
                        "roofline": {"bound": "hbm", "achieved": 2 * (q1 - q0) * 264.0 / (c5_kernel_ms * 1e-3) / 1e9 if c5_kernel_ms else None,
                                     "peak": hbm_peak, "unit": "GB/s",
                                     "frac": 2 * (q1 - q0) * 264.0 / (c5_kernel_ms * 1e-3) / 1e9 / hbm_peak if c5_kernel_ms else None,
                                     "algorithmic_bytes_per_query": 264,
                                     "note": "positions generated and energies reduced on the device: 256 B stencil + 8 B; a 4 GB "
                                             "grid is far beyond the L2, the gather runs at the memory system's random-sector rate"},
                        "sum_vdw_K": sums[0], "sum_coulomb": sums[1], "min_vdw_K": mins[0], "blocked": int(nblk)},
              "e2e": {"value": pts5 / (c5_total_ms * 1e-3), "unit": "grid points/s",
                      "note": "grid pair built, gathered and interpolable on every GPU (build + hand-off)",
                      "h2d_bytes_per_step": int(inp5.positions.nbytes + inp5.charges.nbytes + inp5.types.nbytes),
                      "d2h_bytes_per_step": 0}}

    # ================= guest-guest pair energies (energy.jl:397-427), rank 0 =============================
    guest = None
    if rank == 0:
        guest = bench_guest(ceg, wl, dev)

    # ================= CPU baseline (rank 0, N = 1 only): the oracle on the host cores ===============
    cpu = None
    cpu_i = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, cpu_i = cpu_baseline(inp, cset4, cset1, gK, budget_s=args.cpu_budget)

    total_launches = ceg._lib.kernel_launch_count() - launches0
    d2h_total = sum_over_ranks(d2h_bytes)          # collective: every rank calls it
    if rank == 0:
        line = {
            "metric": "grid points/sec (LJ+Ewald FP64)",
            "value": build_value, "unit": "grid points/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": build_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "real CHA_1.4_3b4eeb96 framework fixture (flattened, tests/golden/cha_na.npz); "
                                     "synthetic query positions",
            "config": {"workload": WORKLOAD_BUILD},
            "parallelism": f"x-slab x{world}", "cpus_bound_per_rank": local_cpus, "host": host_info(),
            "l2": "no L2 flush: inputs (<1 MB image list) are re-read by every tile by design; each step writes 1.8 GB of "
                  "output (> 126 MB L2)",
            "e2e": {"value": e2e_value, "unit": "grid points/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_total),
                    "ms_per_step": e2e_ms, "note": "ceg_probe_create + ceg_grid_build into pinned host arrays + free",
                    "breakdown_ms": breakdown},
            "e2e_device": {"value": total_points / (dr_ms * 1e-3), "unit": "grid points/s", "ms_per_step": dr_ms,
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "replicas_identical": bool(chk_same), "fused_error": fused_error,
                           "note": "per step: ceg_grid_build_records of this rank's x-slab — the build kernel writes every node as a "
                                   "finished interpolation record into the grids of all ranks (NVLink peer stores; all-gather and "
                                   "re-layout fused into the kernel) — device synchronize, barrier -> the VdW + Coulomb grid pair is "
                                   "resident and interpolable on every GPU (the reference: .grid files, raspa.jl:430-438); destination "
                                   "grids allocated and IPC-mapped once",
                           "nccl_baseline": {"ms_per_step": nccl_ms, "value": total_points / (nccl_ms * 1e-3),
                                             "sub_slabs_per_rank": n_pieces,
                                             "note": "same hand-off with library collectives: ceg_grid_build_dev in sub-slabs, NCCL "
                                                     "all-gather overlapped with the next sub-slab's build, "
                                                     "ceg_grid_upload_parts_dev"}},
            "roofline": {"bound": "fp64", "achieved": my_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": my_tflops / fp64_peak if fp64_peak else None,
                         "traffic": tr_b * stats["n_points"] if tr_b else None, "traffic_source": tr_b_src,
                         "traffic_unit": "DRAM bytes per step (this rank)",
                         "peak_source": "ceg_measure_fp64_peak (DFMA chains with uniform operands, this run); nominal 37.2",
                         "peak_register_operands": fp64_peak_regops,
                         "frac_of_register_operand_peak": my_tflops / fp64_peak_regops if fp64_peak_regops else None,
                         "algorithmic_flop_per_point": fl_pp, "pairs_in_cutoff_per_point": n_in_pp,
                         "pairs_tested_per_point": stats["pairs_tested"] / max(1, stats["n_points"]),
                         "kernel": "grid_build_tiled_kernel<256,1,true,true>", "kernel_ms": my_ms,
                         "registers": build_regs},
            "gpu_launches": int(build_launches),
            "clocks": clocks,
            "c1": {"workload": "C1: same framework, 0.15 Å (218x204x190 = 8,449,680 points), VdW + Coulomb grids, 1 GPU",
                   "kernel_ms": c1_kernel_ms, "e2e_ms_pinned_host": c1_e2e_ms,
                   "points_per_s_kernel": 8449680 / (c1_kernel_ms * 1e-3)},
            "interp": {"metric": "tricubic interpolations/sec", "value": interp_value, "unit": "interpolations/s",
                       "ms_per_step": best_ms, "steps": KI,
                       "config": {"workload": WORKLOAD_INTERP},
                       "parallelism": f"queries x{world}",
                       "l2": "grids 2 x 270 MB (pair records: 540 MB) + 240 MB of positions exceed the 126 MB L2; no flush",
                       "e2e": {"value": 2 * N_QUERIES / (interp_e2e_ms * 1e-3), "unit": "interpolations/s",
                               "h2d_bytes_per_step": int(N_QUERIES * 24), "d2h_bytes_per_step": int(2 * N_QUERIES * 8),
                               "note": ("ceg_interpolate_pair" if use_pair else "ceg_interpolate_multi") +
                                       ": host positions in, both grids' energies out, chunked on two streams",
                               "ms_per_step": interp_e2e_ms},
                       "roofline": ({"bound": "hbm", "achieved": pair_gbs, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": pair_gbs / hbm_peak, "peak_source": hbm_src,
                                     "traffic": tr_i * nq if tr_i else None, "traffic_source": tr_i_src,
                                     "traffic_unit": "DRAM bytes per launch (this rank)",
                                     "algorithmic_bytes_per_query": 552,
                                     "algorithmic_note": "one launch = both interpolations of a query: 8 nodes x 64 B {VdW, Coulomb} "
                                                         "records + 24 B position + 16 B of results (two separate lookups: 2 x 288 B)",
                                     "kernel": "interpolate_pair_kernel", "registers": pair_regs} if use_pair else
                                    {"bound": "hbm", "achieved": sep_gbs, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": sep_gbs / hbm_peak, "peak_source": hbm_src,
                                     "traffic": tr_i * nq if tr_i else None, "traffic_source": tr_i_src,
                                     "traffic_unit": "DRAM bytes per launch of 10^7/N queries on one grid (this rank)",
                                     "dram_throughput_frac": (tr_i * nq / (sep_ms / 2 * 1e-3) / 1e9 / hbm_peak) if tr_i else None,
                                     "dram_note": "profiled DRAM bytes (64-byte access granularity against 32-byte records) / "
                                                  "this run's launch time / peak: how close the kernel is to the bandwidth "
                                                  "limit with the bytes its access pattern actually moves",
                                     "algorithmic_bytes_per_query": 288, "kernel": "interpolate_kernel<BLOCK>", "registers": sep_regs}),
                       "separate_kernels": {"value": 2 * N_QUERIES / (sep_ms * 1e-3), "ms_per_step": sep_ms,
                                            "achieved_gbs": sep_gbs, "frac": sep_gbs / hbm_peak,
                                            "algorithmic_bytes_per_query": 288, "kernel": "interpolate_kernel<BLOCK> x 2"},
                       "fused_pair_kernel": {"value": 2 * N_QUERIES / (pair_ms * 1e-3), "ms_per_step": pair_ms,
                                             "achieved_gbs": pair_gbs, "frac": pair_gbs / hbm_peak,
                                             "algorithmic_bytes_per_query": 552, "kernel": "interpolate_pair_kernel",
                                             "footprint_bytes": int(2 * 8 * 4 * nx1 * ny1 * nz1)},
                       "gpu_launches": int(interp_launches), "clocks": clocks2,
                       "blocked_fraction": blocked_fraction, "checksum_K": checksum,
                       "cpu_baseline": cpu_i},
            "ewald": {"metric": "reciprocal-space Ewald placements/sec (one Na+ per placement)",
                      "value": n_rec * world / (rec_ms * 1e-3), "unit": "placements/s", "ms_per_step": rec_ms, "steps": KI,
                      "config": {"workload": f"{N_RECIP} random Na+ placements in the C1 cell, k-space (8,8,8), {n_k} k-vectors "
                                             "(compute_ewald, ewald.jl:555-577)"},
                      "roofline": {"bound": "fp64", "achieved": 15.0 * n_k * n_rec / (rec_ms * 1e-3) / 1e12, "peak": fp64_peak,
                                   "unit": "TFLOP/s", "frac": 15.0 * n_k * n_rec / (rec_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                                   "algorithmic_flop_per_atom_k": 15, "kernel": "reciprocal_rows_kernel<1>"},
                      "energy_points_e2e": {"value": n_rec * world / (ep_ms * 1e-3), "unit": "placements/s", "ms_per_step": ep_ms,
                                            "note": "energy_points (grids.jl:311-327 batched): host positions in, host (vdw, coulomb) "
                                                    "out; fused pair interpolation + reciprocal term per placement",
                                            "h2d_bytes_per_step": int(n_rec * world * 24), "d2h_bytes_per_step": int(n_rec * world * 16)}},
            "c3": {"workload": WORKLOAD_C3, "value": 4 * pts3 / (c3_ms * 1e-3), "unit": "grid points/s (x 4 grids)",
                   "ms_per_step": c3_ms, "points_per_s_per_pass": pts3 / (c3_ms * 1e-3),
                   "e2e": {"value": 4 * pts3 / (c3_e2e_ms * 1e-3), "unit": "grid points/s (x 4 grids)", "ms_per_step": c3_e2e_ms,
                           "h2d_bytes_per_step": int(inp3.positions.nbytes + inp3.charges.nbytes + inp3.types.nbytes + inp3.table.nbytes),
                           "d2h_bytes_per_step": int(4 * 8 * pts3 * 4),
                           "note": "ceg_probe_create + ceg_grid_build of the 4 grids into pinned host arrays + free"},
                   "roofline": {"bound": "fp64", "achieved": c3_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                                "frac": c3_tflops / fp64_peak if fp64_peak else None, "algorithmic_flop_per_point": fl3,
                                "pairs_in_cutoff_per_point": nin3, "kernel": "grid_build_tiled_kernel<256,3,true,true>"},
                   "gpu_launches": int(c3_launches)},
            "c5": c5,
            "latency_us": latency,
            "guest": guest,
            "cpu_baseline": cpu,
            "gpu_launches_total": int(total_launches),
            "device": ceg._lib.device_info(dev),
        }
        if args.headline == "interp":
            # same measurements, the interpolation half promoted to the top-level keys
            top = dict(line["interp"])
            build = {k: line[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "e2e", "e2e_device", "roofline",
                                          "gpu_launches", "clocks", "cpu_baseline", "c1")}
            line = {"metric": top.pop("metric"), "value": top.pop("value"), "unit": top.pop("unit"), "n_gpus": world,
                    "steps": top.pop("steps"), "warmup": W, "ms_per_step": top.pop("ms_per_step"), "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": line["data"], **top,
                    "build": build, "ewald": line["ewald"], "c3": line["c3"], "c5": line["c5"],
                    "latency_us": line["latency_us"], "guest": line["guest"],
                    "gpu_launches_total": line["gpu_launches_total"], "device": line["device"], "host": line["host"]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def bench_guest(ceg, wl, dev):
    """single_contribution_vdw for 10^5 trial CO2 placements against 10^3 CO2 guests (3000 atoms) in the C3 cell."""
    import torch
    fw, ff, _ = wl.mfi_shaped()
    names = ["O_co2", "C_co2"]
    tab = ceg.pair_energy_table(ff, names)
    mol = np.array([[0.0, 0.0, 1.149], [0.0, 0.0, 0.0], [0.0, 0.0, -1.149]])
    idx = np.array([0, 1, 0], dtype=np.int32)
    n_guests, n_trials = 1000, 100_000
    cell = np.asarray(fw.mat) * np.array([2.0, 2.0, 2.0])[None, :]
    envc = wl.random_positions(cell, n_guests, 3)
    env_pos = (envc[:, None, :] + mol[None, :, :]).reshape(-1, 3)
    env = ceg.GuestEnvironment(cell, 12.0, tab, env_pos, np.tile(idx, n_guests), np.repeat(np.arange(n_guests, dtype=np.int32), 3))
    trials = wl.random_positions(cell, n_trials, 11)[:, None, :] + mol[None, :, :]
    ceg.single_contribution_vdw(env, idx, trials, -1)           # warm-up at full size: staging buffers allocated
    torch.cuda.synchronize()
    reps = 5
    k_ms = []
    t0 = time.perf_counter()
    for _ in range(reps):
        e = ceg.single_contribution_vdw(env, idx, trials, -1)
        k_ms.append(ceg._lib.guest_env_last_kernel_ms(env.handle))
    ms = (time.perf_counter() - t0) * 1e3 / reps
    kernel_ms = float(np.mean(k_ms))
    pairs = n_trials * 3 * 3 * n_guests
    vol = abs(float(np.linalg.det(cell)))
    frac_in = (4.0 / 3.0) * np.pi * 12.0 ** 3 / vol              # expected share of the pairs inside the cutoff
    flop = pairs * (24.0 + frac_in * 70.0)                       # wrap + fused 3x3 product + norm; LJ + real-space Ewald rule
    peak = ceg._lib.measure_fp64_peak(dev)
    return {"roofline": {"bound": "fp64", "achieved": flop / (kernel_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flop / (kernel_ms * 1e-3) / 1e12 / peak if peak else None,
                         "algorithmic_flop_per_pair_test": 24, "algorithmic_flop_per_in_cutoff_pair": 70,
                         "expected_in_cutoff_fraction": frac_in, "kernel": "guest_energy_fast_kernel", "kernel_ms": kernel_ms,
                         "note": "kernel time from CUDA events on the handle's stream (ceg_guest_env_last_kernel_ms)"},
            "workload": "10^5 trial CO2 placements x 10^3 CO2 guests (3000 environment atoms), LJ + real-space Ewald rules, 12 Å cutoff",
            "ms_per_step": ms, "kernel_ms": kernel_ms, "pair_tests_per_s": pairs / (ms * 1e-3),
            "pair_tests_per_s_kernel": pairs / (kernel_ms * 1e-3), "placements_per_s": n_trials / (ms * 1e-3),
            "checksum_K": float(np.where(np.isfinite(e), e, 0.0).sum()),
            "note": "ms_per_step: host positions in, host energies out (ceg_guest_env_energy, H2D of 7.2 MB of pageable trial "
                    "positions included).  A trial placement is shared by one thread per molecule atom; every (trial atom, guest "
                    "atom) pair goes through a rint-wrap in fractional coordinates and one fused 3x3 product; pairs within 1e-9 A^2 "
                    "of the cutoff are re-measured with the reference's literal sequence; in-cutoff pairs are queued per thread "
                    "and evaluated warp-wide out of line"}


# ---------------------------------------------------------------------------------------------------
def cpu_baseline(inp, cset4, cset1, gK, budget_s=20.0):
    """The reference's CPU algorithm (oracle port: brute force over all supercell atoms per point, one
    pass per grid, probes.jl:71-117) on a bounded sample of the same workload, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    cores = orc.use_all_cores()
    nx, ny, nz = cset4.nodes
    t0 = time.perf_counter()
    orc.grid_build(inp, cset4, nx // 2, nx // 2 + 1, 1, want_f32=True)
    t1 = time.perf_counter() - t0
    n_planes = int(max(1, min(nx, budget_s / max(t1, 1e-3))))
    stride = max(1, nx // n_planes)
    t0 = time.perf_counter()
    orc.grid_build(inp, cset4, 0, nx, stride, want_f32=True)
    dt = time.perf_counter() - t0
    planes = len(range(0, nx, stride))
    cpu = {"value": planes * ny * nz / dt, "unit": "grid points/s", "cores": cores, "kind": "port", **host_info(),
           "sample": f"{planes} of {nx} x planes (every {stride}th) of the same C4 grid, VdW + Coulomb passes, "
                     f"{dt:.1f} s on {cores} OpenMP threads"}
    cpu_i = None
    if gK is not None:
        from ceg_b200 import workloads as wl
        pos = wl.random_positions(cset1.cell.mat, 1_000_000, 1)
        t0 = time.perf_counter()
        orc.interpolate(cset1, gK[0], True, pos)
        orc.interpolate(cset1, gK[1], False, pos)
        dt = time.perf_counter() - t0
        cpu_i = {"value": 2e6 / dt, "unit": "interpolations/s", "cores": cores, "kind": "port",
                 "sample": f"first 10^6 of the 10^7 C2 positions on both grids, 64x64 matrix form, {dt:.2f} s"}
    return cpu, cpu_i


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Julia is not available
    (SURVEY.md §8c), so this is the oracle port, timed with all host threads on bounded samples of the
    product arm's workload.  Rank 0 only."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from ceg_b200 import workloads as wl
    orc.build()
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: size the team from the affinity mask instead
    cores = orc.use_all_cores()
    inp, cset1, _ = wl.load_fixture("cha_na")
    cset4 = wl.respace(cset1, 0.10)
    nx, ny, nz = cset4.nodes
    K, W = args.steps, args.warmup
    t0 = time.perf_counter()
    orc.grid_build(inp, cset4, nx // 2, nx // 2 + 1, 1, want_f32=True)
    t1 = time.perf_counter() - t0
    per_step_s = max(1.0, args.cpu_budget * 3.0 / (K + W))
    n_planes = int(max(1, min(nx, per_step_s / max(t1, 1e-3))))
    stride = max(1, nx // n_planes)
    planes = len(range(0, nx, stride))
    for w in range(W):
        orc.grid_build(inp, cset4, w % stride, nx, stride, want_f32=True)
    t0 = time.perf_counter()
    for k in range(K):
        orc.grid_build(inp, cset4, k % stride, nx, stride, want_f32=True)
    dt = time.perf_counter() - t0
    pts = sum(len(range(k % stride, nx, stride)) for k in range(K)) * ny * nz
    value = pts / dt
    sample = (f"each step = {planes} of {nx} x planes (every {stride}th) of the C4 grid, VdW + Coulomb passes, "
              f"{cores} OpenMP threads")
    line = {"impl": "reference", "metric": "grid points/sec (LJ+Ewald FP64)", "value": value, "unit": "grid points/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "real CHA_1.4_3b4eeb96 framework fixture (flattened, tests/golden/cha_na.npz)",
            "config": {"workload": WORKLOAD_BUILD},
            "parallelism": f"OpenMP x{cores}", "host": host_info(),
            "cpu_baseline": {"value": value, "unit": "grid points/s", "cores": cores, "kind": "port", "sample": sample, **host_info()},
            "e2e": {"value": value, "unit": "grid points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = CrystalEnergyGrids.jl CPU algorithm restated in C/OpenMP (oracle/ceg_oracle.c); "
                    "Julia is not installed in this image"}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was
    redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ceg", choices=["ceg", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 block (2 x 4 GB grids + 10^9-query sweep)")
    ap.add_argument("--headline", default="build", choices=["build", "interp"],
                    help="which half of BASELINE.json's metric the top-level keys carry: grid construction on C4 (default) "
                         "or tricubic interpolation on C2 (configs[1]); the other half is always in the same line")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ceg" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
