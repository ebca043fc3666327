#!/usr/bin/env python
"""Extract the SASS of one kernel from a .so: tools/sass_fun.py lib.so 'grid_build_tiled_kernelILi256ELi1ELb1ELb1' > out.sass
and print a summary (instruction mix, spills, where the local-memory traffic sits)."""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat in name:
        lines = [l for l in b.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
        ins = []
        for l in lines:
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        if len(sys.argv) > 3:
            open(sys.argv[3], "w").write("\n".join(f"{a:06x}  {t}" for a, t in ins))
        import collections
        mix = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for _, t in ins)
        print(name, len(ins), "instructions")
        print(" ".join(f"{k}:{v}" for k, v in mix.most_common(25)))
        break
