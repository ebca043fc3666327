"""GPU experiment driver: the fused Na+ build kernel, or the VdW-only + Coulomb-only pair of passes.
usage: exp_build.py [spacing] [reps] [fused|split]"""
import dataclasses, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

spacing = float(sys.argv[1]) if len(sys.argv) > 1 else 0.10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "fused"
inp, cset, ex = wl.load_fixture("cha_na")
cs = wl.respace(cset, spacing)
nx, ny, nz = cs.nodes
gs = ceg._lib.make_gridspec(cs)
lam_v, thr_v = ceg.vdw_scaling(); lam_c, thr_c = ceg.coulomb_scaling()
d = [torch.empty((8, nx, ny, nz), dtype=torch.float32, device="cuda") for _ in range(2)]
stream = torch.cuda.current_stream().cuda_stream
if mode == "fused":
    handles = [(ceg.ProbeSystem(None, None, inputs=inp), [d[0].data_ptr(), d[1].data_ptr()])]
else:
    handles = [(ceg.ProbeSystem(None, None, inputs=dataclasses.replace(inp, want_coulomb=False, charges=None)), [d[0].data_ptr()]),
               (ceg.ProbeSystem(None, None, inputs=dataclasses.replace(inp, table=inp.table[:0])), [d[1].data_ptr()])]

def step():
    for ps, ptrs in handles:
        ceg._lib.grid_build_dev(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, ptrs, stream)

for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
st = ceg._lib.probe_last_stats(handles[-1][0].handle)
print(f"{mode} spacing {spacing}: {nx}x{ny}x{nz} = {nx*ny*nz} points, {ms:.2f} ms/build, {nx*ny*nz/ms/1e3:.1f} M points/s, "
      f"pairs in cutoff/point {st['pairs_in_cutoff']/st['n_points']:.1f}, tested {st['pairs_tested']/st['n_points']:.1f}, "
      f"checksum {float(d[0][0].double().clamp(-1e6,1e6).sum()):.6e} {float(d[1][0].double().sum()):.6e}")
