"""GPU experiment: phases of the end-to-end build step (handle create / build+D2H / free)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl
inp, cset, ex = wl.load_fixture("cha_na")
cs = wl.respace(cset, 0.10)
nx, ny, nz = cs.nodes
gs = ceg._lib.make_gridspec(cs)
lam_v, thr_v = ceg.vdw_scaling(); lam_c, thr_c = ceg.coulomb_scaling()
host = [torch.empty((8, nx, ny, nz), dtype=torch.float32).pin_memory() for _ in range(2)]
hn = [h.numpy() for h in host]
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    t1 = time.perf_counter()
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, hn, None)
    t2 = time.perf_counter()
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, hn, None)
    t3 = time.perf_counter()
    ps.close()
    t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms  build#1 {1e3*(t2-t1):.1f} ms  build#2 {1e3*(t3-t2):.1f} ms  free {1e3*(t4-t3):.1f} ms")
tm = ceg._lib.probe_last_timing
ps = ceg.ProbeSystem(None, None, inputs=inp)
for it in range(6):
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, hn, None)
    print("timing (list, kernel span, copy tail, call) ms:", {k: round(v, 3) for k, v in tm(ps.handle).items()})
ps.close()
print("checksum", float(hn[0][0].astype(np.float64).sum()), float(hn[1][0].astype(np.float64).sum()))
