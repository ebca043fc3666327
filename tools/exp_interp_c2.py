"""GPU experiment / ncu target: the C2 interpolation kernels on grids of the C1 size (random data): interpolate_kernel<BLOCK> on
one grid and interpolate_pair_kernel on the fused pair, 10^7 queries each.  usage: exp_interp_c2.py [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
inp, cset, ex = wl.load_fixture("cha_na")
nx, ny, nz = cset.nodes
rng = np.random.default_rng(0)
g = [rng.standard_normal((8, nx, ny, nz)).astype(np.float32) for _ in range(2)]
ev = ceg.EnergyGrid(cset, (1, 1, 1), np.inf, True, g[0]); ec = ceg.EnergyGrid(cset, (1, 1, 1), 1e-6, True, g[1])
hv, hc = ev.handle, ec.handle
hp = ev.pair_with(ec)
n = 10_000_000
pos = torch.from_numpy(wl.random_positions(cset.cell.mat, n, 1)).cuda()
o = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
stream = torch.cuda.current_stream().cuda_stream
for name, fn in (("interpolate_kernel<BLOCK>", lambda: ceg._lib.interpolate_dev(hv, n, pos.data_ptr(), o[0].data_ptr(), stream)),
                 ("interpolate_pair_kernel", lambda: ceg._lib.interpolate_pair_dev(hp, n, pos.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), stream))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / reps:.3f} ms per 10^7 queries")
