"""GPU experiment: interpolation throughput per device record layout (run under gpurun)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

inp, cset, ex = wl.load_fixture("cha_na")
if len(sys.argv) > 1:
    cset = wl.respace(cset, float(sys.argv[1]))
nx, ny, nz = cset.nodes
print("grid", nx, ny, nz, nx * ny * nz * 32 / 1e6, "MB")
rng = np.random.default_rng(0)
grid = rng.standard_normal((8, nx, ny, nz)).astype(np.float32)
n = 10_000_000
pos = torch.from_numpy(wl.random_positions(cset.cell.mat, n, 1)).cuda()
out = torch.empty(n, dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ref = None
for layout in ((0, 3, 1, 2) if len(sys.argv) < 3 else [int(v) for v in sys.argv[2].split(',')]):
    ceg._lib.set_interp_layout(layout)
    eg = ceg.EnergyGrid(cset, (1, 1, 1), 1e-6, True, grid)
    h = eg.handle
    for _ in range(3):
        ceg._lib.interpolate_dev(h, n, pos.data_ptr(), out.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ceg._lib.interpolate_dev(h, n, pos.data_ptr(), out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    o = out.cpu().numpy().copy()
    if ref is None: ref = o
    print(f"layout {layout}: {ms:.3f} ms  {n/ms/1e6:.2f} G interp/s  {n*288/ms/1e6:.0f} GB/s algorithmic  same={np.array_equal(o, ref)}")
    eg.close()
