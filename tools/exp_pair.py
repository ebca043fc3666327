"""GPU experiment: fused {VdW, Coulomb} pair records against two separate lookups, by grid size.
usage: exp_pair.py [spacings...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

spacings = [float(x) for x in sys.argv[1:]] or [0.3, 0.2, 0.17, 0.15]
inp, cset, ex = wl.load_fixture("cha_na")
stream = torch.cuda.current_stream().cuda_stream
n = 10_000_000
for sp in spacings:
    cs = wl.respace(cset, sp)
    nx, ny, nz = cs.nodes
    rng = np.random.default_rng(1)
    g = [rng.standard_normal((8, nx, ny, nz)).astype(np.float32) for _ in range(2)]
    ev = ceg.EnergyGrid(cs, (1, 1, 1), np.inf, True, g[0]); ec = ceg.EnergyGrid(cs, (1, 1, 1), 1e-6, True, g[1])
    hv, hc = ev.handle, ec.handle
    hp = ev.pair_with(ec)
    pos = torch.from_numpy(wl.random_positions(cs.cell.mat, n, 1)).cuda()
    o = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(4)]
    def sep():
        ceg._lib.interpolate_dev(hv, n, pos.data_ptr(), o[0].data_ptr(), stream)
        ceg._lib.interpolate_dev(hc, n, pos.data_ptr(), o[1].data_ptr(), stream)
    def pair():
        ceg._lib.interpolate_pair_dev(hp, n, pos.data_ptr(), o[2].data_ptr(), o[3].data_ptr(), stream)
    res = []
    for fn in (sep, pair):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    assert torch.equal(o[0], o[2]) and torch.equal(o[1], o[3])
    mb = 8 * 4 * nx * ny * nz / 1e6
    print(f"spacing {sp}: grid {nx}x{ny}x{nz} = {mb:.0f} MB each (layout {ceg._lib.grid_layout(hv)}); separate {res[0]:.3f} ms, pair {res[1]:.3f} ms per 10^7 queries on both grids")
    ev.close(); ec.close()
