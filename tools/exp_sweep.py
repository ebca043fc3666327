"""GPU experiment: the 10^9-query interpolation sweep of SURVEY §8d (positions generated on the device,
results reduced on the device) on a C1-sized (270 MB) and a C5-sized (4.05 GB) grid.
usage: exp_sweep.py [n_queries]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
inp, cset1, _ = wl.load_fixture("cha_na")
cases = [("C1 CHA 0.15 Å", cset1), ("C5 50 Å cubic 0.10 Å", ceg.GridCoordinatesSetup.from_cell(np.diag([50.0, 50.0, 50.0]), 0.10))]
for name, cs in cases:
    nx, ny, nz = cs.nodes
    rng = np.random.default_rng(1)
    grid = rng.standard_normal((8, nx, ny, nz), dtype=np.float32)
    grid[0] *= 1000
    grid[0][rng.random((nx, ny, nz)) < 0.2] = 2.4e7          # blocked corners, as in a real VdW grid
    eg = ceg.EnergyGrid(cs, (1, 1, 1), np.inf, True, grid)
    eg.handle
    ceg.interpolate_sweep(eg, 10**7, 4)
    for rep in range(2):
        t0 = time.perf_counter()
        r = ceg.interpolate_sweep(eg, n, 4, n_keep=10**6)
        wall = time.perf_counter() - t0
        print(f"{name}: grid {nx}x{ny}x{nz} ({grid.nbytes/1e9:.2f} GB), {n:.0e} queries: kernel {r['kernel_ms']:.1f} ms "
              f"({n/r['kernel_ms']/1e6:.2f} G/s, {288*n/r['kernel_ms']/1e9:.2f} TB/s algorithmic), wall {wall*1e3:.0f} ms, "
              f"blocked {r['n_blocked']/n:.4f}, sum {r['sum']:.9e}, min {r['min']:.6f}")
    pos = wl.random_positions(cs.cell.mat, 10**6, 4)
    assert np.array_equal(ceg.interpolate_grid(eg, pos), r["keep"])
    eg.close()
    del grid
