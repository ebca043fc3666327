"""GPU experiment: per-call latency of the batched evaluation entry points for small batches (Monte-Carlo steps)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

inp, cset, ex = wl.load_fixture("cha_na")
cs = wl.respace(cset, 0.3)
nx, ny, nz = cs.nodes
n_unit = len(ex["site_charge"])
fw = ceg.RASPASystem(ex["unit_mat"], ex["positions"][:n_unit], ["X"] * n_unit, np.ones(n_unit), ex["charges"][:n_unit], False)
ps = ceg.ProbeSystem(None, None, inputs=inp)
outs = [np.empty((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
lam_v, thr_v = ceg.vdw_scaling(); lam_c, thr_c = ceg.coulomb_scaling()
ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cs), 0, nx, lam_v, thr_v, lam_c, thr_c, outs, None)
ps.close()
K = ceg.GRID_TO_KELVIN
gv = ceg.EnergyGrid(cs, inp.supercell, np.inf, True, (outs[0].astype(np.float64) * K).astype(np.float32))
gc = ceg.EnergyGrid(cs, inp.supercell, 1e-6, True, (outs[1].astype(np.float64) * K).astype(np.float32))
ef = ceg.initialize_ewald(fw)
q = float(ex["probe_charge"])
na = ceg.RASPASystem(np.diag([np.inf] * 3), np.zeros((1, 3)), ["Na"], np.ones(1), np.array([q]), True)
setup = ceg.CrystalEnergySetup(fw, na, gc, np.array([q]), [gv], [0], ef, None, ceg.BlockFile(cs))
for n in (1, 100, 10_000, 100_000):
    pos = wl.random_positions(fw.mat, n, 3)[:, None, :]
    for fn, name in ((lambda: ceg.energy_points(setup, pos), "energy_points"),
                     (lambda: ceg.interpolate_grid(gv, pos[:, 0]), "interpolate_grid"),
                     (lambda: ceg._lib.ewald_reciprocal(ef.handle, pos, np.array([q]), 0.0, 0.0), "ewald_reciprocal")):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        reps = 200
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        print(f"{name:18s} n = {n:6d}: {dt*1e6:8.1f} us per call")
