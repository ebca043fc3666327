"""GPU experiment: throughput of the reciprocal-space Ewald kernel (compute_ewald batched over placements)
and of the fused energy_points, on the C1 k-space (8,8,8), 1368 k-vectors.
usage: exp_recip.py [n_placements]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
inp, cset, ex = wl.load_fixture("cha_na")
n_unit = len(ex["site_charge"])
fw = ceg.RASPASystem(ex["unit_mat"], ex["positions"][:n_unit], ["X"] * n_unit, np.ones(n_unit), ex["charges"][:n_unit], False)
ef = ceg.initialize_ewald(fw)
stream = torch.cuda.current_stream().cuda_stream
geom = {1: np.zeros((1, 3)), 3: np.array([[0, 0, 1.149], [0, 0, 0.0], [0, 0, -1.149]])}
charges = {1: np.array([0.9094]), 3: np.array([-0.3256, 0.6512, -0.3256])}
for na in (1, 3):
    centers = wl.random_positions(fw.mat, n, 5)
    pos = centers[:, None, :] + geom[na][None]
    d_xyz = torch.from_numpy(pos).cuda()
    d_out = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        ceg._lib.ewald_reciprocal_dev(ef.handle, n, charges[na], d_xyz.data_ptr(), 1.5, -2.5, d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        ceg._lib.ewald_reciprocal_dev(ef.handle, n, charges[na], d_xyz.data_ptr(), 1.5, -2.5, d_out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nk = ef.kspace.num_kvecs
    print(f"reciprocal: {n} placements x {na} atoms, {nk} k-vectors: {ms:.2f} ms, {n/ms/1e3:.2f} M placements/s, "
          f"{n*na*nk/ms/1e6:.1f} G (atom,k) terms/s, checksum {float(d_out.sum()):.9e}")

# ---- fused energy_points / energy_grid through the host API (grids.jl:311-424), wall clock -----------------
import time
nx, ny, nz = cset.nodes
ps = ceg.ProbeSystem(None, None, inputs=inp)
outs = [np.empty((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
lam_v, thr_v = ceg.vdw_scaling(); lam_c, thr_c = ceg.coulomb_scaling()
ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cset), 0, nx, lam_v, thr_v, lam_c, thr_c, outs, None)
ps.close()
k = np.float32(ceg.GRID_TO_KELVIN)
gv = ceg.EnergyGrid(cset, inp.supercell, np.inf, True, (outs[0].astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32))
gc = ceg.EnergyGrid(cset, inp.supercell, 1e-6, True, (outs[1].astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32))
qna = float(ex["probe_charge"])
na_sys = ceg.RASPASystem(np.diag([np.inf] * 3), np.zeros((1, 3)), ["Na"], np.ones(1), np.array([qna]), True)
setup = ceg.CrystalEnergySetup(fw, na_sys, gc, np.array([qna]), [gv], [0], ef, None, ceg.BlockFile(cset))
pos = wl.random_positions(fw.mat, n, 6)[:, None, :]
ceg.energy_points(setup, pos[:1000])
for rep in range(2):
    t0 = time.perf_counter(); v, c = ceg.energy_points(setup, pos); t1 = time.perf_counter()
    print(f"energy_points (host API, {n} Na+ placements): {1e3*(t1-t0):.2f} ms, {n/(t1-t0)/1e6:.1f} M points/s")
for rep in range(2):
    t0 = time.perf_counter(); eg = ceg.energy_grid(setup, 0.3); t1 = time.perf_counter()
    print(f"energy_grid(0.3 Å) {eg.shape}: {1e3*(t1-t0):.1f} ms, min {eg.min():.6f} at {np.unravel_index(eg.argmin(), eg.shape)}")
