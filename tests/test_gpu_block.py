"""GPU parity of the blocking masks (parse_blockfile, BlockFile(g)) and their use in
energy_point (coordinates.jl:83-167, grids.jl:188-204, 311-314)."""
import os

import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _write_block(tmp_path, centers_frac, radii):
    p = tmp_path / "x.block"
    with open(p, "w") as f:
        f.write(f"{len(radii)}\n")
        for c, r in zip(centers_frac, radii):
            f.write(f"{c[0]}   {c[1]}   {c[2]}   {r}\n")
    return str(p)


def test_blocking_sphere_across_periodic_boundary(cuda_lib, orc, tmp_path):
    """test/runtests.jl:269-272: CIT7block, 3 Å sphere at fractional (0.1,0.2,0.3): both listed points are
    blocked; the device mask equals the oracle's bit for bit (triclinic cell, literal periodic distance)."""
    z = np.load(os.path.join(wl.GOLDEN_DIR, "cit7block.npz"))
    cset = ceg.GridCoordinatesSetup.from_cell(z["unit_mat"], float(z["spacing"]))
    path = _write_block(tmp_path, [(0.1, 0.2, 0.3)], [3.0])
    blk = ceg.parse_blockfile(path, cset)
    want = orc.parse_blockfile(path, cset)
    assert not blk.empty and np.array_equal(blk.block, want)
    c, r2 = ceg.coordinates.read_blockfile(path, cset)
    assert np.allclose(c, z["centers"], atol=1e-12) and np.allclose(r2, z["radius2"])
    for p in z["blocked_points"]:
        assert blk[p]
    # energy_point returns (1e100, 0) there, through the fused device path
    nx, ny, nz = cset.nodes
    g = np.zeros((8, nx, ny, nz), dtype=np.float32)
    fw = ceg.RASPASystem(z["unit_mat"], np.zeros((0, 3)), [], np.zeros(0), np.zeros(0), False)
    ar = ceg.RASPASystem(np.diag([np.inf] * 3), np.zeros((1, 3)), ["Ar"], np.ones(1), np.zeros(1), True)
    gv = ceg.EnergyGrid(cset, (2, 3, 3), np.inf, True, g)
    setup = ceg.CrystalEnergySetup(fw, ar, ceg.EnergyGrid.zero(), np.zeros(1), [gv], [0], ceg.EwaldFramework.empty(fw.mat),
                                   None, blk)
    for p in z["blocked_points"]:
        assert ceg.energy_point(setup, p[None]) == (1e100, 0.0)
    assert ceg.energy_point(setup, np.array([[6.0, 5.0, 4.5]])) == (0.0, 0.0)
    pos = wl.random_positions(cset.cell.mat, 50000, 3) * 2.0 - 5.0
    v, _ = ceg.energy_points(setup, pos[:, None, :])
    assert np.array_equal(v == 1e100, orc.block_lookup(cset, blk.block, pos))
    gv.close()


def test_several_spheres_and_empty(cuda_lib, orc, tmp_path):
    cset = ceg.GridCoordinatesSetup.from_cell(ceg.mat_from_parameters((9.0, 8.0, 7.0), (75, 95, 110)), 0.35)
    path = _write_block(tmp_path, [(0.02, 0.5, 0.97), (0.6, 0.01, 0.4), (0.5, 0.5, 0.5)], [1.7, 2.4, 0.9])
    assert np.array_equal(ceg.parse_blockfile(path, cset).block, orc.parse_blockfile(path, cset))
    assert ceg.parse_blockfile(_write_block(tmp_path, [], []), cset).empty


def test_blockfile_from_grid(cuda_lib, orc):
    """BlockFile(g) (grids.jl:188-204) on the device == numpy restatement."""
    cset = ceg.GridCoordinatesSetup.from_cell(np.diag([6.0, 5.0, 7.0]), 0.25)
    nx, ny, nz = cset.nodes
    rng = np.random.default_rng(2)
    g = np.zeros((8, nx, ny, nz), dtype=np.float32)
    g[0] = np.where(rng.random((nx, ny, nz)) < 0.03, 7e6, rng.standard_normal((nx, ny, nz)) * 1e3).astype(np.float32)
    g[0, nx - 1] = 9e6            # the last x plane is never scanned as a lower corner
    eg = ceg.EnergyGrid(cset, (1, 1, 1), np.inf, True, g)
    b = ceg.blockfile_from_grid(eg)
    assert np.array_equal(b.block, orc.blockfile_from_values(g[0]))
    g0 = np.zeros_like(g)
    assert ceg.blockfile_from_grid(ceg.EnergyGrid(cset, (1, 1, 1), np.inf, True, g0)).empty
