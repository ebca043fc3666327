"""GPU parity of tricubic interpolation against the CPU oracle (64x64 matrix form)."""
import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import workloads as wl

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # north_star: interpolated energies within 1e-9 relative


def _random_grid(cset, seed, hot_fraction=0.0):
    rng = np.random.default_rng(seed)
    nx, ny, nz = cset.nodes
    g = rng.standard_normal((8, nx, ny, nz)).astype(np.float32)
    g[0] *= 1000.0
    if hot_fraction:
        hot = rng.random((nx, ny, nz)) < hot_fraction
        g[0][hot] = np.float32(2.4e7)
    return g


def _check(orc, cset, grid, is_vdw, pos, what):
    eg = ceg.EnergyGrid(cset, (1, 1, 1), np.inf if is_vdw else 1e-6, True, grid)
    got = ceg.interpolate_grid(eg, pos)
    want = orc.interpolate(cset, grid, is_vdw, pos)
    blocked = want == 1e100
    assert np.array_equal(got == 1e100, blocked), f"{what}: blocked masks differ"
    ok = ~blocked
    scale = np.maximum(np.abs(want[ok]), 1e-3 * np.abs(grid[0]).mean() if not is_vdw else 1.0)
    err = np.abs(got[ok] - want[ok]) / scale
    assert err.max() <= RTOL, f"{what}: worst rel err {err.max():.3e}"
    eg.close()
    return blocked.mean()


@pytest.fixture(params=[0, 1, 2, 3, 4], ids=["node", "quad", "cell", "pair", "block"])
def layout(request, cuda_lib):
    """Every device record layout (include/ceg_cuda.h: ceg_set_interp_layout) must give the same answers."""
    ceg._lib.set_interp_layout(request.param)
    yield request.param
    ceg._lib.set_interp_layout(-1)


@pytest.mark.parametrize("shape", ["triclinic", "ortho", "rhombo"])
def test_random_grids(cuda_lib, orc, shape, layout):
    mats = {"triclinic": ceg.mat_from_parameters((13.0187, 11.2063, 9.3758), (92.82, 107.2, 103.26)),
            "ortho": np.diag([7.3, 9.1, 5.2]),
            "rhombo": ceg.mat_from_parameters((9.0, 9.0, 9.0), (94.07, 94.07, 94.07))}
    cset = ceg.GridCoordinatesSetup.from_cell(mats[shape], 0.41)
    grid = _random_grid(cset, 5, hot_fraction=0.02)
    # positions inside, outside (wrapped) and far outside the cell
    pos = np.concatenate([wl.random_positions(cset.cell.mat, 20000, 21),
                          wl.random_positions(cset.cell.mat, 2000, 22) * 7.0 - 30.0])
    frac_blocked = _check(orc, cset, grid, True, pos, shape + " vdw")
    assert 0.02 < frac_blocked < 0.5
    _check(orc, cset, grid, False, pos, shape + " coulomb")


def test_nodes_edges_and_empty(cuda_lib, orc, layout):
    """Exact nodes (r = 0), the far faces of the bounding box (p1 clamps, grids.jl:218) and n = 0."""
    cset = ceg.GridCoordinatesSetup.from_cell(np.diag([6.0, 5.0, 4.0]), 0.5)
    grid = _random_grid(cset, 9)
    eg = ceg.EnergyGrid(cset, (1, 1, 1), 1e-6, True, grid)
    assert ceg._lib.grid_layout(eg.handle) == layout
    nx, ny, nz = cset.nodes
    idx = np.array([(i, j, k) for i in (0, 3, nx - 2) for j in (0, 5, ny - 2) for k in (0, 2, nz - 2)])
    pos = np.array([ceg.abc_to_xyz(cset, *t) for t in idx]) + 1e-9   # just inside the cell of node t
    got = ceg.interpolate_grid(eg, pos)
    want = grid[0][idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float64)
    assert np.allclose(got, want, rtol=0, atol=2e-5)
    assert np.allclose(got, orc.interpolate(cset, grid, False, pos), rtol=1e-9, atol=1e-9)
    # points on the cell faces: wrap to 0 / stay at the top face
    face = np.array([[6.0, 2.0, 1.0], [0.0, 0.0, 0.0], [5.999999999, 4.999999999, 3.999999999], [-1e-12, 2.0, 3.0]])
    assert np.allclose(ceg.interpolate_grid(eg, face), orc.interpolate(cset, grid, False, face), rtol=1e-9, atol=1e-9)
    assert ceg.interpolate_grid(eg, np.zeros((0, 3))).shape == (0,)
    assert isinstance(ceg.interpolate_grid(eg, np.array([1.0, 2.0, 3.0])), float)
    eg.close()
    assert ceg.interpolate_grid(ceg.EnergyGrid.zero(), np.array([1.0, 2.0, 3.0])) == 0.0
    with pytest.raises(ValueError):
        ceg.interpolate_grid(ceg.EnergyGrid.invalid(), np.array([1.0, 2.0, 3.0]))


def test_interpolate_multi_pipelined(cuda_lib, orc):
    """ceg_interpolate_multi: several grids at the same positions, more queries than one pipeline chunk."""
    cset = ceg.GridCoordinatesSetup.from_cell(ceg.mat_from_parameters((9.0, 8.0, 7.0), (85, 95, 100)), 0.5)
    g1, g2 = _random_grid(cset, 1, hot_fraction=0.05), _random_grid(cset, 2)
    e1 = ceg.EnergyGrid(cset, (1, 1, 1), np.inf, True, g1)
    e2 = ceg.EnergyGrid(cset, (1, 1, 1), 1e-6, True, g2)
    pos = wl.random_positions(cset.cell.mat, 2_300_000, 77)
    both = ceg._lib.interpolate_multi([e1.handle, e2.handle], pos)
    assert np.array_equal(both[0], ceg.interpolate_grid(e1, pos)) and np.array_equal(both[1], ceg.interpolate_grid(e2, pos))
    sub = slice(0, 50000)
    want = orc.interpolate(cset, g1, True, pos[sub])
    assert np.array_equal(both[0][sub] == 1e100, want == 1e100)
    ok = want != 1e100
    assert (np.abs(both[0][sub][ok] - want[ok]) <= 1e-9 * np.abs(want[ok]) + 1e-6).all()
    e1.close(); e2.close()


def test_c2_on_built_grids_and_goldens(cuda_lib, orc, tmp_path):
    """C1 -> C2 end to end on the device path: build the Ar grid of CHA+Na and the Na+/Coulomb grids of
    CHA at 0.15 Å, write/parse the .grid files, interpolate; check the reference's own known
    answers (test/runtests.jl:38,44) and 2e5 random positions against the oracle on the same
    Float32 grids."""
    import json, os
    gold = json.load(open(os.path.join(wl.GOLDEN_DIR, "goldens.json")))["reference_values"]
    # --- Ar in CHA+Na: minimum of energy_grid(0.3 Å) at lattice index (52,50,91)
    inp, cset, ex = wl.load_fixture("chana_ar")
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    nx, ny, nz = cset.nodes
    out = [np.empty((8, nx, ny, nz), dtype=np.float32)]
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cset), 0, nx, lam_v, thr_v, lam_c, thr_c, out, None)
    ps.close()
    path = str(tmp_path / "ar.grid")
    ceg.write_grid(path, cset, inp.supercell, out[0])
    eg = ceg.parse_grid(path, False)            # cell matrix read back from the file tail
    assert np.array_equal(eg.csetup.dims, cset.dims) and np.allclose(eg.csetup.cell.mat, cset.cell.mat, rtol=0, atol=0)
    mat = cset.cell.mat
    nums = [int(np.floor(np.linalg.norm(mat[:, c]) / 0.3)) + 1 for c in range(3)]
    p = (51 * mat[:, 0] / nums[0] + 49 * mat[:, 1] / nums[1]) + 90 * mat[:, 2] / nums[2]
    e = ceg.interpolate_grid(eg, p)
    assert abs(e - gold["ar_chana_min"]) <= 1e-9 * abs(gold["ar_chana_min"]), (e, gold["ar_chana_min"])
    assert ceg._lib.grid_layout(eg.handle) == 4        # 270 MB grid: 2x2x2-node bricks by default
    pos = wl.random_positions(mat, 200000, 1)
    got = ceg.interpolate_grid(eg, pos)
    want = orc.interpolate(eg.csetup, eg.grid, True, pos)
    assert np.array_equal(got == 1e100, want == 1e100)
    ok = want != 1e100
    assert (np.abs(got[ok] - want[ok]) <= 1e-9 * np.abs(want[ok]) + 1e-6).all()
    eg.close()
    # --- Na+ in CHA at the origin: vdw golden (runtests.jl:44)
    inp, cset, ex = wl.load_fixture("cha_na")
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    nx, ny, nz = cset.nodes
    outs = [np.empty((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cset), 0, nx, lam_v, thr_v, lam_c, thr_c, outs, None)
    ps.close()
    pv, pc = str(tmp_path / "na.grid"), str(tmp_path / "coul.grid")
    ceg.write_grid(pv, cset, inp.supercell, outs[0])
    ceg.write_grid(pc, cset, inp.supercell, outs[1], ewald_precision=1e-6)
    gv = ceg.parse_grid(pv, False, cset.cell.mat)
    gc = ceg.parse_grid(pc, True, cset.cell.mat)
    assert gc.ewald_precision == 1e-6 and gv.ewald_precision == np.inf
    v0 = ceg.interpolate_grid(gv, np.zeros(3))
    assert abs(v0 - gold["na_cha_origin_vdw"]) <= 1e-9 * abs(gold["na_cha_origin_vdw"])
    gv.close(); gc.close()


def test_interpolate_sweep_device_generated_positions(cuda_lib, orc):
    """ceg_interpolate_sweep: positions generated on the device equal workloads.random_positions bit
    for bit, so the kept energies equal the batched ceg_interpolate of the host-generated positions exactly; the
    device reduction (sum, min, blocked count) matches numpy; a split sweep adds up; and the oracle agrees."""
    inp, cset, _ = wl.load_fixture("cit7_multi")          # triclinic cell
    cs = wl.respace(cset, 0.5)
    grid = _random_grid(cs, 5, hot_fraction=0.02)
    n, seed = 300_000, 4
    for is_vdw in (True, False):
        eg = ceg.EnergyGrid(cs, (1, 1, 1), np.inf if is_vdw else 1e-6, True, grid)
        pos = wl.random_positions(cs.cell.mat, n, seed)
        want = ceg.interpolate_grid(eg, pos)
        r = ceg.interpolate_sweep(eg, n, seed, n_keep=n)
        assert np.array_equal(r["keep"], want)
        ok = want != 1e100
        assert r["n_blocked"] == int((~ok).sum()) and (is_vdw == bool(r["n_blocked"]))
        assert r["min"] == want[ok].min()
        assert abs(r["sum"] - want[ok].sum()) <= 1e-12 * np.abs(want[ok]).sum()
        # same sweep in two pieces (how ranks shard it), and run-to-run determinism
        a = ceg.interpolate_sweep(eg, 100_000, seed, first=0, n_keep=10)
        b = ceg.interpolate_sweep(eg, n - 100_000, seed, first=100_000, n_keep=10)
        assert np.array_equal(b["keep"], want[100_000:100_010])
        assert a["n_blocked"] + b["n_blocked"] == r["n_blocked"] and min(a["min"], b["min"]) == r["min"]
        assert abs(a["sum"] + b["sum"] - r["sum"]) <= 1e-12 * np.abs(want[ok]).sum()
        assert ceg.interpolate_sweep(eg, n, seed)["sum"] == r["sum"]
        # oracle on a sample
        o = orc.interpolate(cs, grid, is_vdw, pos[:5000])
        assert np.array_equal(o == 1e100, r["keep"][:5000] == 1e100)
        m = o != 1e100
        assert (np.abs(r["keep"][:5000][m] - o[m]) <= RTOL * np.maximum(np.abs(o[m]), 1.0)).all()
        eg.close()
    z = ceg.interpolate_sweep(ceg.EnergyGrid.zero(), 10, 1, n_keep=4)
    assert z["sum"] == 0.0 and z["n_blocked"] == 0 and len(z["keep"]) == 4


def test_grid_upload_from_device_equals_file_round_trip(cuda_lib, tmp_path):
    """ceg_grid_upload_dev: a grid left on the device by ceg_grid_build_dev becomes an interpolation handle with
    parse_grid's unit conversion applied on the device — same energies, bit for bit, as build -> write_grid ->
    parse_grid -> upload."""
    import torch
    inp, cset, _ = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 0.5)
    nx, ny, nz = cs.nodes
    gs = ceg._lib.make_gridspec(cs)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    d = [torch.empty((8, nx, ny, nz), dtype=torch.float32, device="cuda") for _ in range(2)]
    ceg._lib.grid_build_dev(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, [t.data_ptr() for t in d],
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = [t.cpu().numpy() for t in d]
    ps.close()
    pos = wl.random_positions(cs.cell.mat, 50000, 3)
    for g, (prec, coul) in enumerate([(np.inf, False), (1e-6, True)]):
        path = str(tmp_path / f"g{g}.grid")
        ceg.write_grid(path, cs, inp.supercell, host[g], ewald_precision=prec if coul else None)
        via_file = ceg.parse_grid(path, coul)
        on_dev = ceg.EnergyGrid.from_device(cs, inp.supercell, prec, d[g].data_ptr())
        assert np.array_equal(ceg.interpolate_grid(on_dev, pos), ceg.interpolate_grid(via_file, pos))
        on_dev.close(); via_file.close()


def test_non_finite_positions_are_memory_safe(cuda_lib, orc, layout):
    """NaN, infinite or astronomically large query positions must not read out of bounds or poison later calls
    (the reference would throw an InexactError / BoundsError there; the device clamps the cell index)."""
    cset = ceg.GridCoordinatesSetup.from_cell(ceg.mat_from_parameters((7.0, 6.0, 5.0), (80.0, 95.0, 105.0)), 0.5)
    grid = _random_grid(cset, 3)
    eg = ceg.EnergyGrid(cset, (1, 1, 1), 1e-6, True, grid)
    bad = np.array([[np.nan, 0, 0], [np.inf, 1, 2], [-np.inf, np.inf, np.nan], [1e300, -1e300, 1e308], [1e19, 3.0, -2e18]])
    good = wl.random_positions(cset.cell.mat, 1000, 2)
    out = ceg.interpolate_grid(eg, np.concatenate([good[:500], bad, good[500:]]))
    assert out.shape == (1005,)
    want = orc.interpolate(cset, grid, False, good)
    got = np.concatenate([out[:500], out[505:]])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9)
    assert np.allclose(ceg.interpolate_grid(eg, good), want, rtol=1e-9, atol=1e-9)
    eg.close()


def test_small_batches_take_the_latency_path_with_the_same_bits(cuda_lib, orc, layout):
    """ceg_interpolate with at most 256 points is one launch of the latency kernel (csrc/latency.cuh) instead of the
    pipelined batch path: same stencil code, so the same bits — for every record layout, VdW (blocked corners included)
    and Coulomb grids, wrapped and far-away positions, non-finite positions."""
    cset = ceg.GridCoordinatesSetup.from_cell(ceg.mat_from_parameters((13.0187, 11.2063, 9.3758), (92.82, 107.2, 103.26)), 0.41)
    grid = _random_grid(cset, 9, hot_fraction=0.02)
    pos = np.concatenate([wl.random_positions(cset.cell.mat, 3000, 23), wl.random_positions(cset.cell.mat, 300, 24) * 7.0 - 30.0,
                          np.array([[np.nan, 0.0, 0.0], [np.inf, 1.0, 2.0]])])
    for is_vdw in (True, False):
        eg = ceg.EnergyGrid(cset, (1, 1, 1), np.inf if is_vdw else 1e-6, True, grid)
        batch = ceg._lib.interpolate(eg.handle, pos)
        if is_vdw:
            assert (batch == 1e100).any()
        for a, n in ((0, 1), (5, 2), (40, 31), (1000, 256), (len(pos) - 40, 40)):
            small = ceg._lib.interpolate(eg.handle, pos[a:a + n])
            assert np.array_equal(small, batch[a:a + n], equal_nan=True), (is_vdw, a, n)
        eg.close()
