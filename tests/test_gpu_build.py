"""GPU parity of grid construction against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import workloads as wl
from util import assert_close_f64, assert_f32_grids_match

pytestmark = pytest.mark.gpu


def _probe(inp):
    return ceg.ProbeSystem(None, None, inputs=inp)


def _rand_positions(mat, n, seed):
    return wl.random_positions(mat, n, seed)


@pytest.mark.parametrize("fixture", ["cha_na", "chana_ar", "cit7_multi"])
def test_literal_kernel_vs_oracle(cuda_lib, orc, fixture):
    """ceg_compute_derivatives == probes.jl:71-117 restated, at random positions."""
    inp, cset, _ = wl.load_fixture(fixture)
    pos = _rand_positions(cset.cell.mat, 300, 11)
    # also a few grid nodes and points very close to atoms (hard-sphere / 1 Å / Inf paths)
    near = inp.positions[:20] + np.array([0.3, -0.2, 0.1])
    exact = inp.positions[:3].copy()   # r = 0: Inf / NaN paths
    pos = np.concatenate([pos, near, exact])
    ps = _probe(inp)
    got = ps.compute_derivatives(pos)
    want = orc.compute_derivatives(inp, pos)
    ng = got.shape[0]
    for g in range(ng):
        coul = inp.want_coulomb and g == ng - 1
        assert_close_f64(got[g], want[g], coul, f"{fixture} grid {g}")
    ps.close()


def _check_slab(orc, inp, cset, i0, i1, what):
    nx, ny, nz = cset.nodes
    ng = inp.table.shape[0] + int(inp.want_coulomb)
    ps = _probe(inp)
    gs = ceg._lib.make_gridspec(cset)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    outs = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(ng)]
    raw = np.empty((ng, 8, i1 - i0, ny, nz))
    ceg._lib.grid_build(ps.handle, gs, i0, i1, lam_v, thr_v, lam_c, thr_c, outs, raw)
    stats = ceg._lib.probe_last_stats(ps.handle)
    ps.close()
    o_outs, o_raw = orc.grid_build(inp, cset, i0, i1, 1, want_f32=True, want_raw=True)
    for g in range(ng):
        coul = inp.want_coulomb and g == ng - 1
        assert_close_f64(raw[g], o_raw[g], coul, f"{what} raw grid {g}")
        lam, thr = (lam_c, thr_c) if coul else (lam_v, thr_v)
        assert_f32_grids_match(outs[g][:, i0:i1], o_outs[g][:, i0:i1], o_raw[g], lam, thr, f"{what} f32 grid {g}")
        # planes outside the slab are untouched
        assert not outs[g][:, :i0].any() and not outs[g][:, i1:].any()
    assert stats["n_points"] == (i1 - i0) * ny * nz
    assert 0 < stats["pairs_in_cutoff"] <= stats["pairs_tested"]
    return stats


@pytest.mark.parametrize("i0,i1", [(0, 2), (107, 110), (216, 218)])
def test_tiled_slab_cha_na(cuda_lib, orc, i0, i1):
    """C1 (CHA, Na+: Buckingham+HardSphere VdW fused with the Ewald real-space grid), x-slabs
    including the first/last planes and a ragged final tile."""
    inp, cset, _ = wl.load_fixture("cha_na")
    _check_slab(orc, inp, cset, i0, i1, f"cha_na[{i0}:{i1}]")


def test_tiled_slab_chana_ar_shifted_lj(cuda_lib, orc):
    inp, cset, _ = wl.load_fixture("chana_ar")
    _check_slab(orc, inp, cset, 50, 52, "chana_ar")


def test_tiled_slab_cit7_triclinic_multi_probe(cuda_lib, orc):
    """Triclinic cell, supercell (2,3,3), three fused VdW probes + Coulomb."""
    inp, cset, _ = wl.load_fixture("cit7_multi")
    assert inp.supercell == (2, 3, 3)
    cs = wl.respace(cset, 0.25)
    nx = cs.nodes[0]
    _check_slab(orc, inp, cs, 0, 3, "cit7[0:3]")
    _check_slab(orc, inp, cs, nx - 5, nx, "cit7[last]")


def test_tiled_slab_mfi_synthetic(cuda_lib, orc):
    """C3: orthorhombic synthetic framework, CH4 + CO2 sites (LJ) + Coulomb, supercell (2,2,2)."""
    inp, cset, fw = wl.synthetic_inputs("mfi", 0.3, ("C_ch4", "C_co2", "O_co2"), True)
    assert inp.ortho and inp.supercell == (2, 2, 2)
    _check_slab(orc, inp, cset, 10, 14, "mfi")


def test_tiled_slab_mof_synthetic(cuda_lib, orc):
    """C5: 50 Å cubic cell, 5000 atoms of four species, supercell (1,1,1); LJ probe + Coulomb."""
    inp, cset, fw = wl.synthetic_inputs("mof", 0.8, ("C_ch4",), True)
    assert len(inp.positions) >= 4900 and inp.supercell == (1, 1, 1)
    _check_slab(orc, inp, cset, 30, 33, "mof")


def test_tiled_slab_rhombohedral_image_search(cuda_lib, orc):
    """safemin < cutoff: pairs between the two radii go through the reference's +-1 image search
    (utils.jl:226-246), whose exits leave either the first smaller image or a stale displacement vector in
    the buffer the derivative tensors are built from.  Tiled kernel and literal kernel vs the oracle."""
    inp, cset, fw = wl.synthetic_inputs("rhombo60", 0.4, ("C_ch4", "Na"), True)
    assert not inp.ortho and inp.supercell == (3, 3, 3) and 100.0 < inp.safemin2 < inp.cutoff2
    nx = cset.nodes[0]
    _check_slab(orc, inp, cset, 0, 5, "rhombo60[0:5]")
    _check_slab(orc, inp, cset, nx - 3, nx, "rhombo60[last]")
    pos = _rand_positions(cset.cell.mat, 400, 3)
    ps = _probe(inp)
    got = ps.compute_derivatives(pos)
    ps.close()
    want = orc.compute_derivatives(inp, pos)
    for g in range(got.shape[0]):
        assert_close_f64(got[g], want[g], g == got.shape[0] - 1, f"rhombo60 literal grid {g}")


def test_ragged_slab_uses_every_tile_shape(cuda_lib, orc):
    """13 planes = one 8-plane row + the 4-, and 1-plane tail shapes; 7 planes = 4 + 2 + 1."""
    inp, cset, _ = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 0.45)
    _check_slab(orc, inp, cs, 5, 18, "ragged 13")
    _check_slab(orc, inp, cs, 40, 47, "ragged 7")


def test_single_plane_and_vdw_only_and_coulomb_only(cuda_lib, orc):
    inp, cset, _ = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 0.6)
    import dataclasses
    vdw_only = dataclasses.replace(inp, want_coulomb=False, charges=None)
    _check_slab(orc, vdw_only, cs, 7, 8, "vdw-only single plane")
    coul_only = dataclasses.replace(inp, table=inp.table[:0])
    _check_slab(orc, coul_only, cs, 30, 33, "coulomb-only")


def test_full_grid_properties_and_slab_split(cuda_lib, orc):
    """Full-size C1 build (8.45 M points x 2 grids): slab-split invariance (bit-exact), clamp
    semantics, and parity on sampled x planes against the oracle."""
    inp, cset, _ = wl.load_fixture("cha_na")
    nx, ny, nz = cset.nodes
    assert (nx, ny, nz) == (218, 204, 190)
    ps = _probe(inp)
    gs = ceg._lib.make_gridspec(cset)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    full = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, full, None)
    st = ceg._lib.probe_last_stats(ps.handle)
    assert st["n_points"] == nx * ny * nz
    # in-cutoff pairs per point: N_sc * (4/3) pi 12^3 / V  (~310)
    per_point = st["pairs_in_cutoff"] / st["n_points"]
    assert 300 < per_point < 320, per_point
    split = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    for a, b in [(0, 77), (77, 150), (150, nx)]:
        ceg._lib.grid_build(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, split, None)
    ps.close()
    for g in range(2):
        assert np.array_equal(full[g].view(np.uint32), split[g].view(np.uint32)), "slab split changes bits"
        lam, thr = (lam_c, thr_c) if g == 1 else (lam_v, thr_v)
        clamp = full[g][0] == np.float32(2 * thr * lam)
        assert clamp.any()
        assert not full[g][4:, clamp].any(), "second/third derivatives must be zero where clamped"
        assert (np.abs(full[g][1:4, clamp]) <= np.float32(thr * lam) * (1 + 1e-6)).all()
        assert np.isfinite(full[g][0]).all()
    planes = [3, 64, 131, 217]
    for i in planes:
        o_outs, _ = orc.grid_build(inp, cset, i, i + 1, 1, want_f32=True, want_raw=False)
        for g in range(2):
            lam, thr = (lam_c, thr_c) if g == 1 else (lam_v, thr_v)
            assert_f32_grids_match(full[g][:, i:i + 1], o_outs[g][:, i:i + 1], None, lam, thr, f"full plane {i} grid {g}")


def test_error_paths(cuda_lib):
    inp, cset, _ = wl.load_fixture("cha_na")
    ps = _probe(inp)
    gs = ceg._lib.make_gridspec(cset)
    with pytest.raises(ceg.CegError):
        ceg._lib.grid_build(ps.handle, gs, 5, 5, 1, 1, 1, 1, [None, None], None)      # empty slab
    with pytest.raises(ceg.CegError):
        ceg._lib.grid_build(ps.handle, gs, 0, 10 ** 6, 1, 1, 1, 1, [None, None], None)  # out of range
    ps.close()
    import dataclasses
    bad = dataclasses.replace(inp, cutoff2=20.0 ** 2)   # cutoff beyond half the supercell width
    ps = _probe(bad)
    with pytest.raises(ceg.CegError):
        ceg._lib.grid_build(ps.handle, gs, 0, 1, 1, 1, 1, 1, [None, None], None)
    ps.close()


def test_registered_host_output_and_cache_trim(cuda_lib):
    """ceg_host_register pins a caller-owned output array (what a multi-process driver does with its slab runs);
    the build into it gives the same bytes; ceg_trim_device_cache releases the pooled device scratch and the next
    call simply re-allocates."""
    inp, cset, _ = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 0.6)
    nx, ny, nz = cs.nodes
    gs = ceg._lib.make_gridspec(cs)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    ps = _probe(inp)
    plain = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, plain, None)
    pinned = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    for p in pinned:
        for c in range(8):
            ceg._lib.host_register(p[c, 3:nx - 2])          # the 8 runs of an x-slab
    n0 = ceg._lib.kernel_launch_count()
    ceg._lib.grid_build(ps.handle, gs, 3, nx - 2, lam_v, thr_v, lam_c, thr_c, pinned, None)
    assert ceg._lib.kernel_launch_count() > n0
    for p in pinned:
        for c in range(8):
            ceg._lib.host_unregister(p[c, 3:nx - 2])
    for g in range(2):
        assert np.array_equal(pinned[g][:, 3:nx - 2], plain[g][:, 3:nx - 2]) and not pinned[g][:, :3].any()
    ceg._lib.trim_device_cache()
    again = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, again, None)
    assert np.array_equal(again[0], plain[0]) and np.array_equal(again[1], plain[1])
    ps.close()
    info = ceg._lib.device_info(0)
    assert info["sm_count"] > 0 and info["mem_bytes"] > 0


def test_thick_slab_of_a_large_cell_uses_x_windows(cuda_lib):
    """A 50 Å cell with 5000 atoms at 0.25 Å: the image list of the whole 201-plane grid (> 12k images) is replaced
    by per-window lists (ceg_cuda.cu: list_windows).  Same bytes as the planes built as separate thin slabs."""
    inp, cset, fw = wl.synthetic_inputs("mof", 0.25, ("C_ch4",), True)
    nx, ny, nz = cset.nodes
    assert nx > 128
    gs = ceg._lib.make_gridspec(cset)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    ps = _probe(inp)
    full = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, full, None)
    st = ceg._lib.probe_last_stats(ps.handle)
    assert st["n_points"] == nx * ny * nz
    for (a, b) in [(0, 5), (60, 70), (126, 133), (nx - 3, nx)]:
        thin = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
        ceg._lib.grid_build(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, thin, None)
        for g in range(2):
            assert np.array_equal(thin[g][:, a:b], full[g][:, a:b]), (g, a, b)
    ps.close()


def test_cutoff_beyond_the_probe_supercell_uses_the_literal_kernel(cuda_lib, orc):
    """A force-field cutoff above half the width of the 12 Å probe supercell (probes.jl:24): the reference still builds the
    grid through its minimum-image search; build_grids routes such slabs through ceg_compute_derivatives + _set_gridpoint!."""
    from util import assert_close_f64
    fw, _, species = wl.mfi_shaped()
    ff14 = wl._lj_forcefield(species, cutoff=14.0)
    inp = ceg.probe_inputs(fw, ff14, ("C_ch4",))
    assert 2 * 14.0 > min(ceg.perpendicular_lengths(inp.mat))
    cset, grids, raw = ceg.build_grids(fw, ff14, 1.2, atoms=("C_ch4",), xrange=(3, 5), want_raw=True)
    o_outs, o_raw = orc.grid_build(inp, cset, 3, 5, 1, want_f32=True, want_raw=True)
    assert_close_f64(raw[0], o_raw[0], False, "cutoff 14 raw")
    lam, thr = ceg.vdw_scaling()
    assert_f32_grids_match(grids[0][:, 3:5], o_outs[0][:, 3:5], o_raw[0], lam, thr, "cutoff 14 f32")
