"""Host-side logic (no GPU): parsers, geometry, k-space, .grid I/O, C-ABI surface."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import workloads as wl

REF = "/root/reference/test/raspa"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference fixtures not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constants():
    assert ceg.GRID_TO_KELVIN == 1.202723550011625
    assert ceg.COULOMBIC_CONVERSION_FACTOR == 167100.9468073144
    lam, thr = ceg.coulomb_scaling()
    assert abs(lam - 138935.46) < 0.1 and abs(thr - 71.976) < 1e-3


def test_grid_coordinates_cha():
    """C1 218x204x190 nodes; dims always odd; abc_to_xyz order of operations."""
    inp, cset, _ = wl.load_fixture("cha_na")
    assert cset.nodes == (218, 204, 190) and cset.npoints == 8449680
    assert all(int(d) % 2 == 1 for d in cset.dims)
    c4 = wl.respace(cset, 0.1)
    assert c4.nodes == (326, 306, 284)
    p = ceg.abc_to_xyz(cset, 7, 9, 11)
    assert p[0] == 7 * cset.size[0] / cset.dims[0] + cset.shift[0]
    centre = cset.cell.mat @ np.array([0.5, 0.5, 0.5])
    s = ceg.offsetpoint(centre, cset)
    assert np.allclose(ceg.inverse_offsetpoint(s, cset), centre, atol=1e-9)
    # a point outside the cell is wrapped back before the affine map (coordinates.jl:58-66)
    assert np.allclose(ceg.offsetpoint(centre + cset.cell.mat[:, 0] - 2 * cset.cell.mat[:, 2], cset), s, atol=1e-9)


def test_supercell_and_min_image_setup():
    inp, cset, _ = wl.load_fixture("cit7_multi")
    assert ceg.find_supercell(cset.cell.mat, 12.0) == (2, 3, 3)
    ortho, safemin = ceg.prepare_periodic_distance_computations(inp.mat)
    assert not ortho and abs(safemin - 12.009) < 2e-3
    inp1, cs1, _ = wl.load_fixture("cha_na")
    ortho, safemin = ceg.prepare_periodic_distance_computations(inp1.mat)
    assert not ortho and abs(safemin - 14.08) < 1e-2
    # Float16 angle rule: 91.7 deg counts as orthogonal, 92.0 does not
    assert ceg.prepare_periodic_distance_computations(ceg.mat_from_parameters((10, 10, 10), (90, 90, 91.7)))[0]
    assert not ceg.prepare_periodic_distance_computations(ceg.mat_from_parameters((10, 10, 10), (90, 90, 92.0)))[0]


def test_replicate_order():
    pos = np.array([[0.1, 0.2, 0.3], [1.0, 1.1, 1.2]])
    mat = np.diag([2.0, 3.0, 4.0])
    rep = ceg.replicate_framework(pos, mat, (2, 1, 3))
    assert rep.shape == (12, 3)
    # index = ix*n*ny*nz + iy*n*nz + iz*n + i   (probes.jl:37-52)
    assert np.allclose(rep[1 * 2 * 3 + 0 + 2 * 2 + 1], pos[1] + [2.0, 0, 8.0])


def test_kspace_cha():
    inp, cset, ex = wl.load_fixture("cha_na")
    alpha, tol1 = ceg.ewald_parameters(1e-6)
    assert abs(alpha - 0.2650583) < 1e-7
    ks = ceg.build_kspace(inp.mat, alpha, tol1)
    assert ks.ks == (8, 8, 8) and ks.num_kvecs == 1368
    assert np.array_equal(ks.rows_array(), ex["krows"])
    kv = ks.kvec_array()
    assert (kv[:, 0] >= 0).all() and not ((kv == 0).all(axis=1)).any()
    kf = ceg.build_kfactors(ks, inp.mat, inp.invmat, alpha)
    assert np.array_equal(kf, ex["kfactors"])


def test_grid_file_roundtrip(tmp_path):
    """header 128 B (VdW) / 136 B (Coulomb), payload, 72 B cell-matrix tail;
    parse multiplies by GRID_TO_KELVIN in Float32."""
    cset = ceg.GridCoordinatesSetup.from_cell(ceg.mat_from_parameters((5.0, 6.0, 7.0), (80, 95, 100)), 0.9)
    nx, ny, nz = cset.nodes
    g = np.random.default_rng(1).standard_normal((8, nx, ny, nz)).astype(np.float32)
    pv, pc = tmp_path / "v.grid", tmp_path / "c.grid"
    ceg.write_grid(pv, cset, (2, 2, 2), g)
    ceg.write_grid(pc, cset, (2, 2, 2), g, ewald_precision=1e-6)
    assert os.path.getsize(pv) == 128 + g.nbytes + 72
    assert os.path.getsize(pc) == 136 + g.nbytes + 72
    ev = ceg.parse_grid(pv, False)
    ec = ceg.parse_grid(pc, True, cset.cell.mat)
    assert ev.num_unitcell == (2, 2, 2) and ev.ewald_precision == math.inf and ec.ewald_precision == 1e-6
    want = (g.astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32)
    assert np.array_equal(ev.grid, want) and np.array_equal(ec.grid, want)
    assert np.array_equal(ev.csetup.cell.mat, cset.cell.mat)
    with open(tmp_path / "short.grid", "wb") as f:
        f.write(open(pv, "rb").read()[:-72])
    with pytest.raises(ValueError):
        ceg.parse_grid(tmp_path / "short.grid", False)            # no mat given, none in the file
    assert ceg.parse_grid(tmp_path / "short.grid", False, cset.cell.mat).grid.shape == (8, nx, ny, nz)


def test_blockfile_from_values_oracle(orc):
    """grids.jl:188-204 (oracle restatement; the product computes it on the device)."""
    v = np.zeros((7, 7, 7), dtype=np.float32)
    v[2, 3, 1] = 6e6
    b = orc.blockfile_from_values(v)
    assert b.sum() == 8 and b[2:4, 3:5, 1:3].all()
    v[2, 3, 1] = 0
    v[6, 0, 0] = 6e6        # last node along x is not scanned (grids.jl:191)
    assert not orc.blockfile_from_values(v).any()


def test_pair_table_flattening():
    lj = ceg.forcefields.make_rule(ceg.Kind.LennardJones, [100.0, 3.0], True, 12.0)
    assert lj.shift == lj.__class__(ceg.Kind.LennardJones, (100.0, 3.0))(12.0)
    s = ceg.InteractionRuleSum.make([ceg.forcefields.Buckingham(1e7, 4.0, 1e5), ceg.forcefields.HardSphere(1.5, 0.0),
                                     ceg.forcefields.CoulombEwaldDirect(0.26, 1.0, -1.0)])
    assert [r.kind for r in s.rules] == [ceg.Kind.HardSphere, ceg.Kind.CoulombEwaldDirect, ceg.Kind.Buckingham]
    p = ceg.forcefields.flatten_pair(s)
    assert p["flags"] == 1 | 4 and p["hs_r"] == 1.5 and p["buck_B"] == 4.0
    assert ceg.forcefields.flatten_pair(ceg.forcefields.NoInteraction())["flags"] == 0
    with pytest.raises(ceg.UndefinedInteractionError):
        ceg.forcefields.flatten_pair(ceg.forcefields.Undefined())
    assert ceg.pair_dtype.itemsize == 72


@needs_ref
def test_forcefield_parser_on_reference_fixture():
    """SURVEY §8a A5': what the kernels see for the BoulfelfelSholl2021 force field."""
    ceg.setdir_RASPA(REF)
    ff = ceg.parse_forcefield_RASPA("BoulfelfelSholl2021")
    na_oz = ff["Na", "Oz"]
    kinds = [r.kind for r in na_oz.rules]
    assert kinds == [ceg.Kind.HardSphere, ceg.Kind.CoulombEwaldDirect, ceg.Kind.Buckingham]
    assert na_oz.rules[2].params == (5.581e7, 3.985, 9.167e5) and na_oz.rules[2].shift == 0.0   # truncated
    assert ff["Na", "Siz"].kind is ceg.Kind.CoulombEwaldDirect                                # `none` + charges
    ar_o = ff["Ar", "Oz"]
    assert ar_o.kind is ceg.Kind.LennardJones and ar_o.params == (107.69, 3.15) and ar_o.shift != 0.0  # shifted
    assert ff["Ar", "Siz"].kind is ceg.Kind.NoInteraction
    ch = ff["C_ch4", "H_ch4"]                                                                   # Jorgensen override
    lj = [r for r in ch.rules if r.kind is ceg.Kind.LennardJones][0]
    assert abs(lj.params[0] - math.sqrt(58.0 * 7.55)) < 1e-12 and abs(lj.params[1] - math.sqrt(3.825 * 2.4)) < 1e-12
    assert ceg.needsvdwgrid(ff, "Na") and ceg.needsvdwgrid(ff, "Ar")
    fw = ceg.load_framework_RASPA("CHA_1.4_3b4eeb96", "BoulfelfelSholl2021")
    assert len(fw) == 972
    inp = ceg.probe_inputs(fw, ff, ("Na",), True, ceg.ewald_parameters()[0])
    gold, _, _ = wl.load_fixture("cha_na")
    assert np.array_equal(inp.positions, gold.positions) and np.array_equal(inp.table.tobytes(), gold.table.tobytes())
    coul_path, vdws = ceg.grid_locations("CHA_1.4_3b4eeb96", "BoulfelfelSholl2021", ff, ["Na"], 0.15, (1, 1, 1))
    assert coul_path.endswith("grids/BoulfelfelSholl2021/CHA_1.4_3b4eeb96/0.150000/1x1x1/CHA_1.4_3b4eeb96_Electrostatics_Ewald.grid")
    assert vdws[0].endswith("0.150000/CHA_1.4_3b4eeb96_Na_truncated.grid")


def test_cabi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ceg_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ceg_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 24
    assert sorted(ceg._lib.EXPORTS) == declared
    lib = ctypes.CDLL(ceg.LIB_PATH)          # must load without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ceg_abi_version() == 2


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point fails loudly (no oracle, no numpy path)."""
    try:
        n = ceg._lib.device_count()
    except ceg.CegError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    inp, cset, _ = wl.load_fixture("cha_na")
    with pytest.raises(ceg.CegError):
        ceg.ProbeSystem(None, None, inputs=inp)
    with pytest.raises(ceg.CegError):
        ceg.interpolate_grid(ceg.EnergyGrid(cset, (1, 1, 1), math.inf, True, np.zeros((8,) + cset.nodes, np.float32)), np.zeros(3))
    src = "".join(open(os.path.join(ROOT, "crystalenergygrids.jl_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "crystalenergygrids.jl_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src and "libceg_oracle" not in src
