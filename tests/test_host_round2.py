"""Round-2 host logic that needs no GPU: grid-cache naming of customised force fields (raspa.jl:196-204), the
energy_grid orientation contract (grids.jl:385-390), _set_gridpoint! on arrays (grids.jl:118-135) against the oracle,
the EnergyGrid host-copy hand-shake, the setup_montecarlo argument checks."""
import dataclasses
import os
import threading
import time

import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import forcefields as FF
from test_host_raspa import raspa_dir  # noqa: F401  (fixture)


def test_customised_forcefield_gets_its_own_grid_cache(raspa_dir):  # noqa: F811
    ff = ceg.parse_forcefield_RASPA("ToyFF")
    by_name, _ = ceg.grid_locations("Toy", "ToyFF", ff, ["O_co2"], 0.15, (2, 2, 3))
    as_object, vd = ceg.grid_locations("Toy", ff, ff, ["O_co2"], 0.15, (2, 2, 3))
    assert "/ToyFF/Toy/" in by_name
    assert f"/ToyFF-{ceg.hash_forcefield(ff)}/Toy/" in as_object and as_object != by_name
    assert vd[0].endswith("Toy_O_co2_truncated.grid")
    # a modified rule changes the hash (and so the cache directory); an identical copy does not
    inter = [list(row) for row in ff.interactions]
    i, j = ff.sdict["O_co2"], ff.sdict["Oz"]
    inter[i][j] = inter[j][i] = FF.LennardJones(99.0, 3.0)
    ff2 = ceg.ForceField(inter, ff.sdict, ff.symbols, ff.cutoff, ff.name)
    ff3 = ceg.ForceField([list(row) for row in ff.interactions], dict(ff.sdict), list(ff.symbols), ff.cutoff, ff.name)
    assert ceg.hash_forcefield(ff2) != ceg.hash_forcefield(ff) == ceg.hash_forcefield(ff3)
    tup, _ = ceg.grid_locations("Toy", ("ToyFF", ff2), ff2, ["O_co2"], 0.15, (2, 2, 3))
    assert f"/ToyFF-{ceg.hash_forcefield(ff2)}/" in tup


def test_energy_grid_refuses_to_drop_the_orientations():
    mol = ceg.RASPASystem(np.diag([np.inf] * 3), np.array([[0, 0, 1.1], [0, 0, 0], [0, 0, -1.1]]), ["O", "C", "O"],
                          np.ones(3), np.array([-0.3, 0.6, -0.3]), True)
    fw = ceg.RASPASystem(np.diag([10.0, 10.0, 10.0]), np.zeros((1, 3)), ["X"], np.ones(1), np.zeros(1), False)
    setup = ceg.CrystalEnergySetup(fw, mol, ceg.EnergyGrid.zero(), np.array([-0.3, 0.6, -0.3]), [ceg.EnergyGrid.zero()] * 2,
                                   [0, 1, 0], None, None, ceg.BlockFile(ceg.GridCoordinatesSetup.from_cell(fw.mat, 1.0)))
    with pytest.raises(ValueError, match="rotations"):
        ceg.energy_grid(setup, 1.0)            # num_rotate = 40, polyatomic, no rotation set: grids.jl:385-390 would use >= 40


def test_set_gridpoint_matches_the_oracle(orc):
    from ceg_b200.grids import set_gridpoint
    rng = np.random.default_rng(3)
    raw = rng.standard_normal((500, 8)) * 10.0 ** rng.integers(-3, 9, size=(500, 1))
    raw[::7, 0] = np.inf
    raw[3::11, 0] = np.nan
    delta = np.array([0.149, 0.151, 0.148])
    for lam, thr in (ceg.vdw_scaling(), ceg.coulomb_scaling()):
        got = set_gridpoint(raw, delta, lam, thr)
        want = np.stack([orc.set_gridpoint(delta, lam, thr, r) for r in raw])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_energy_grid_host_copy_waits_for_its_writer():
    g = ceg.EnergyGrid(None, (1, 1, 1), np.inf, True, np.zeros((0, 0, 0, 0), dtype=np.float32))
    done = []

    def fill():
        time.sleep(0.05)
        g._grid = np.ones((8, 2, 2, 2), dtype=np.float32)
        done.append(1)
    g._pending = threading.Thread(target=fill)
    g._pending.start()
    assert g.grid.shape == (8, 2, 2, 2) and done == [1] and g._pending is None


def test_setup_montecarlo_argument_checks(raspa_dir):  # noqa: F811
    with pytest.raises(ValueError, match="one blockfiles element per system"):
        ceg.setup_montecarlo("Toy", "ToyFF", ["Na"], blockfiles=[None, None])


def test_julia_shim_binds_only_declared_and_exported_symbols():
    """julia/CEGCuda.jl (the reference-side binding) may only `ccall` entry points that include/ceg_cuda.h
    declares and libceg_cuda.so exports; every method override INTEGRATION.md lists is present in the file."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    jl = open(os.path.join(root, "julia", "CEGCuda.jl"), encoding="utf-8").read()
    header = open(os.path.join(root, "include", "ceg_cuda.h"), encoding="utf-8").read()
    called = set(re.findall(r"ccall\(\(:(ceg_[a-z0-9_]+)", jl))
    assert len(called) >= 25
    declared = set(re.findall(r"\b(ceg_[a-z0-9_]+)\s*\(", header))
    assert called <= declared, called - declared
    assert called <= set(ceg._lib.EXPORTS), called - set(ceg._lib.EXPORTS)
    lib = ceg._lib.load()
    for name in called:
        assert hasattr(lib, name)
    for method in ("create_grid_vdw", "create_grid_coulomb", "interpolate_grid", "energy_point", "energy_grid", "compute_ewald"):
        assert re.search(r"function CrystalEnergyGrids\." + method + r"\(", jl), method
    assert {"ceg_ipc_export", "ceg_ipc_open", "ceg_ipc_close", "ceg_grid_build_records", "ceg_grid_create_empty"} <= called
    for fn in ("build!", "close_peers!"):
        assert ("function " + fn + "(") in jl, fn
    for fn in ("create_grids_fused", "create_grids_resident", "structure_factor", "single_contributions", "guest_energies",
               "block_from_spheres", "block_from_values", "handle_from_device", "interpolate_sweep"):
        assert re.search(r"function " + fn + r"\(", jl), fn


def test_committed_ncu_traffic_table_belongs_to_the_committed_kernels():
    """profiles/traffic.json (DRAM bytes per unit from `ncu --set full`, quoted by bench.py as roofline.traffic) carries
    the digest of the kernel sources it was captured on; bench.py drops the figure when the digest or the register count
    of the loaded library differs.  The committed table must belong to the committed sources."""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    table = json.load(open(os.path.join(root, "profiles", "traffic.json")))
    assert set(table) >= {"build", "interp", "interp_pair"}
    for key, row in table.items():
        assert row["csrc_digest"] == bench.csrc_digest(key), key
        assert row["dram_bytes_per_unit"] > 0 and row["registers"] > 0
        assert os.path.exists(os.path.join(root, row["source"].split(" ")[0])), row["source"]
    # the build kernel writes 64 B per grid point and reads next to nothing; the interpolation kernel moves ~400 B per query
    assert 64.0 <= table["build"]["dram_bytes_per_unit"] < 72.0
    assert 288.0 < table["interp"]["dram_bytes_per_unit"] < 450.0
