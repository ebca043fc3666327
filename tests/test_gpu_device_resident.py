"""Round-2 device-resident paths: fused (VdW, Coulomb) pair records, build -> gather -> upload hand-off
(build_grids_device, ceg_grid_gather), setup_RASPA / setup_montecarlo without the file round trip
(raspa.jl:420-439, montecarlo.jl:269-334), in-process multi-GPU builds."""
import os

import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import workloads as wl
from test_gpu_setup_raspa import _make_raspa_dir

pytestmark = pytest.mark.gpu


def _two_grids(spacing=0.5):
    inp, cset, ex = wl.load_fixture("cha_na")
    cs = wl.respace(cset, spacing)
    nx, ny, nz = cs.nodes
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    outs = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cs), 0, nx, lam_v, thr_v, lam_c, thr_c, outs, None)
    ps.close()
    gK = [(g.astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32) for g in outs]
    return inp, cs, ex, outs, gK


@pytest.mark.parametrize("layouts", [(0, 0), (4, 4), (0, 4)], ids=["node", "block", "mixed"])
def test_pair_records_equal_separate_lookups(cuda_lib, layouts):
    """ceg_grid_pair_create / ceg_interpolate_pair: bit-identical to ceg_interpolate on each grid, from either
    source layout; blocked positions (VdW corner above 5e6 K) keep their Coulomb value."""
    inp, cs, ex, outs, gK = _two_grids()
    ceg._lib.set_interp_layout(layouts[0])
    eg_v = ceg.EnergyGrid(cs, (1, 1, 1), np.inf, True, gK[0])
    eg_v.handle
    ceg._lib.set_interp_layout(layouts[1])
    eg_c = ceg.EnergyGrid(cs, (1, 1, 1), 1e-6, True, gK[1])
    eg_c.handle
    ceg._lib.set_interp_layout(-1)
    pos = np.concatenate([wl.random_positions(cs.cell.mat, 60000, 5), wl.random_positions(cs.cell.mat, 3000, 6) * 5.0 - 40.0,
                          np.zeros((1, 3))])
    v, c = ceg.interpolate_pair(eg_v, eg_c, pos)
    assert np.array_equal(v, ceg.interpolate_grid(eg_v, pos))
    assert np.array_equal(c, ceg.interpolate_grid(eg_c, pos))
    assert 0.05 < (v == 1e100).mean() < 0.6
    # the fused records are what energy_points uses once the pair exists — same bits as without it
    fw = ceg.RASPASystem(ex["unit_mat"], ex["positions"][:len(ex["site_charge"])], ["X"] * len(ex["site_charge"]),
                         np.ones(len(ex["site_charge"])), ex["charges"][:len(ex["site_charge"])], False)
    q = float(ex["probe_charge"])
    na = ceg.RASPASystem(np.diag([np.inf] * 3), np.zeros((1, 3)), ["Na"], np.ones(1), np.array([q]), True)
    ef = ceg.initialize_ewald(fw)
    eg_v.unpair()
    setup = ceg.CrystalEnergySetup(fw, na, eg_c, np.array([q]), [eg_v], [0], ef, None, ceg.BlockFile(cs))
    v0, c0 = ceg.energy_points(setup, pos[:20000, None, :])
    setup.fused_pairs = True
    v1, c1 = ceg.energy_points(setup, pos[:20000, None, :])
    assert eg_v._pair is not None
    assert np.array_equal(v0, v1) and np.array_equal(c0, c1)
    eg_v.close(); eg_c.close()


def test_pair_create_rejects_mismatched_grids(cuda_lib):
    inp, cs, ex, outs, gK = _two_grids(0.8)
    cs2 = wl.respace(cs, 0.9)
    a = ceg.EnergyGrid(cs, (1, 1, 1), np.inf, True, gK[0])
    b = ceg.EnergyGrid(cs2, (1, 1, 1), 1e-6, True, np.zeros((8, *cs2.nodes), dtype=np.float32))
    with pytest.raises(ceg.CegError):
        ceg._lib.grid_pair_create(a.handle, b.handle)
    with pytest.raises(ceg.CegError):
        ceg._lib.grid_pair_create(b.handle, a.handle)      # (Coulomb, VdW) order
    a.close(); b.close()


def test_gather_and_device_build_equal_host_build(cuda_lib):
    """ceg_grid_gather (here: two parts on one device) + ceg_grid_upload_dev == build into host arrays -> parse_grid
    scaling -> upload; build_grids_device returns the same grids with their `.grid`-unit host copies."""
    inp, cs, ex, outs, gK = _two_grids(0.5)
    nx, ny, nz = cs.nodes
    gs = ceg._lib.make_gridspec(cs)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    cut = nx // 3
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    parts = []
    for a, b in ((0, cut), (cut, nx)):
        d = [ceg._lib.device_malloc(0, 8 * (b - a) * ny * nz * 4) for _ in range(2)]
        ceg._lib.grid_build_dev(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, d, 0)
        parts.append((a, b, d))
    ceg._lib.device_synchronize(0)
    ps.close()
    pos = wl.random_positions(cs.cell.mat, 30000, 8)
    for g, prec in enumerate((np.inf, 1e-6)):
        d_full = ceg._lib.device_malloc(0, 8 * nx * ny * nz * 4)
        ceg._lib.grid_gather([0, 0], [p[2][g] for p in parts], [p[0] for p in parts], [p[1] for p in parts], gs, 0, d_full)
        back = np.empty((8, nx, ny, nz), dtype=np.float32)
        ceg._lib.device_download(0, d_full, back)
        assert np.array_equal(back, outs[g])
        on_dev = ceg.EnergyGrid.from_device(cs, (1, 1, 1), prec, d_full)
        ceg._lib.device_free(d_full)
        ref = ceg.EnergyGrid(cs, (1, 1, 1), prec, True, gK[g])
        assert np.array_equal(ceg.interpolate_grid(on_dev, pos), ceg.interpolate_grid(ref, pos))
        on_dev.close(); ref.close()
    # the same straight from the parts (ceg_grid_upload_parts_dev), exact and padded strides: no assembled copy
    import torch
    for g, prec in enumerate((np.inf, 1e-6)):
        ref = ceg.EnergyGrid(cs, (1, 1, 1), prec, True, gK[g])
        want = ceg.interpolate_grid(ref, pos)
        eg = ceg.EnergyGrid.from_device_parts(cs, (1, 1, 1), prec, [p[2][g] for p in parts], [(p[0], p[1]) for p in parts])
        assert np.array_equal(ceg.interpolate_grid(eg, pos), want)
        eg.close()
        pad = []
        for a, b, _ in parts:
            t = torch.full((8, b - a + 3, ny, nz), float("nan"), dtype=torch.float32, device="cuda")
            t[:, :b - a] = torch.from_numpy(outs[g][:, a:b]).cuda()
            pad.append(t)
        eg = ceg.EnergyGrid.from_device_parts(cs, (1, 1, 1), prec, [t.data_ptr() for t in pad], [(p[0], p[1]) for p in parts],
                                              [t.shape[1] for t in pad])
        assert np.array_equal(ceg.interpolate_grid(eg, pos), want)
        eg.close()
        # many narrow parts (8 ranks x 4 sub-slabs in bench.py): up to 64, the 65th is refused
        edges = np.unique(np.linspace(0, nx, min(nx, 64) + 1).astype(int))
        many = [torch.from_numpy(np.ascontiguousarray(outs[g][:, a:b])).cuda() for a, b in zip(edges[:-1], edges[1:])]
        eg = ceg.EnergyGrid.from_device_parts(cs, (1, 1, 1), prec, [t.data_ptr() for t in many],
                                              list(zip(edges[:-1].tolist(), edges[1:].tolist())))
        assert np.array_equal(ceg.interpolate_grid(eg, pos), want)
        eg.close(); ref.close()
    with pytest.raises(ceg.CegError):
        ceg.EnergyGrid.from_device_parts(cs, (1, 1, 1), np.inf, [parts[0][2][0]] * 65, [(i, i + 1) for i in range(65)])
    with pytest.raises(ceg.CegError):       # parts must tile the x range
        ceg.EnergyGrid.from_device_parts(cs, (1, 1, 1), np.inf, [parts[1][2][0]], [(parts[1][0], parts[1][1])])
    for p in parts:
        for d in p[2]:
            ceg._lib.device_free(d)
    with pytest.raises(ceg.CegError):
        ceg._lib.device_free(12345)


def test_setup_raspa_device_handoff_equals_file_round_trip(cuda_lib, tmp_path):
    """§8(f)4: the grids setup_RASPA creates go from the build kernels to the interpolator on the device; the files it
    writes in the background and the energies it returns are identical to the reference's write-then-parse path."""
    _make_raspa_dir(tmp_path)
    ceg.setdir_RASPA(str(tmp_path))
    a = ceg.setup_RASPA("Toy", "ToyFF", "CO2", "TraPPE", gridstep=0.4, blockfile=False, grids_root=str(tmp_path / "ga"))
    b = ceg.setup_RASPA("Toy", "ToyFF", "CO2", "TraPPE", gridstep=0.4, blockfile=False, grids_root=str(tmp_path / "gb"),
                        keep_on_device=False)
    assert a.grids[0]._handle is not None and b.grids[0]._handle is None      # resident before any host copy exists
    mol = np.asarray(a.molecule.position)
    pos = wl.random_positions(a.framework.mat, 5000, 4)[:, None, :] + mol[None, :, :]
    va, ca = ceg.energy_points(a, pos)
    vb, cb = ceg.energy_points(b, pos)
    assert np.array_equal(va, vb) and np.array_equal(ca, cb)
    a.wait()
    fa = sorted(p.relative_to(tmp_path / "ga") for p in (tmp_path / "ga").rglob("*.grid"))
    fb = sorted(p.relative_to(tmp_path / "gb") for p in (tmp_path / "gb").rglob("*.grid"))
    assert fa == fb and len(fa) == 3
    for f in fa:
        assert (tmp_path / "ga" / f).read_bytes() == (tmp_path / "gb" / f).read_bytes()
    for ga, gb in zip(a.grids + [a.coulomb], b.grids + [b.coulomb]):
        assert np.array_equal(ga.grid, gb.grid) and ga.num_unitcell == gb.num_unitcell
        assert ga.ewald_precision == gb.ewald_precision
    # a customised force field with the stock name must not share the stock cache (raspa.jl:196-204)
    ff = ceg.parse_forcefield_RASPA("ToyFF")
    cpath, _ = ceg.grid_locations("Toy", ff, ff, ["O_co2"], 0.4, (2, 2, 3), str(tmp_path / "ga"))
    assert f"ToyFF-{ceg.hash_forcefield(ff)}" in cpath


def test_setup_montecarlo_grids_and_framework_interactions(cuda_lib, tmp_path):
    """montecarlo.jl:269-334 (grid part) and :490-520: two molecule kinds share the Coulomb grid and the O_co2 / C_co2 /
    Na VdW grids; framework_interactions == energy_points without the reciprocal term."""
    _make_raspa_dir(tmp_path)
    ceg.setdir_RASPA(str(tmp_path))
    co2 = ceg.load_molecule_RASPA("CO2", "TraPPE", "ToyFF")
    mc = ceg.setup_montecarlo("Toy.cif", "ToyFF", [co2, ("Na", 3)], gridstep=0.4, blockfiles=[False, False])
    ff = mc.forcefield
    assert [g is not None for g in mc.grids] == [s in ("C_co2", "O_co2", "Na") for s in ff.symbols]
    assert mc.num_framework_atoms[ff.sdict["Oz"]] == 32 * 12 and mc.num_framework_atoms[ff.sdict["Siz"]] == 16 * 12
    assert mc.coulomb.ewald_precision == 1e-6 and mc.cell.shape == (3, 3)
    ref = ceg.setup_RASPA("Toy", "ToyFF", "CO2", "TraPPE", gridstep=0.4, blockfile=False)
    mol = np.asarray(co2.position)
    pos = wl.random_positions(ref.framework.mat, 3000, 12)[:, None, :] + mol[None, :, :]
    v, d = ceg.framework_interactions(mc, 0, pos)
    v2, d2 = ceg.energy_points(ref, pos, reciprocal=False)
    assert np.array_equal(v, v2) and np.array_equal(d, d2)
    v1, d1 = ceg.framework_interactions(mc, 0, pos[7])
    assert v1 == v[7] and d1 == d[7]
    vn, dn = ceg.framework_interactions(mc, 1, pos[:100, :1])
    assert vn.shape == (100,) and np.isfinite(dn).all()


def test_many_framework_species(cuda_lib, orc):
    """More framework species than round 1's 24-entry shared-memory table (the reference has no limit)."""
    import dataclasses
    inp, cset, _ = wl.synthetic_inputs("mfi", 0.6, ("C_ch4",), True)
    n = len(inp.positions)
    n_types = 40
    types = (np.arange(n) % n_types).astype(np.int32)
    table = np.zeros((1, n_types), dtype=inp.table.dtype)
    for t in range(n_types):
        table[0, t] = inp.table[0, inp.types[0]]
        table[0, t]["lj_eps"] *= 1.0 + 0.01 * t
    inp2 = dataclasses.replace(inp, types=types, n_types=n_types, table=table)
    nx, ny, nz = cset.nodes
    ps = ceg.ProbeSystem(None, None, inputs=inp2)
    raw = np.empty((2, 8, 2, ny, nz))
    outs = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cset), 10, 12, *ceg.vdw_scaling(), *ceg.coulomb_scaling(), outs, raw)
    ps.close()
    _, o_raw = orc.grid_build(inp2, cset, 10, 12, 1, want_f32=False, want_raw=True)
    from util import assert_close_f64
    assert_close_f64(raw[0], o_raw[0], False, "40 species vdw")
    assert_close_f64(raw[1], o_raw[1], True, "40 species coulomb")


def test_error_budget_of_the_summed_channels(cuda_lib, orc):
    """The hot-path math is shorter than round 1's (erfcx degree 11 on [0.25, 3.25], the Ewald Gaussian from
    a product of three tabulated factors): on SUMS of ~310 partly cancelling pair terms per channel of the C1 grids the
    worst error stays below a third of the north-star tolerance (1e-9 relative + 1e-6 K per grid point)."""
    from util import ATOL_K, RTOL, atol_for
    inp, cset, _ = wl.load_fixture("cha_na")
    nx, ny, nz = cset.nodes
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    raw = np.empty((2, 8, 3, ny, nz))
    ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cset), 60, 63, *ceg.vdw_scaling(), *ceg.coulomb_scaling(),
                        [None, None], raw)
    ps.close()
    _, o_raw = orc.grid_build(inp, cset, 60, 63, 1, want_f32=False, want_raw=True)
    worst = []
    for g, coul in ((0, False), (1, True)):
        fin = np.isfinite(o_raw[g])
        err = np.abs(raw[g][fin] - o_raw[g][fin])
        tol = RTOL * np.abs(o_raw[g][fin]) + atol_for(coul)
        worst.append(float((err / tol).max()))
    print("worst error / tolerance (VdW, Coulomb):", worst)
    assert max(worst) < 1.0 / 3.0, f"worst error / tolerance = {worst}"


def _n_devices():
    try:
        return ceg._lib.device_count()
    except Exception:  # noqa: BLE001  (no driver / no device: the CPU-only collection run)
        return 0


@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs")
def test_two_gpus_in_one_process(cuda_lib):
    """build_grids(devices=[0, 1]) (the analogue of @threads, grids.jl:144,171) and the device-resident
    build -> peer gather -> upload on both GPUs equal the single-GPU results bit for bit."""
    fwk, ffk, _ = wl.mfi_shaped()
    from ceg_b200.ewald import ewald_parameters
    alpha = ewald_parameters(1e-6)[0]
    cs1, g1, _ = ceg.build_grids(fwk, ffk, 0.4, atoms=("C_ch4",), coulomb=True, alpha=alpha, devices=[0])
    cs2, g2, _ = ceg.build_grids(fwk, ffk, 0.4, atoms=("C_ch4",), coulomb=True, alpha=alpha, devices=[0, 1])
    assert all(np.array_equal(a, b) for a, b in zip(g1, g2))
    cs3, egs, host = ceg.build_grids_device(fwk, ffk, 0.4, atoms=("C_ch4",), coulomb=True, alpha=alpha, devices=[0, 1],
                                            download=True)
    assert all(np.array_equal(a, b) for a, b in zip(g1, host))
    pos = wl.random_positions(fwk.mat, 20000, 2)
    for g, eg in enumerate(egs):
        assert set(eg._replicas) == {1}
        ref = ceg.EnergyGrid(cs1, eg.num_unitcell, eg.ewald_precision, True,
                             (g1[g].astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32))
        want = ceg.interpolate_grid(ref, pos)
        assert np.array_equal(ceg._lib.interpolate(eg.handle_on(0), pos), want)
        assert np.array_equal(ceg._lib.interpolate(eg.handle_on(1), pos), want)
        ref.close(); eg.close()


@pytest.mark.parametrize("layout", [0, 4], ids=["node", "block"])
def test_build_records_equals_build_file_parse_upload(cuda_lib, layout):
    """ceg_grid_build_records: the build kernel writes finished interpolation records (parse_grid's conversion and the
    device layout applied in its epilogue) into several destination grids at once — bit-identical to build -> `.grid`
    values -> parse_grid scaling -> ceg_grid_upload, for slabs of any parity, with the `.grid`-order slab on the side."""
    inp, cs, ex, outs, gK = _two_grids(0.5)
    nx, ny, nz = cs.nodes
    gs = ceg._lib.make_gridspec(cs)
    uc = ceg.grids.unitcell_struct(cs.cell)
    lam = ceg.vdw_scaling() + ceg.coulomb_scaling()
    pos = wl.random_positions(cs.cell.mat, 30000, 11)
    ceg._lib.set_interp_layout(layout)
    try:
        dests = [[ceg._lib.grid_create_empty(0, uc, gs, g == 0) for g in range(2)] for _ in range(3)]
        recs = [[ceg._lib.grid_records(h) for h in hs] for hs in dests]
        assert all(r[2] == layout and r[1] >= 32 * nx * ny * nz for rs in recs for r in rs)
        ps = ceg.ProbeSystem(None, None, inputs=inp)
        cut = nx // 3 | 1                                   # an odd first plane for the second slab: bricks straddle it
        for a, b in ((0, cut), (cut, nx)):
            d_out = [ceg._lib.device_malloc(0, 8 * (b - a) * ny * nz * 4) for _ in range(2)]
            ceg._lib.grid_build_records(ps.handle, gs, a, b, *lam, [[r[0] for r in rs] for rs in recs], layout,
                                        [ceg.GRID_TO_KELVIN] * 2, d_out, 0)
            ceg._lib.device_synchronize(0)
            for g in range(2):
                back = np.empty((8, b - a, ny, nz), dtype=np.float32)
                ceg._lib.device_download(0, d_out[g], back)
                assert np.array_equal(back.view(np.uint32), outs[g][:, a:b].view(np.uint32))
                ceg._lib.device_free(d_out[g])
        for g, prec in enumerate((np.inf, 1e-6)):
            ref = ceg.EnergyGrid(cs, (1, 1, 1), prec, True, gK[g])
            want = ceg.interpolate_grid(ref, pos)
            for hs in dests:
                assert np.array_equal(ceg._lib.interpolate(hs[g], pos), want)
            ref.close()
        # argument checks
        good = [[r[0] for r in recs[0]]]
        with pytest.raises(ceg.CegError):
            ceg._lib.grid_build_records(ps.handle, gs, 0, nx, *lam, good * 9, layout, [1.0, 1.0], None, 0)
        with pytest.raises(ceg.CegError):
            ceg._lib.grid_build_records(ps.handle, gs, 0, nx, *lam, good, 1, [1.0, 1.0], None, 0)       # QUAD layout
        with pytest.raises(ceg.CegError):
            ceg._lib.grid_build_records(ps.handle, gs, 0, nx, *lam, [[good[0][0] + 4, good[0][1]]], layout, [1.0, 1.0], None, 0)
        with pytest.raises(ceg.CegError):
            ceg._lib.grid_build_records(ps.handle, gs, 0, nx, *lam, good, layout, [1.0, 0.0], None, 0)
        ps.close()
        for hs in dests:
            for h in hs:
                ceg._lib.grid_free(h)
    finally:
        ceg._lib.set_interp_layout(-1)


def _sharded_worker(rank, world, port, outdir):
    import sys
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    dev = rank % torch.cuda.device_count()               # one GPU: both ranks share it (IPC works within a device too)

    def exchange(o):
        got = [None] * world
        dist.all_gather_object(got, o)
        return got

    inp, cset, _ = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 0.5)
    nx = cs.nodes[0]
    sb = ceg.ShardedGridBuild(inp, cs, (1, 1, 1), (np.inf, 1e-6), rank, world, exchange, device=dev)
    sb.build()
    ceg._lib.device_synchronize(dev)
    dist.barrier()                                        # every rank's slab has landed in every replica
    # reference: the whole grid built by this process alone, through host arrays
    ps = ceg.ProbeSystem(None, None, inputs=inp, device=dev)
    outs = [np.zeros((8,) + tuple(cs.nodes), dtype=np.float32) for _ in range(2)]
    ceg._lib.grid_build(ps.handle, sb.gs, 0, nx, *sb.lam, outs, None)
    ps.close()
    pos = wl.random_positions(cs.cell.mat, 20000, 5 + rank)
    for g, eg in enumerate(sb.grids):
        ref = ceg.EnergyGrid(cs, (1, 1, 1), eg.ewald_precision, True,
                             (outs[g].astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32), device=dev)
        assert np.array_equal(ceg.interpolate_grid(eg, pos), ceg.interpolate_grid(ref, pos))
        ref.close()
    sb.close(barrier=dist.barrier)
    for eg in sb.grids:
        eg.close()
    open(os.path.join(outdir, f"ok{rank}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_build_across_processes_writes_every_replica(cuda_lib, tmp_path):
    """ShardedGridBuild with one process per rank: each rank builds its x-slab and its kernel writes the records into
    the grids of BOTH processes (CUDA IPC mappings; NVLink peer stores when the ranks sit on different GPUs)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_record_build_and_ipc_argument_errors(cuda_lib):
    """Error behaviour of the round-2 entry points: nothing is dereferenced before the checks pass."""
    inp, cset, _ = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 1.0)
    gs = ceg._lib.make_gridspec(cs)
    uc = ceg.grids.unitcell_struct(cs.cell)
    h = ceg._lib.grid_create_empty(0, uc, gs, True)
    ptr, nbytes, layout = ceg._lib.grid_records(h)
    assert ptr % 32 == 0 and nbytes >= 32 * int(np.prod(cs.nodes)) and layout in (0, 4)
    hd, off = ceg._lib.ipc_export(0, ptr)
    assert len(hd) == 64 and off == 0
    hd2, off2 = ceg._lib.ipc_export(0, ptr + 4096)       # an interior pointer: same allocation, its offset
    assert off2 == 4096
    with pytest.raises(ceg.CegError):
        ceg._lib.ipc_export(0, 0)
    with pytest.raises(ceg.CegError):
        ceg._lib.ipc_close(0, ptr)                         # not a mapping obtained from ceg_ipc_open
    with pytest.raises(ceg.CegError):
        ceg._lib.ipc_open(0, bytes(64), 0)                 # not a handle
    lam = ceg.vdw_scaling() + ceg.coulomb_scaling()
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    with pytest.raises(ceg.CegError):                      # host memory as a record destination
        buf = np.zeros(1 << 16, dtype=np.uint8)
        a = (buf.ctypes.data + 31) & ~31
        ceg._lib.grid_build_records(ps.handle, gs, 0, 2, *lam, [[a, a]], layout, [1.0, 1.0], None, 0)
    with pytest.raises(ceg.CegError):                      # bad slab
        ceg._lib.grid_build_records(ps.handle, gs, 3, 3, *lam, [[ptr, ptr]], layout, [1.0, 1.0], None, 0)
    ps.close()
    ceg._lib.grid_free(h)
    # a ShardedGridBuild of one rank is the single-GPU device-resident build
    sb = ceg.ShardedGridBuild(inp, cs, (1, 1, 1), (np.inf, 1e-6), 0, 1, exchange=None, device=0)
    sb.build()
    sb.close()
    outs = [np.zeros((8,) + tuple(cs.nodes), dtype=np.float32) for _ in range(2)]
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    ceg._lib.grid_build(ps.handle, gs, 0, cs.nodes[0], *lam, outs, None)
    ps.close()
    pos = wl.random_positions(cs.cell.mat, 5000, 3)
    for g, eg in enumerate(sb.grids):
        ref = ceg.EnergyGrid(cs, (1, 1, 1), eg.ewald_precision, True,
                             (outs[g].astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32))
        assert np.array_equal(ceg.interpolate_grid(eg, pos), ceg.interpolate_grid(ref, pos))
        ref.close(); eg.close()


@pytest.mark.parametrize("fused", [False, True], ids=["separate_grids", "pair_records"])
def test_latency_path_equals_batch_path(cuda_lib, fused):
    """ceg_energy_points takes its one-launch latency path (csrc/latency.cuh) for <= 256 placements of <= 32 atoms and
    the batched kernels above: the same placements must give the same VdW bits (blocked placements included) and the
    same Coulomb energy to 1e-9 (the reciprocal sum is reduced in a different fixed order), for a monatomic and a
    three-atom molecule, with and without the reciprocal term, from inline (kernel parameter) and mapped-memory inputs."""
    inp, cs, ex, outs, gK = _two_grids(0.5)
    n_fw = len(ex["site_charge"])
    fw = ceg.RASPASystem(ex["unit_mat"], ex["positions"][:n_fw], ["X"] * n_fw, np.ones(n_fw), ex["charges"][:n_fw], False)
    q = float(ex["probe_charge"])
    eg_v = ceg.EnergyGrid(cs, (1, 1, 1), np.inf, True, gK[0])
    eg_c = ceg.EnergyGrid(cs, (1, 1, 1), 1e-6, True, gK[1])
    ef = ceg.initialize_ewald(fw)
    block = ceg.blockfile_from_grid(eg_v)
    rng = np.random.default_rng(3)
    for charges, idx in ((np.array([q]), [0]), (np.array([q, -0.5 * q, -0.5 * q]), [0, 0, 0])):
        na = len(idx)
        template = np.array([[0.0, 0.0, 0.0], [1.1, 0.0, 0.0], [-1.1, 0.0, 0.0]])[:na]       # rigid-molecule geometry (static term)
        mol = ceg.RASPASystem(np.diag([np.inf] * 3), template, ["Na"] * na, np.ones(na), charges, True)
        setup = ceg.CrystalEnergySetup(fw, mol, eg_c, charges, [eg_v], idx, ef, None, block)
        setup.fused_pairs = fused
        centers = wl.random_positions(fw.mat, 20000, 17)
        pos = centers[:, None, :] + (rng.normal(size=(20000, na, 3)) * 0.8 if na > 1 else 0.0)
        for recip in (True, False):
            vb, cb = ceg.energy_points(setup, pos, reciprocal=recip)                 # batched kernels
            assert 0.02 < (vb == 1e100).mean() < 0.9 and np.isfinite(cb).all()
            for n in (1, 2, 5, 17, 64, 256):                                          # 17 x 3 atoms: beyond the inline limit
                off = 100 * n
                v, c = ceg.energy_points(setup, pos[off:off + n], reciprocal=recip)   # latency kernel
                assert np.array_equal(v, vb[off:off + n])
                assert (np.abs(c - cb[off:off + n]) <= 1e-9 * np.abs(cb[off:off + n]) + 1e-6).all()
        eg_v.unpair()
    eg_v.close(); eg_c.close()
