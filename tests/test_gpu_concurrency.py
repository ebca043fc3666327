"""The C ABI is called from concurrent host threads by the reference's callers (`interpolate_grid` and
`compute_ewald` run in parallel tasks, montecarlo.jl:573-578; grid builders run under `@threads`).  Several
Python threads (ctypes releases the GIL during the call) share one grid handle, one Ewald handle and the
process-wide device allocation pool; every result must equal the serial one bit for bit."""
import threading

import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _run_threads(fn, n):
    out, err = [None] * n, []

    def work(i):
        try:
            out[i] = fn(i)
        except Exception as e:   # noqa: BLE001
            err.append(e)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not err, err
    return out


def test_shared_handles_from_many_threads(cuda_lib):
    inp, cset, ex = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 0.6)
    rng = np.random.default_rng(0)
    nx, ny, nz = cs.nodes
    grid = rng.standard_normal((8, nx, ny, nz)).astype(np.float32)
    eg = ceg.EnergyGrid(cs, (1, 1, 1), 1e-6, True, grid)
    n_unit = len(ex["site_charge"])
    fw = ceg.RASPASystem(ex["unit_mat"], ex["positions"][:n_unit], ["X"] * n_unit, np.ones(n_unit), ex["charges"][:n_unit], False)
    ef = ceg.initialize_ewald(fw)
    q = np.array([-0.3256, 0.6512, -0.3256])
    geom = np.array([[0, 0, 1.149], [0, 0, 0.0], [0, 0, -1.149]])
    batches = [wl.random_positions(cs.cell.mat, 20000 + 1000 * i, 100 + i) for i in range(8)]

    def interp(i):
        return ceg.interpolate_grid(eg, batches[i])

    def recip(i):
        return ceg._lib.ewald_reciprocal(ef.handle, batches[i][:3000, None, :] + geom[None], q, 0.5 * i, -1.0)

    serial_i = [interp(i) for i in range(8)]
    serial_r = [recip(i) for i in range(8)]
    for _ in range(3):
        got = _run_threads(lambda i: (interp(i), recip(i)), 8)
        for i in range(8):
            assert np.array_equal(got[i][0], serial_i[i]) and np.array_equal(got[i][1], serial_r[i])
    eg.close()


def test_concurrent_builds_of_disjoint_slabs(cuda_lib):
    """Two host threads build disjoint x-slabs of the same host array through two probe handles (the multi-GPU
    recipe of INTEGRATION.md on one device): same bytes as one call over the whole range."""
    inp, cset, _ = wl.load_fixture("cha_na")
    cs = wl.respace(cset, 0.5)
    nx, ny, nz = cs.nodes
    gs = ceg._lib.make_gridspec(cs)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    ref = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    ceg._lib.grid_build(ps.handle, gs, 0, nx, lam_v, thr_v, lam_c, thr_c, ref, None)
    ps.close()
    outs = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
    cuts = [0, nx // 3, (2 * nx) // 3 + 1, nx]

    def build(i):
        p = ceg.ProbeSystem(None, None, inputs=inp)
        ceg._lib.grid_build(p.handle, gs, cuts[i], cuts[i + 1], lam_v, thr_v, lam_c, thr_c, outs, None)
        p.close()
    _run_threads(build, 3)
    for g in range(2):
        assert np.array_equal(outs[g], ref[g])
