"""GPU parity of the guest–guest pair energies (energy.jl:339-427) and the batched
Monte-Carlo trial evaluation built from the three device pieces (montecarlo.jl:563-606)."""
import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import forcefields as FF
from ceg_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _forcefield():
    """CO2 sites + Na+ in a force field with shifted LJ, Buckingham+HardSphere and the real-space Ewald term."""
    alpha = ceg.raspa.ewald_alpha_forcefield(1e-6, 12.0)
    q = {"C_co2": 0.6512, "O_co2": -0.3256, "Na": 1.0, "Ar": 0.0}
    names = list(q)
    sd = {s: i for i, s in enumerate(names)}
    inputs = [(("C_co2", "C_co2"), FF.make_rule(ceg.Kind.LennardJones, [28.129, 2.757], True, 12.0)),
              (("O_co2", "O_co2"), FF.make_rule(ceg.Kind.LennardJones, [80.507, 3.033], True, 12.0)),
              (("Ar", "Ar"), FF.make_rule(ceg.Kind.LennardJones, [124.07, 3.38], True, 12.0)),
              (("Na", "Na"), FF.NoInteraction()),
              (("Na", "O_co2"), ceg.InteractionRuleSum.make([FF.make_rule(ceg.Kind.Buckingham, [5.581e7, 3.985, 9.167e5], False, 12.0),
                                                             FF.make_rule(ceg.Kind.HardSphere, [1.5, 0.0], False, 12.0)]))]
    ff = ceg.ForceField.build(inputs, ceg.Mixing.LorentzBerthelot, 12.0, True, False, sdict=sd)
    for a in names:
        for b in names:
            if q[a] != 0 and q[b] != 0:
                i, j = sd[a], sd[b]
                ff.interactions[i][j] = FF.sum_rules(ff.interactions[i][j],
                                                     ceg.InteractionRule(ceg.Kind.CoulombEwaldDirect, (alpha, q[a], q[b]), 0.0, False))
    return ff, names


def _environment(ff, names, n_mol, seed):
    rng = np.random.default_rng(seed)
    mat = ceg.mat_from_parameters((26.0, 25.0, 27.0), (88.0, 95.0, 101.0))
    tab = ceg.pair_energy_table(ff, names)
    geom = np.array([[0, 0, 1.149], [0, 0, 0.0], [0, 0, -1.149]])
    pos, sp, mol = [], [], []
    for m in range(n_mol):
        c = mat @ rng.random(3)
        kind = m % 3
        if kind == 0:       # CO2 with a random orientation
            a = rng.standard_normal(3); a /= np.linalg.norm(a)
            pos += [c + 1.149 * a, c, c - 1.149 * a]; sp += [1, 0, 1]; mol += [m] * 3
        elif kind == 1:
            pos.append(c); sp.append(2); mol.append(m)   # Na
        else:
            pos.append(c); sp.append(3); mol.append(m)   # Ar
    return ceg.GuestEnvironment(mat, 12.0, tab, np.array(pos), np.array(sp, dtype=np.int32), np.array(mol, dtype=np.int32)), geom


def test_single_contribution_vdw_batched(cuda_lib, orc):
    ff, names = _forcefield()
    env, geom = _environment(ff, names, 60, 1)
    assert {int(f) for f in np.unique(env.table["flags"])} >= {0, 2, 8, 10}
    rng = np.random.default_rng(2)
    centers = (env.mat @ rng.random((3, 4000))).T
    trials = centers[:, None, :] + geom[None, :, :]
    got = ceg.single_contribution_vdw(env, [1, 0, 1], trials, skip_molecule=0)
    want = orc.guest_energy(env, [1, 0, 1], trials, 0)
    assert np.array_equal(np.isinf(got), np.isinf(want))
    fin = np.isfinite(want)
    assert (np.abs(got[fin] - want[fin]) <= 1e-9 * np.abs(want[fin]) + 1e-6).all()
    # a Na+ trial: Buckingham + hard sphere + Ewald pairs, some placements inside a hard sphere (Inf)
    na_trials = np.concatenate([centers[:500, None, :], env.positions[env.species == 1][:20, None, :] + 0.4])
    got = ceg.single_contribution_vdw(env, [2], na_trials, skip_molecule=1)
    want = orc.guest_energy(env, [2], na_trials, 1)
    assert np.isinf(want).any() and np.array_equal(np.isinf(got), np.isinf(want))
    fin = np.isfinite(want)
    assert (np.abs(got[fin] - want[fin]) <= 1e-9 * np.abs(want[fin]) + 1e-6).all()


def test_compute_vdw_and_move_identity(cuda_lib, orc):
    """compute_vdw (energy.jl:339-385) and the identity the Monte-Carlo loop relies on: moving molecule m
    changes the total by single(new) - single(old)."""
    ff, names = _forcefield()
    env, geom = _environment(ff, names, 45, 3)
    total = ceg.compute_vdw(env)
    # oracle: explicit double loop over pairs of different molecules
    ref = 0.5 * sum(orc.guest_energy(env, env.species[env.molecule == m], env.positions[env.molecule == m][None], int(m))[0]
                    for m in np.unique(env.molecule))
    assert np.isfinite(total) and abs(total - ref) <= 1e-9 * abs(ref) + 1e-6
    m = 6                                                # a CO2 molecule (6 % 3 == 0)
    sel = env.molecule == m
    new = env.positions[sel] + np.array([1.3, -0.7, 2.1])
    old_c = ceg.single_contribution_vdw(env, env.species[sel], env.positions[sel][None], m)[0]
    new_c = ceg.single_contribution_vdw(env, env.species[sel], new[None], m)[0]
    import dataclasses
    pos2 = env.positions.copy(); pos2[sel] = new
    total2 = ceg.compute_vdw(dataclasses.replace(env, positions=pos2))
    assert abs((total2 - total) - (new_c - old_c)) <= 1e-9 * max(abs(total), abs(total2)) + 1e-6
    # more than one shared-memory tile of environment atoms
    big, _ = _environment(ff, names, 1500, 5)
    assert len(big.positions) > 2048
    t = (big.mat @ np.random.default_rng(0).random((3, 300))).T[:, None, :]
    g = ceg.single_contribution_vdw(big, [3], t, -1)
    w = orc.guest_energy(big, [3], t, -1)
    assert (np.abs(g - w) <= 1e-9 * np.abs(w) + 1e-6).all()


def test_grid_vs_explicit_supercell_cit7(cuda_lib, orc):
    """test/runtests.jl:186-198 ("Triclinic"): the energy of Ar in CIT-7 from the interpolated grid agrees (rtol
    5e-3, the reference's tolerance) with the pair sum over the explicit (2,3,3) supercell atoms treated as guest
    molecules; and that pair sum (ceg_guest_energy) equals the un-interpolated grid-point sum of the literal
    derivative kernel (ceg_compute_derivatives, channel 0) to 1e-9 — two independent device code paths."""
    inp, cset, _ = wl.load_fixture("cit7_multi")                 # probes Ar, C_co2, O_co2 + Coulomb
    nt = inp.n_types
    tab = np.zeros((nt + 1, nt + 1), dtype=ceg.pair_energy_dtype)
    for t in range(nt):
        for name in ("flags", "hs_r", "lj_eps", "lj_sigma", "lj_shift", "buck_A", "buck_B", "buck_C", "buck_shift"):
            tab[t, nt][name] = tab[nt, t][name] = inp.table[0, t][name]
    n = len(inp.positions)
    env = ceg.GuestEnvironment(inp.mat, 12.0, tab, inp.positions, inp.types, np.arange(n, dtype=np.int32))
    pts = np.concatenate([np.array([[5.86207, 16.23616, 17.52779]]), wl.random_positions(cset.cell.mat, 200, 8)])
    pair_sum = ceg.single_contribution_vdw(env, [nt], pts[:, None, :])
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    exact = ps.compute_derivatives(pts)[0][:, 0]
    both = np.isfinite(pair_sum) & np.isfinite(exact)
    assert both.sum() > 150 and np.array_equal(np.isfinite(pair_sum), np.isfinite(exact))
    assert (np.abs(pair_sum[both] - exact[both]) <= 1e-9 * np.abs(exact[both]) + 1e-6).all()
    # the Ar grid at 0.15 Å and its interpolation at the reference's point
    nx, ny, nz = cset.nodes
    outs = [np.empty((8, nx, ny, nz), dtype=np.float32) for _ in range(4)]
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cset), 0, nx, lam_v, thr_v, lam_c, thr_c, outs, None)
    ps.close()
    g = ceg.EnergyGrid(cset, inp.supercell, np.inf, True, orc.scale_to_kelvin(outs[0]))
    e_grid = ceg.interpolate_grid(g, pts[:1])[0]
    g.close()
    # (a repulsive site, ~2.5e5 K: the tricubic interpolant is 1.4e-3 off the exact sum there)
    assert abs(e_grid - pair_sum[0]) <= 5e-3 * abs(pair_sum[0]), (e_grid, pair_sum[0])


def test_resident_environment_update_and_stateless_form(cuda_lib, orc):
    """ceg_guest_env_*: the environment stays on the device, accepted moves update it in place; the stateless
    ceg_guest_energy gives the same numbers."""
    ff, names = _forcefield()
    env, geom = _environment(ff, names, 45, 3)
    rng = np.random.default_rng(9)
    trials = (env.mat @ rng.random((3, 3000))).T[:, None, :] + geom[None]
    e0 = ceg.single_contribution_vdw(env, [1, 0, 1], trials, skip_molecule=0)
    stateless = ceg._lib.guest_energy(0, env.cell_struct(), 144.0, env.table, env.positions, env.species, env.molecule,
                                      [1, 0, 1], 0, trials)
    assert np.array_equal(e0, stateless)
    # move molecule 3 (a CO2: atoms of molecule id 3) somewhere else and compare with a fresh environment / the oracle
    sel = np.nonzero(env.molecule == 3)[0]
    new_pos = env.positions[sel] + np.array([3.1, -2.2, 1.7])
    env.move(int(sel[0]), new_pos)
    e1 = ceg.single_contribution_vdw(env, [1, 0, 1], trials, skip_molecule=0)
    want = orc.guest_energy(env, [1, 0, 1], trials, 0)
    fin = np.isfinite(want)
    assert np.array_equal(np.isinf(e1), np.isinf(want)) and (np.abs(e1[fin] - want[fin]) <= 1e-9 * np.abs(want[fin]) + 1e-6).all()
    assert not np.array_equal(e0, e1)
    env.close()


def test_guest_kernel_shapes_rotated_cell_and_cutoff_band(cuda_lib, orc):
    """The round-2 guest kernel (one thread per molecule atom, queued in-cutoff pairs): molecules whose atom count does
    not divide the CTA (5), exceeds the threads per placement (40 > 32), a cell that is not upper triangular (the general
    3x3 product instead of the triangular one), a skipped molecule, and pairs sitting on the cutoff, where the literal
    re-measurement decides exactly as the oracle does."""
    import dataclasses
    ff, names = _forcefield()
    env, geom = _environment(ff, names, 400, 11)              # > 512 atoms: several tiles
    rng = np.random.default_rng(12)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))           # a proper rotation of the whole system
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    rot = dataclasses.replace(env, mat=q @ env.mat, positions=env.positions @ q.T)
    assert abs(rot.mat[1, 0]) > 1e-3 and abs(rot.mat[2, 1]) > 1e-3        # not triangular
    for n_at in (5, 40):
        sp = rng.integers(0, 4, n_at).astype(np.int32)
        shape = rng.standard_normal((n_at, 3)) * 1.5
        centers = (env.mat @ rng.random((3, 700))).T
        trials = centers[:, None, :] + shape[None]
        want = orc.guest_energy(env, sp, trials, 7)
        got = ceg.single_contribution_vdw(env, sp, trials, skip_molecule=7)
        got_rot = ceg.single_contribution_vdw(rot, sp, trials @ q.T, skip_molecule=7)
        fin = np.isfinite(want)
        assert fin.sum() > 50 and np.array_equal(np.isfinite(got), fin) and np.array_equal(np.isfinite(got_rot), fin)
        for g in (got, got_rot):
            assert (np.abs(g[fin] - want[fin]) <= 1e-9 * np.abs(want[fin]) + 1e-6).all()
    # pairs on the cutoff: one O_co2 guest alone in a big cell, Na trials at 12 Å -/+ a few ulps along several directions.
    # The pair energy is -0.031 K just inside the cutoff and 0 beyond it, so the in/out decision is visible.
    tab = env.table
    big = ceg.GuestEnvironment(np.diag([40.0, 41.0, 42.0]), 12.0, tab, np.array([[20.0, 20.0, 20.0]]),
                               np.array([1], dtype=np.int32), np.array([0], dtype=np.int32))
    dirs = rng.standard_normal((200, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    r = 12.0 * (1.0 + rng.integers(-3, 4, 200) * 2.2e-16)
    trials = (20.0 + dirs * r[:, None])[:, None, :]
    want = orc.guest_energy(big, [2], trials, -1)
    got = ceg.single_contribution_vdw(big, [2], trials, -1)
    assert (want == 0).any() and (want < -0.01).any()
    assert np.array_equal(got == 0, want == 0)
    assert (np.abs(got - want) <= 1e-9 * np.abs(want) + 1e-6).all()
