"""setup_RASPA end to end on a synthetic RASPA directory written by the test itself (raspa.jl:472-530,
403-439): CIF + force-field .def files -> fused creation of the missing `.grid` files on the device ->
parse -> energy_point, with the file cache reused on the second call."""
import os

import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import workloads as wl

pytestmark = pytest.mark.gpu

MIXING = """# general rule for shifted vs truncated
truncated
# general rule tailcorrections
no
# number of defined interactions
5
# type interaction
Oz      lennard-jones   53.0   3.30
Siz     lennard-jones   22.0   2.30
C_co2   lennard-jones   27.0   2.80
O_co2   lennard-jones   79.0   3.05
Na      none
# general mixing rule for Lennard-Jones
Lorentz-Berthelot
"""
FFDEF = """# rules to overwrite
0
# number of defined interactions
2
# type type2 interaction
Na   Oz   buckingham2   5.581e+07   3.9850   9.167e+05   1.50
O_co2 Siz  lennard-jones 39.05 3.49
# mixing rules to overwrite
0
"""
PSEUDO = """#number of pseudo atoms
5
#type print as chem oxidation mass charge polarization B-factor radii connectivity anisotropic anisotropic-type tinker-type
Oz     yes O  O  0 15.9994 -1.025  0.0 1.0 0.5  2 0 absolute 0
Siz    yes Si Si 0 28.0855  2.05   0.0 1.0 1.18 4 0 absolute 0
C_co2  yes C  C  0 12.0107  0.6512 0.0 1.0 0.72 0 0 absolute 0
O_co2  yes O  O  0 15.9994 -0.3256 0.0 1.0 0.68 0 0 absolute 0
Na     yes Na Na 0 22.98977 1.0    0.0 1.0 1.0  0 0 absolute 0
"""
CO2 = """# critical constants
304.1282
7377300.0
0.22394
#Number Of Atoms
 3
# Number of groups
1
# CO2-group
rigid
# number of atoms
3
# atomic positions
0 O_co2     0.0           0.0           1.149
1 C_co2     0.0           0.0           0.0
2 O_co2     0.0           0.0          -1.149
"""


def _make_raspa_dir(root):
    rng = np.random.default_rng(7)
    os.makedirs(root / "structures" / "cif")
    os.makedirs(root / "forcefield" / "ToyFF")
    os.makedirs(root / "molecules" / "TraPPE")
    lines = ["data_Toy", "_cell_length_a 13.1", "_cell_length_b 12.4", "_cell_length_c 11.7", "_cell_angle_alpha 90",
             "_cell_angle_beta 97.5", "_cell_angle_gamma 90", "loop_", "_atom_site_label", "_atom_site_type_symbol",
             "_atom_site_fract_x", "_atom_site_fract_y", "_atom_site_fract_z"]
    n = 0
    for i in range(4):
        for j in range(4):
            for k in range(3):
                f = (np.array([i, j, k]) + 0.5) / np.array([4, 4, 3]) + (rng.random(3) - 0.5) * 0.08
                n += 1
                sp = ("Siz", "Si") if n % 3 == 0 else ("Oz", "O")
                lines.append(f"{sp[0]}_{n} {sp[1]} {f[0]:.6f} {f[1]:.6f} {f[2]:.6f}")
    (root / "structures" / "cif" / "Toy.cif").write_text("\n".join(lines) + "\n")
    (root / "forcefield" / "ToyFF" / "force_field_mixing_rules.def").write_text(MIXING)
    (root / "forcefield" / "ToyFF" / "force_field.def").write_text(FFDEF)
    (root / "forcefield" / "ToyFF" / "pseudo_atoms.def").write_text(PSEUDO)
    (root / "molecules" / "TraPPE" / "CO2.def").write_text(CO2)


def test_setup_raspa_co2_fused_grids_and_cache(cuda_lib, orc, tmp_path):
    _make_raspa_dir(tmp_path)
    ceg.setdir_RASPA(str(tmp_path))
    setup = ceg.setup_RASPA("Toy", "ToyFF", "CO2", "TraPPE", gridstep=0.4, blockfile=False)
    assert setup.atomsidx == [0, 1, 0] and len(setup.grids) == 2               # O_co2, C_co2
    assert setup.coulomb.ewald_precision == 1e-6 and all(g.ewald_precision == np.inf for g in setup.grids)
    fw = setup.framework
    assert ceg.find_supercell(fw.mat, 12.0) == (2, 2, 3) == setup.grids[0].num_unitcell
    setup.wait()        # the files of freshly created grids are written by a background thread
    gdir = tmp_path / "grids" / "ToyFF" / "Toy" / "0.400000"
    files = sorted(p.name for p in gdir.rglob("*.grid"))
    assert files == ["Toy_C_co2_truncated.grid", "Toy_Electrostatics_Ewald.grid", "Toy_O_co2_truncated.grid"]
    # the files equal what the oracle builds grid by grid (one pass per grid, like grids.jl:137-185)
    ff = setup.forcefield
    cset = setup.grids[0].csetup
    nx, ny, nz = cset.nodes
    for atom, eg in (("O_co2", setup.grids[0]), ("C_co2", setup.grids[1])):
        inp = ceg.probe_inputs(fw, ff, (atom,))
        want, _ = orc.grid_build(inp, cset)
        raw = np.fromfile(gdir / f"Toy_{atom}_truncated.grid", dtype="<f4", offset=128, count=8 * nx * ny * nz)
        got = raw.reshape(8, nx, ny, nz)
        neq = got != want[0]
        assert neq.mean() < 1e-4 and np.abs(got - want[0])[neq].max(initial=0) <= 1e-6 * np.abs(want[0]).max()
        assert np.array_equal(eg.grid, orc.scale_to_kelvin(got))
    inp = ceg.probe_inputs(fw, ff, (), True, setup.ewald.alpha)
    want, _ = orc.grid_build(inp, cset)
    got = np.fromfile(gdir / "2x2x3" / "Toy_Electrostatics_Ewald.grid", dtype="<f4", offset=136,
                      count=8 * nx * ny * nz).reshape(8, nx, ny, nz)
    assert (got != want[0]).mean() < 1e-4
    # energy_point vs the oracle composition on the same Float32 grids
    centers = wl.random_positions(fw.mat, 2000, 9)
    mol = np.asarray(setup.molecule.position)
    pos = centers[:, None, :] + mol[None, :, :]
    v, c = ceg.energy_points(setup, pos)
    oe = orc.OracleEwald(fw)
    ctx = setup.ewald_ctx()
    wv = np.zeros(len(pos))
    wc = oe.reciprocal(pos, setup.charges, ctx.energy_net_charges, ctx.static_contribution)
    blocked = np.zeros(len(pos), dtype=bool)
    for a in range(3):
        e = orc.interpolate(cset, setup.grids[setup.atomsidx[a]].grid, True, pos[:, a])
        blocked |= e == 1e100
        wv += e
        wc += setup.charges[a] * orc.interpolate(cset, setup.coulomb.grid, False, pos[:, a])
    ok = ~blocked
    assert np.array_equal(v >= 1e99, blocked)
    assert (np.abs(v[ok] - wv[ok]) <= 1e-9 * np.abs(wv[ok]) + 1e-6).all()
    assert (np.abs(c[ok] - wc[ok]) <= 1e-9 * np.abs(wc[ok]) + 1e-6).all()
    # second call: the cache is reused (no file rewritten)
    mt = {p: os.path.getmtime(p) for p in gdir.rglob("*.grid")}
    setup2 = ceg.setup_RASPA("Toy", "ToyFF", "CO2", "TraPPE", gridstep=0.4, blockfile=False)
    assert {p: os.path.getmtime(p) for p in gdir.rglob("*.grid")} == mt
    assert np.array_equal(setup2.grids[0].grid, setup.grids[0].grid)
    # a single-atom probe defined in the framework force field (raspa.jl:526-529)
    s3 = ceg.setup_RASPA("Toy", "ToyFF", "Na", gridstep=0.4)
    s3.wait()
    assert s3.block.empty and s3.charges[0] == 1.0 and (gdir / "Toy_Na_truncated.grid").exists()
    e = ceg.energy_point(s3, np.array([[3.0, 4.0, 5.0]]))
    assert np.isfinite(e[1]) and (np.isfinite(e[0]) or e[0] == 1e100)
    # batched Monte-Carlo trial evaluation (montecarlo.jl:563-606) = the three device pieces side by side
    names = ["O_co2", "C_co2"]
    tab = ceg.pair_energy_table(ff, names)
    envc = wl.random_positions(fw.mat, 12, 3)
    env_pos = (envc[:, None, :] + mol[None, :, :]).reshape(-1, 3)
    env = ceg.GuestEnvironment(fw.mat, 12.0, tab, env_pos, np.tile(np.array(setup.atomsidx, dtype=np.int32), 12),
                               np.repeat(np.arange(12, dtype=np.int32), 3))
    rest = setup.ewald.StoreRigidChargeFramework + ceg.adsorbate_sums(setup.ewald, np.tile(setup.charges, 11), env_pos[3:])
    fv, fc, rec, gg = ceg.movement_energies(setup, env, rest, setup.atomsidx, pos[:500], skip_molecule=0)
    assert np.array_equal(fv, v[:500])
    wdir = sum(setup.charges[a] * orc.interpolate(cset, setup.coulomb.grid, False, pos[:500, a]) for a in range(3))
    okk = ok[:500]
    assert (np.abs(fc[okk] - wdir[okk]) <= 1e-9 * np.abs(wdir[okk]) + 1e-6).all()
    oe.S = rest
    wrec = oe.reciprocal(pos[:500], setup.charges, 0.0, 0.0)
    assert np.abs(rec - wrec).max() <= 1e-9 * np.abs(wrec).max()
    wgg = orc.guest_energy(env, setup.atomsidx, pos[:500], 0)
    assert np.abs(wgg).max() > 0 and (np.abs(gg - wgg) <= 1e-9 * np.abs(wgg) + 1e-6).all()
    # energy_grid (grids.jl:346-424): lattice generated on the device == energy_point over the host-built lattice
    th = 0.7
    rots = [np.eye(3), np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]])]
    eg = ceg.energy_grid(setup, 1.1, num_rotate=2, rotations=rots)
    m = np.asarray(fw.mat, dtype=np.float64)
    nums = [int(np.floor(np.linalg.norm(m[:, d]) / 1.1)) + 1 for d in range(3)]
    steps = [m[:, d] / nums[d] for d in range(3)]
    assert eg.shape == (2, *nums)
    iA, iB, iC = np.meshgrid(*[np.arange(k) for k in nums], indexing="ij")
    ofs = ((iA[..., None] * steps[0] + iB[..., None] * steps[1]) + iC[..., None] * steps[2]).reshape(-1, 3)
    for k, R in enumerate(rots):
        v2, c2 = ceg.energy_points(setup, ofs[:, None, :] + (mol @ R.T)[None, :, :])
        assert np.array_equal(eg[k].ravel(), v2 + c2)
    assert (eg == 1e100).any() and (eg < 0).any()
