"""RASPA input parsing on a synthetic RASPA directory written by the test itself (no GPU, no reference checkout):
CIF -> framework, pseudo atoms, molecule definition, force-field files -> ForceField with the reference's rules
(raspa.jl:533-702, forcefields.jl:204-265), grid cache locations (raspa.jl:403-418), pair-table flattening."""
import math
import os

import numpy as np
import pytest

import ceg_b200 as ceg
from ceg_b200 import forcefields as FF
from test_gpu_setup_raspa import _make_raspa_dir


@pytest.fixture()
def raspa_dir(tmp_path):
    _make_raspa_dir(tmp_path)
    old = ceg.raspa._RASPADIR[0]
    ceg.setdir_RASPA(str(tmp_path))
    yield tmp_path
    ceg.raspa._RASPADIR[0] = old


def test_framework_from_cif_and_pseudo_atoms(raspa_dir):
    fw = ceg.load_framework_RASPA("Toy", "ToyFF")
    assert len(fw.position) == 48 and fw.mat.shape == (3, 3)
    # a || x, b in the xy plane (Chemfiles convention)
    assert fw.mat[1, 0] == 0 and fw.mat[2, 0] == 0 and fw.mat[2, 1] == 0 and abs(fw.mat[0, 0] - 13.1) < 1e-12
    beta = math.degrees(math.acos(fw.mat[:, 0] @ fw.mat[:, 2] / (13.1 * 11.7)))
    assert abs(beta - 97.5) < 1e-9
    sym = [FF.get_atom_name(str(s)) for s in fw.atomic_symbol]     # labels are Oz_1, Siz_3, ...: `_123` is stripped
    assert sym.count("Siz") == 16 and sym.count("Oz") == 32
    q = np.asarray(fw.atomic_charge)
    assert np.allclose(q[[s == "Siz" for s in sym]], 2.05) and np.allclose(q[[s == "Oz" for s in sym]], -1.025)
    assert abs(q.sum()) < 1e-9
    # Cartesian = mat . fractional, no wrapping
    frac = np.asarray(fw.position) @ np.linalg.inv(fw.mat).T
    assert (frac > -0.05).all() and (frac < 1.05).all()


def test_molecule_and_default_system(raspa_dir):
    co2 = ceg.load_molecule_RASPA("CO2", "TraPPE", "ToyFF")
    assert [str(s) for s in co2.atomic_symbol] == ["O_co2", "C_co2", "O_co2"]
    assert np.allclose(co2.position, [[0, 0, 1.149], [0, 0, 0], [0, 0, -1.149]])
    assert np.allclose(co2.atomic_charge, [-0.3256, 0.6512, -0.3256])
    na = ceg.raspa.default_system("Na", "ToyFF")
    assert len(na.position) == 1 and na.atomic_charge[0] == 1.0


def test_forcefield_rules(raspa_dir):
    """Mixing (Lorentz-Berthelot), `none`, the explicit pairs of force_field.def (buckingham2 = Buckingham + HardSphere),
    the Ewald term on charged pairs and the alpha of raspa.jl:682-687."""
    ff = ceg.parse_forcefield_RASPA("ToyFF")
    sd = ff.sdict
    assert list(sd) == ["Oz", "Siz", "C_co2", "O_co2", "Na"]          # line order of the mixing-rules file
    kinds = lambda r: [x.kind for x in FF._rules_of(r)]
    # mixed pair: eps = sqrt(e1 e2), sigma = (s1 + s2) / 2, truncated -> no shift
    lj = [x for x in FF._rules_of(ff.interactions[sd["Oz"]][sd["C_co2"]]) if x.kind is ceg.Kind.LennardJones][0]
    assert abs(lj.params[0] - math.sqrt(53.0 * 27.0)) < 1e-12 and abs(lj.params[1] - (3.30 + 2.80) / 2) < 1e-12 and lj.shift == 0.0
    # explicit override
    lj2 = [x for x in FF._rules_of(ff.interactions[sd["O_co2"]][sd["Siz"]]) if x.kind is ceg.Kind.LennardJones][0]
    assert lj2.params == (39.05, 3.49)
    # buckingham2 and symmetry
    for a, b in (("Na", "Oz"), ("Oz", "Na")):
        k = kinds(ff.interactions[sd[a]][sd[b]])
        assert ceg.Kind.Buckingham in k and ceg.Kind.HardSphere in k and ceg.Kind.CoulombEwaldDirect in k
    # `none` with anything is no VdW interaction (only the Ewald term remains on charged pairs)
    assert set(kinds(ff.interactions[sd["Na"]][sd["Siz"]])) <= {ceg.Kind.CoulombEwaldDirect, ceg.Kind.NoInteraction}
    alpha = ceg.raspa.ewald_alpha_forcefield(1e-6, 12.0)
    eps = math.log(1e-6 * 12.0)
    assert abs(alpha - math.sqrt(abs(eps + math.log(math.sqrt(abs(eps))))) / 12.0) < 1e-15
    ew = [x for x in FF._rules_of(ff.interactions[sd["Na"]][sd["Oz"]]) if x.kind is ceg.Kind.CoulombEwaldDirect][0]
    assert abs(ew.params[0] - alpha) < 1e-15 and ew.params[1] * ew.params[2] == pytest.approx(-1.025)


def test_pair_table_flattening_and_grid_needs(raspa_dir):
    ff = ceg.parse_forcefield_RASPA("ToyFF")
    fw = ceg.load_framework_RASPA("Toy", "ToyFF")
    inp = ceg.probe_inputs(fw, ff, ("Na", "O_co2"), True, 0.265)
    assert inp.table.shape[0] == 2 and inp.supercell == ceg.find_supercell(fw.mat, 12.0)
    names = {int(t): FF.get_atom_name(str(s)) for t, s in zip(inp.types[:48], fw.atomic_symbol)}
    for t, base in names.items():
        na_row, o_row = inp.table[0, t], inp.table[1, t]
        if base == "Oz":
            assert na_row["flags"] == 5 and na_row["hs_r"] == 1.5 and na_row["buck_A"] == 5.581e7      # Buckingham + HS
            assert o_row["flags"] == 2
        else:
            assert na_row["flags"] == 0 and o_row["flags"] == 2 and o_row["lj_eps"] == 39.05
    assert len(inp.positions) == 48 * int(np.prod(inp.supercell)) and inp.want_coulomb
    assert FF.needsvdwgrid(ff, "Na") and FF.needsvdwgrid(ff, "C_co2")


def test_grid_locations(raspa_dir):
    ff = ceg.parse_forcefield_RASPA("ToyFF")
    coul, vdw = ceg.raspa.grid_locations("Toy", "ToyFF", ff, ["O_co2", "C_co2"], 0.4, (2, 2, 3))
    base = os.path.join(str(raspa_dir), "grids", "ToyFF", "Toy", "0.400000")
    assert coul == os.path.join(base, "2x2x3", "Toy_Electrostatics_Ewald.grid")
    assert vdw == [os.path.join(base, "Toy_O_co2_truncated.grid"), os.path.join(base, "Toy_C_co2_truncated.grid")]


def test_guest_pair_energy_table_and_oracle_rule(raspa_dir, orc):
    """energy.py: the (species x species) table of rule(d^2) (interactions.jl:367-409) built from the parsed force
    field, and the oracle's pair energy against the host-side InteractionRule.__call__ (written from the same
    reference lines, independently of the C code)."""
    import ctypes as C
    ff = ceg.parse_forcefield_RASPA("ToyFF")
    names = ["O_co2", "C_co2", "Na", "Oz"]
    tab = ceg.pair_energy_table(ff, names)
    assert tab.shape == (4, 4) and np.array_equal(tab["flags"], tab["flags"].T)
    na_o = tab[2, 3]
    assert na_o["flags"] == 1 | 4 | 8 and na_o["hs_r"] == 1.5 and na_o["buck_B"] == 3.985       # HS + Buckingham + Ewald
    assert na_o["ewald_qq"] == pytest.approx(ceg.COULOMBIC_CONVERSION_FACTOR * 1.0 * -1.025)
    assert tab[0, 1]["flags"] == 2 | 8 and tab[0, 1]["ewald_qq"] == pytest.approx(ceg.COULOMBIC_CONVERSION_FACTOR * -0.3256 * 0.6512)
    # oracle rule(d^2) == host rule(r) on a few distances (outside the hard sphere)
    sd = ff.sdict
    env = ceg.GuestEnvironment(np.diag([60.0, 60.0, 60.0]), 12.0, tab, np.zeros((1, 3)), np.array([3], dtype=np.int32),
                               np.array([0], dtype=np.int32))
    for (a, b) in ((2, 3), (0, 3), (1, 3)):
        env.species[:] = b
        rule = ff.interactions[sd[names[a]]][sd[names[b]]]
        for r in (1.7, 2.9, 5.5, 11.0):
            got = orc.guest_energy(env, [a], np.array([[[r, 0.0, 0.0]]]), 1)[0]
            assert got == pytest.approx(rule(r), rel=1e-12, abs=1e-12), (names[a], names[b], r)
    # inside the hard sphere: +Inf; beyond the cutoff: 0
    env.species[:] = 3
    assert np.isinf(orc.guest_energy(env, [2], np.array([[[1.2, 0.0, 0.0]]]), 1)[0])
    assert orc.guest_energy(env, [2], np.array([[[12.5, 0.0, 0.0]]]), 1)[0] == 0.0
