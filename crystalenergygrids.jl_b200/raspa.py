"""RASPA2 input front-end: CIF frameworks, force-field .def files, molecule .def files,
.grid cache locations and `setup_RASPA`.

Mirrors /root/reference/src/raspa.jl (host-side text parsing; runs once per setup):
  setdir_RASPA! / getdir_RASPA          raspa.jl:12-34
  parse_pseudoatoms_RASPA               raspa.jl:95-134
  parse_molecule_RASPA                  raspa.jl:192-214
  RASPASystem                           raspa.jl:222-230
  load_framework_RASPA                  raspa.jl:303-317   (Chemfiles CIF convention)
  load_molecule_RASPA / default_system  raspa.jl:344-385
  decide_parse_block / parse_block      raspa.jl:381-400
  grid_locations                        raspa.jl:403-418
  retrieve_or_create_grid               raspa.jl:420-439
  setup_RASPA                           raspa.jl:472-530
  parse_interaction_RASPA … parse_forcefield_RASPA   raspa.jl:533-702
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import forcefields as FF
from .constants import GRID_TO_KELVIN
from .cell import find_supercell, mat_from_parameters
from .constants import CUTOFF_COULOMB
from .forcefields import ForceField, InteractionRule, InteractionRuleSum, Kind, Mixing, get_atom_name

_RASPADIR = [os.path.join(os.path.expanduser("~"), "RASPA2", "simulations", "share", "raspa")]


def setdir_RASPA(path: str) -> None:
    _RASPADIR[0] = str(path)


def getdir_RASPA() -> str:
    ret = _RASPADIR[0]
    if not os.path.isdir(ret):
        raise FileNotFoundError(f"Could not find raspa directory at the given path {ret}. "
                                "Please set the correct path through the `setdir_RASPA` function.")
    return ret


# --------------------------------------------------------------------------------------------
@dataclass
class PseudoAtomInfo:
    type: str
    printas: str
    symbol: str
    oxidation: float
    mass: float
    charge: float
    polarization: float
    Bfactor: float
    radius: float
    connectivity: float
    anisotropic: float
    anisotropy_absolute: bool
    tinker_type: int


@dataclass
class PseudoAtomListing:
    exact: dict
    open_ended: list    # (prefix, line index)
    info: list

    def get_strict(self, atom) -> PseudoAtomInfo:
        name = str(atom)
        prio = self.exact.get(name, -1)
        for prefix, j in self.open_ended:
            if name.startswith(prefix) and j > prio:
                prio = j
        if prio < 0:
            raise KeyError(f"Atom {atom} not found in pseudo_atoms.def")
        return self.info[prio]

    def __getitem__(self, atom) -> PseudoAtomInfo:
        return self.get_strict(get_atom_name(atom))


def parse_pseudoatoms_RASPA(file) -> PseudoAtomListing:
    with open(file) as f:
        lines = [l for l in f.read().splitlines() if l.strip() and not l.startswith("#")]
    n = int(lines.pop(0))
    if n != len(lines):
        raise ValueError(f"Found {len(lines)} non-empty lines but {n} declared")
    exact, open_ended, info = {}, [], []
    for i, _l in enumerate(lines):
        l = _l.split()
        if len(l) != 14:
            raise ValueError(f'malformed line "{_l}" does not contain 14 fields.')
        is_open = l[0].endswith("_")
        name = l[0][:-1] if is_open else l[0]
        if is_open:
            open_ended.append((name, i))
        else:
            exact[name] = i
        info.append(PseudoAtomInfo(name, l[2] if l[1] == "yes" else "", l[3], float(l[4]), float(l[5]),
                                   float(l[6]), float(l[7]), float(l[8]), float(l[9]), float(l[10]),
                                   float(l[11]), l[12] == "absolute", int(l[13])))
    return PseudoAtomListing(exact, open_ended, info)


# --------------------------------------------------------------------------------------------
@dataclass
class RASPASystem:
    """raspa.jl:222-230 — plain arrays in Å / e.  ``mat`` columns are the cell vectors."""
    mat: np.ndarray                # 3x3 (np.inf diagonal for a free molecule)
    position: np.ndarray           # (n, 3) Cartesian
    atomic_symbol: list
    atomic_mass: np.ndarray
    atomic_charge: np.ndarray
    ismolecule: bool = False

    def __len__(self):
        return len(self.atomic_symbol)

    def with_positions(self, positions) -> "RASPASystem":
        """ChangePositionSystem (ChangePositionSystems.jl)."""
        return RASPASystem(self.mat, np.asarray(positions, dtype=np.float64).reshape(-1, 3), list(self.atomic_symbol),
                           self.atomic_mass, self.atomic_charge, self.ismolecule)


def read_cif_P1(path):
    """Minimal P1 CIF reader: cell parameters and the atom_site loop (label, fract_x/y/z).

    Stands in for `load_system(AtomsIO.ChemfilesParser(), cif)` (raspa.jl:306): cell matrix
    a=(a,0,0), b=(b cos g, b sin g, 0), c=(...), Cartesian = mat * fractional, no wrapping.
    """
    params = {}
    labels, fracs = [], []
    with open(path) as f:
        lines = f.read().splitlines()
    i = 0
    while i < len(lines):
        l = lines[i].strip()
        if l.startswith("_cell_length_") or l.startswith("_cell_angle_"):
            key, val = l.split()[:2]
            params[key] = float(val.split("(")[0])
            i += 1
        elif l == "loop_":
            i += 1
            heads = []
            while i < len(lines) and lines[i].strip().startswith("_"):
                heads.append(lines[i].strip().split()[0])
                i += 1
            if "_atom_site_fract_x" in heads:
                il = heads.index("_atom_site_label")
                ix, iy, iz = (heads.index("_atom_site_fract_" + c) for c in "xyz")
                while i < len(lines):
                    s = lines[i].split()
                    if len(s) < len(heads) or s[0].startswith("_") or s[0] in ("loop_",) or s[0].startswith("#"):
                        break
                    labels.append(s[il])
                    fracs.append([float(s[ix].split("(")[0]), float(s[iy].split("(")[0]), float(s[iz].split("(")[0])])
                    i += 1
            else:
                while i < len(lines) and lines[i].strip() and not lines[i].strip().startswith(("loop_", "_", "#")):
                    i += 1
        else:
            i += 1
    mat = mat_from_parameters((params["_cell_length_a"], params["_cell_length_b"], params["_cell_length_c"]),
                              (params["_cell_angle_alpha"], params["_cell_angle_beta"], params["_cell_angle_gamma"]))
    fr = np.array(fracs, dtype=np.float64).reshape(-1, 3)
    pos = fr @ mat.T
    return mat, labels, pos


def _ffname(pff) -> str:
    if isinstance(pff, str):
        return pff
    if isinstance(pff, ForceField):
        return pff.name
    return pff[0]


def hash_forcefield(ff: ForceField) -> str:
    """Stable content hash of a ForceField (rules, shifts, cutoff, species order): the role hash_string(forcefield)
    plays in raspa.jl:196-204 (the reference hashes Julia's serialisation; any stable digest serves the cache name)."""
    import hashlib
    h = hashlib.sha1()
    h.update(repr((sorted(ff.sdict.items(), key=lambda kv: kv[1]), float(ff.cutoff))).encode())
    for row in ff.interactions:
        for rule in row:
            parts = rule.rules if isinstance(rule, InteractionRuleSum) else (rule,)
            h.update(repr([(int(r.kind), tuple(float(x).hex() for x in r.params), float(r.shift).hex(), bool(r.tailcorrection))
                           for r in parts]).encode())
    return h.hexdigest()[:16]


def _fffullname(pff, forcefield: ForceField) -> str:
    """raspa.jl:196-204: the stock name for a force field given by name, '<name>-<hash>' for a customised ForceField
    or (name, ForceField / PseudoAtomListing) tuple, so that its grids never collide with the stock ones."""
    if isinstance(pff, str):
        return pff
    return f"{_ffname(pff)}-{hash_forcefield(forcefield)}"


def _ffpal(pff) -> PseudoAtomListing:
    if isinstance(pff, tuple) and isinstance(pff[1], PseudoAtomListing):
        return pff[1]
    return parse_pseudoatoms_RASPA(os.path.join(getdir_RASPA(), "forcefield", _ffname(pff), "pseudo_atoms.def"))


def _ff(pff, cutoff=None, ewald_precision=None) -> ForceField:
    if isinstance(pff, ForceField):
        if cutoff is not None and pff.cutoff != cutoff:
            raise ValueError(f"Inconsistent cutoffs: choose between {pff.cutoff} from force field or given {cutoff}")
        return pff
    if not isinstance(pff, str) and isinstance(pff[1], ForceField):
        return pff[1]
    kwargs = {}
    if cutoff is not None:
        kwargs["cutoff"] = cutoff
    if ewald_precision is not None:
        kwargs["ewald_precision"] = ewald_precision
    return parse_forcefield_RASPA(pff, **kwargs)


def _ffcharge(pff, atom, pal, pseudo) -> float:
    """raspa.jl:177-189."""
    if isinstance(pff, str) or (not isinstance(pff, ForceField) and isinstance(pff[1], PseudoAtomListing)):
        return pseudo.charge
    ff = _ff(pff)
    inter = ff[pseudo.type, pseudo.type]
    for r in FF._rules_of(inter):
        if r.kind is Kind.CoulombEwaldDirect:
            return r.params[1]
    return pseudo.charge


def load_framework_RASPA(name, pff) -> RASPASystem:
    """raspa.jl:303-326."""
    if not isinstance(name, str):
        mat = np.array(name, dtype=np.float64).reshape(3, 3)
        return RASPASystem(mat, np.zeros((0, 3)), [], np.zeros(0), np.zeros(0), False)
    cif = os.path.join(getdir_RASPA(), "structures", "cif", name if name.endswith(".cif") else name + ".cif")
    mat, labels, pos = read_cif_P1(cif)
    pal = _ffpal(pff)
    mass = np.empty(len(labels))
    charges = np.empty(len(labels))
    for i, lab in enumerate(labels):
        symb = get_atom_name(lab)
        pseudo = pal[symb]
        mass[i] = pseudo.mass
        charges[i] = _ffcharge(pff, symb, pal, pseudo)
    return RASPASystem(mat, pos, labels, mass, charges, False)


def parse_molecule_RASPA(file):
    with open(file) as f:
        lines = [l for l in f.read().splitlines() if l.strip() and not l.startswith("#")]
    num_atoms = int(lines[3])
    num_groups = int(lines[4])
    assert num_groups == 1
    assert num_atoms == int(lines[6])
    assert num_atoms == 1 or lines[5].strip() == "rigid"
    symbols, positions = [], np.zeros((num_atoms, 3))
    for i in range(num_atoms):
        l = lines[7 + i].split()
        symbols.append(l[1])
        if num_atoms != 1:
            positions[i] = [float(l[2]), float(l[3]), float(l[4])]
    return symbols, positions


_INFMAT = np.diag([np.inf, np.inf, np.inf])


def load_molecule_RASPA(name: str, ffname_molecule: str, pff, framework_system: RASPASystem | None = None) -> RASPASystem:
    """raspa.jl:344-366."""
    path = os.path.join(getdir_RASPA(), "molecules", ffname_molecule, name if name.endswith(".def") else name + ".def")
    symbols, positions = parse_molecule_RASPA(path)
    pal = _ffpal(pff)
    mass = np.array([pal[a].mass for a in symbols])
    charges = np.array([pal[a].charge for a in symbols])
    mat = framework_system.mat if framework_system is not None else _INFMAT
    return RASPASystem(mat, positions, symbols, mass, charges, True)


def default_system(s, pff) -> RASPASystem:
    """raspa.jl:368-382 — a single atom defined in the framework force field."""
    if isinstance(s, RASPASystem):
        return s
    forcefield = _ff(pff)
    symbol = get_atom_name(s)
    if symbol not in forcefield.sdict:
        return load_molecule_RASPA(s, "Default", forcefield)
    pseudo = _ffpal(pff)[symbol]
    return RASPASystem(_INFMAT, np.zeros((1, 3)), [symbol], np.array([pseudo.mass]), np.array([pseudo.charge]), True)


def decide_parse_block(blockfile, molecule: RASPASystem, framework_name) -> str:
    """raspa.jl:381-392."""
    if not isinstance(framework_name, str):
        return ""
    if isinstance(blockfile, (bool, str)):
        new = blockfile
    else:
        new = not (len(molecule) == 1 and molecule.atomic_charge[0] > np.finfo(float).eps)
    if isinstance(new, bool):
        if not new:
            return ""
        return framework_name.split("_")[0]
    return str(new)


def parse_block(blockfile, framework_name, framework: RASPASystem, molecule, spacing):
    """raspa.jl:394-400."""
    from .coordinates import BlockFile, GridCoordinatesSetup, parse_blockfile
    blockpath = decide_parse_block(blockfile, molecule, framework_name)
    csetup = GridCoordinatesSetup.from_cell(framework.mat, spacing)
    if not blockpath:
        return BlockFile(csetup)
    return parse_blockfile(os.path.join(getdir_RASPA(), "structures", "block", blockpath) + ".block", csetup)


def grid_locations(framework, pff, forcefield: ForceField, atoms, gridstep: float, supercell, root: str | None = None):
    """raspa.jl:403-418 — `.grid` cache paths: grids/<ff>/<framework>/<step %.6f>/[<AxBxC>/]…"""
    if not isinstance(framework, str):
        return "", []
    rawname = _ffname(pff)          # locates force_field_mixing_rules.def
    raspa = getdir_RASPA()
    rootdir = os.path.join(root if root is not None else os.path.join(raspa, "grids"), _fffullname(pff, forcefield), framework)
    grid_dir = os.path.join(rootdir, "%.6f" % gridstep)
    supercell_name = "x".join(str(int(x)) for x in supercell)
    coulomb_grid_path = os.path.join(grid_dir, supercell_name, framework + "_Electrostatics_Ewald.grid")
    with open(os.path.join(raspa, "forcefield", rawname, "force_field_mixing_rules.def")) as f:
        trunc_or_shift = next(l for l in f.read().splitlines() if not l.startswith("#")).strip()
    vdws = [os.path.join(grid_dir, f"{framework}_{atom}_{trunc_or_shift}.grid") for atom in atoms]
    return coulomb_grid_path, vdws


# --------------------------------------------------------------------------------------------
def parse_interaction_RASPA(l: str, mixingrules: bool, shift: bool, cutoff: float, tailcorrection: bool):
    """raspa.jl:533-554."""
    splits = l.split()
    for i, x in enumerate(splits):
        if x[0] == "#" or x.startswith("//"):
            splits = splits[:i]
            break
    m = 1 if mixingrules else 0
    kind = splits[2 - m].lower()
    key = (splits[0], splits[1 - m])
    if kind == "lennard-jones":
        rule = FF.make_rule(Kind.LennardJones, [float(x) for x in splits[3 - m:]], shift, cutoff, tailcorrection)
    elif kind == "none":
        rule = FF.NoInteraction()
    elif kind == "buckingham":
        rule = FF.make_rule(Kind.Buckingham, [float(x) for x in splits[3 - m:]], shift, cutoff, tailcorrection)
    elif kind == "buckingham2":
        rule = InteractionRuleSum.make([
            FF.make_rule(Kind.Buckingham, [float(x) for x in splits[3 - m:6 - m]], shift, cutoff, tailcorrection),
            FF.make_rule(Kind.HardSphere, [float(splits[6 - m]), 0.0], shift, cutoff, tailcorrection)])
    else:
        raise ValueError(f"{kind} interaction potential not implemented")
    return key, rule


def parse_mixingrule_RASPA(l: str) -> Mixing:
    x = l.lower()
    if x == "lorentz-berthelot":
        return Mixing.LorentzBerthelot
    if x in ("jorgensen", "good-hope", "geometric"):
        return Mixing.Geometric
    raise ValueError(f"Unknown mixing rule {l}")


def _parse_shift(s):
    s2 = s.strip().lower()
    if s2 == "shifted":
        return True
    assert s2 == "truncated", s
    return False


def _parse_yesno(s):
    s2 = s.strip().lower()
    if s2 == "yes":
        return True
    assert s2 == "no", s
    return False


def _next_noncomment_line(lines, i):
    j = i + 1
    while lines[j].startswith("#"):
        j += 1
    return j


def ewald_alpha_forcefield(ewald_precision: float, cutoff: float) -> float:
    """raspa.jl:682-687 — the α stored in CoulombEwaldDirect rules (not the α of the grid,
    which comes from initialize_ewald)."""
    if ewald_precision == 0:
        return 0.0
    eps = math.log(float(ewald_precision * cutoff))
    return math.sqrt(abs(eps + math.log(math.sqrt(abs(eps))))) / float(cutoff)


def parse_forcefield_RASPA(name, cutoff: float = 12.0, ewald_precision: float | None = None) -> ForceField:
    """raspa.jl:611-702."""
    if ewald_precision is None:
        ewald_precision = 0.0 if math.isinf(cutoff) else 1e-6
    pseudoatoms = name[1] if (not isinstance(name, str) and isinstance(name[1], PseudoAtomListing)) else _ffpal(name)
    rawname = _ffname(name)
    inputs = []
    general_mixing = Mixing.ErrorOnMix
    shift, tailcorrection = True, False
    sdict = {}
    ffdir = os.path.join(getdir_RASPA(), "forcefield", rawname)
    mr = os.path.join(ffdir, "force_field_mixing_rules.def")
    if os.path.isfile(mr):
        with open(mr) as f:
            lines = f.read().splitlines()
        idx = _next_noncomment_line(lines, 0)
        shift = _parse_shift(lines[idx])
        idx = _next_noncomment_line(lines, idx)
        tailcorrection = _parse_yesno(lines[idx])
        idx = _next_noncomment_line(lines, idx)
        num = int(lines[idx])
        for i in range(num):
            idx = _next_noncomment_line(lines, idx)
            key, rule = parse_interaction_RASPA(lines[idx], True, shift, cutoff, tailcorrection)
            sdict[key[0]] = i
            inputs.append((key, rule))
        idx = _next_noncomment_line(lines, idx)
        general_mixing = parse_mixingrule_RASPA(lines[idx].strip())
    ff = ForceField.build(inputs, general_mixing, cutoff, shift, tailcorrection, sdict=sdict, name=rawname)
    inter = ff.interactions
    ffdef = os.path.join(ffdir, "force_field.def")
    if os.path.isfile(ffdef):
        with open(ffdef) as f:
            lines = f.read().splitlines()
        jdx = _next_noncomment_line(lines, 0)
        numnewrules = int(lines[jdx])
        startnewrules = jdx
        jdx = _next_noncomment_line(lines, jdx)
        for _ in range(numnewrules):
            jdx = _next_noncomment_line(lines, jdx)
        numnewinter = int(lines[jdx])

        def idxs(x, y):
            for a in (x, y):
                if a not in sdict:
                    raise KeyError(f"Atom {a} absent from force_field_mixing_rules.def")
            return sdict[x], sdict[y]

        for _ in range(numnewinter):
            jdx = _next_noncomment_line(lines, jdx)
            (x1, y1), newinter = parse_interaction_RASPA(lines[jdx], False, shift, cutoff, tailcorrection)
            i1, j1 = idxs(x1, y1)
            inter[i1][j1] = inter[j1][i1] = newinter
        jdx = _next_noncomment_line(lines, jdx)
        numnewmixing = int(lines[jdx])
        for _ in range(numnewmixing):
            jdx = _next_noncomment_line(lines, jdx)
            x2, y2, _newmix = lines[jdx].split()
            i2, j2 = idxs(x2, y2)
            inter[i2][j2] = inter[j2][i2] = FF.mix_rules(inter[i2][i2], inter[j2][j2], parse_mixingrule_RASPA(_newmix))
        jdx = startnewrules
        for _ in range(numnewrules):
            jdx = _next_noncomment_line(lines, jdx)
            x3, y3, _newshift, _newtail = lines[jdx].split()
            newshift, newtail = _parse_shift(_newshift), _parse_yesno(_newtail)
            i3, j3 = idxs(x3, y3)
            inter[i3][j3] = inter[j3][i3] = FF.map_rule(
                lambda r: FF.make_rule(r.kind, r.params, newshift, cutoff, newtail), inter[i3][j3])
    alpha = ewald_alpha_forcefield(ewald_precision, cutoff)
    for ati, i4 in sdict.items():
        qi = pseudoatoms[ati].charge
        if qi == 0.0:
            continue
        for atj, j4 in sdict.items():
            qj = pseudoatoms[atj].charge
            if qj == 0.0:
                continue
            extra = (InteractionRule(Kind.Coulomb, (qi, qj), 0.0, False) if ewald_precision == 0
                     else InteractionRule(Kind.CoulombEwaldDirect, (alpha, qi, qj), 0.0, False))
            inter[i4][j4] = FF.sum_rules(inter[i4][j4], extra)
    return ForceField(inter, ff.sdict, ff.symbols, ff.cutoff, rawname)


# --------------------------------------------------------------------------------------------
def retrieve_or_create_grid(grid_path, syst_framework, forcefield, gridstep, atom_or_eframework, mat, new, cutoff,
                            devices=None):
    """raspa.jl:420-439."""
    from .ewald import EwaldFramework
    from .grids import EnergyGrid, create_grid_coulomb, create_grid_vdw, parse_grid
    if not grid_path or math.isinf(cutoff):
        return EnergyGrid.invalid()
    if cutoff != 12.0:
        raise ValueError("Cutoff other than 12 Å or infinity is not supported.")
    iscoulomb = isinstance(atom_or_eframework, EwaldFramework)
    if not iscoulomb and not FF.needsvdwgrid(forcefield, atom_or_eframework):
        return EnergyGrid.zero()
    if new or not os.path.isfile(grid_path):
        os.makedirs(os.path.dirname(grid_path), exist_ok=True)
        if iscoulomb:
            create_grid_coulomb(grid_path, syst_framework, forcefield, gridstep, atom_or_eframework, devices=devices)
        else:
            create_grid_vdw(grid_path, syst_framework, forcefield, gridstep, atom_or_eframework, devices=devices)
    return parse_grid(grid_path, iscoulomb, mat)


def _create_missing_grids_fused(syst_framework, forcefield, gridstep, cutoff, new, devices, coulomb_path, ewald,
                                vdw_atoms_paths, keep_on_device: bool = True):
    """Create the `.grid` files of one setup that do not exist yet (or all of them if `new`), fusing
    up to four VdW probes and the Coulomb grid per pass (grids.jl:137-185 writes them one by one).

    keep_on_device: the reference writes each file and parses it straight back (raspa.jl:430-438: disk -> host ->
    scale -> EnergyGrid).  Here the freshly built grids stay in HBM: the build output is gathered (several GPUs) and
    turned into interpolation handles on the device (ceg_grid_upload_dev applies parse_grid's conversion,
    grids.jl:78), and the files are written from one downloaded copy by a background thread.
    -> {path: EnergyGrid} of the grids created (device-resident), empty with keep_on_device=False."""
    import threading
    from .grids import build_grids, build_grids_device, write_grid
    if cutoff != 12.0:
        raise ValueError("Cutoff other than 12 Å or infinity is not supported.")
    todo_vdw = [(a, p) for a, p in vdw_atoms_paths if p and (new or not os.path.isfile(p))]
    todo_coul = bool(coulomb_path) and (new or not os.path.isfile(coulomb_path))
    made = {}
    if not todo_vdw and not todo_coul:
        return made
    num_vdw = find_supercell(syst_framework.mat, forcefield.cutoff)
    num_coul = find_supercell(syst_framework.mat, CUTOFF_COULOMB)
    first = True
    while todo_vdw or (todo_coul and first):
        batch, todo_vdw = todo_vdw[:4], todo_vdw[4:]
        with_coul = todo_coul and first
        atoms = tuple(a for a, _ in batch)
        alpha = ewald.alpha if with_coul else 0.0
        paths = [p for _, p in batch] + ([coulomb_path] if with_coul else [])
        for p in paths:
            os.makedirs(os.path.dirname(p), exist_ok=True)
        if keep_on_device:
            cset, egs, host = build_grids_device(syst_framework, forcefield, gridstep, atoms=atoms, coulomb=with_coul, alpha=alpha,
                                                 devices=devices, ewald_precision=ewald.precision if with_coul else 1e-6,
                                                 download=True)

            def finish(eg, arr, path, nuc, prec):
                write_grid(path, cset, nuc, arr, ewald_precision=prec)
                eg._grid = (arr.astype(np.float64) * GRID_TO_KELVIN).astype(np.float32)     # grids.jl:78

            for k, (eg, arr, path) in enumerate(zip(egs, host, paths)):
                is_coul = with_coul and k == len(egs) - 1
                t = threading.Thread(target=finish, args=(eg, arr, path, num_coul if is_coul else num_vdw,
                                                          ewald.precision if is_coul else None))
                eg._pending = t
                t.start()
                made[path] = eg
        else:
            cset, grids, _ = build_grids(syst_framework, forcefield, gridstep, atoms=atoms, coulomb=with_coul, alpha=alpha,
                                         devices=devices)
            for (a, p), g in zip(batch, grids):
                write_grid(p, cset, num_vdw, g)
            if with_coul:
                write_grid(coulomb_path, cset, num_coul, grids[-1], ewald_precision=ewald.precision)
        first = False
    return made


def setup_RASPA(framework, pff, molecule, ffname_molecule: str | None = None, *, gridstep: float = 0.15,
                supercell=None, blockfile=None, new: bool = False, cutoff: float = 12.0, grids_root: str | None = None,
                devices=None, keep_on_device: bool = True):
    """raspa.jl:472-530 — returns a CrystalEnergySetup (grids.jl:284-294).

    Grids that have to be created are handed from the construction kernels to the interpolator on the device
    (keep_on_device, see _create_missing_grids_fused); their `.grid` files are written in the background, in the
    reference's byte layout.  Cached files are parsed as in the reference."""
    from .ewald import EwaldFramework, initialize_ewald
    from .grids import CrystalEnergySetup, EnergyGrid
    syst_framework = load_framework_RASPA(framework, pff)
    if ffname_molecule is not None:
        syst_mol = load_molecule_RASPA(molecule, ffname_molecule, pff, syst_framework)
    else:
        syst_mol = default_system(molecule, pff)
    if supercell is None:
        supercell = (1, 1, 1) if math.isinf(cutoff) else find_supercell(syst_framework.mat, cutoff)
    supercell = tuple(int(x) for x in supercell)
    mat = syst_framework.mat
    forcefield = _ff(pff, cutoff=cutoff)
    block = parse_block(blockfile, framework, syst_framework, syst_mol, gridstep)

    atomdict = {}
    atoms = list(syst_mol.atomic_symbol)
    for atom in atoms:
        atomdict.setdefault(atom, len(atomdict))
    atomsidx = [atomdict[a] for a in atoms]
    rev_atomdict = [None] * len(atomdict)
    for at, i in atomdict.items():
        rev_atomdict[i] = at
    coulomb_grid_path, vdws = grid_locations(framework, pff, forcefield, rev_atomdict, gridstep, supercell, grids_root)

    needcoulomb = bool(np.any(syst_mol.atomic_charge != 0))
    ewald = initialize_ewald(syst_framework, supercell) if needcoulomb else None
    if isinstance(framework, str) and not math.isinf(cutoff):
        # Every missing grid of this setup is built in fused passes over the framework atoms (up to four
        # VdW probes + the Coulomb grid share one distance computation), instead of one pass per grid as
        # the loop below would do through retrieve_or_create_grid; the files written are the same.
        made = _create_missing_grids_fused(syst_framework, forcefield, gridstep, cutoff, new, devices,
                                           coulomb_grid_path if needcoulomb else None, ewald,
                                           [(a, p) for a, p in zip(rev_atomdict, vdws) if FF.needsvdwgrid(forcefield, a)],
                                           keep_on_device)
        new = False
    else:
        made = {}
    if needcoulomb:
        if not isinstance(framework, str):
            coulomb = EnergyGrid.zero()
        elif coulomb_grid_path in made:
            coulomb = made[coulomb_grid_path]
        else:
            coulomb = retrieve_or_create_grid(coulomb_grid_path, syst_framework, forcefield, gridstep, ewald, mat, new,
                                              cutoff, devices)
    else:
        coulomb, ewald = EnergyGrid.zero(), EwaldFramework.empty(mat)
    grids = []
    for i, atom in enumerate(rev_atomdict):
        if not isinstance(framework, str):
            grids.append(EnergyGrid.zero())
        elif vdws[i] in made:
            grids.append(made[vdws[i]])
        else:
            grids.append(retrieve_or_create_grid(vdws[i], syst_framework, forcefield, gridstep, atom, mat, new, cutoff,
                                                 devices))
    charges = np.array(syst_mol.atomic_charge, dtype=np.float64)
    return CrystalEnergySetup(syst_framework, syst_mol, coulomb, charges, grids, atomsidx, ewald, forcefield, block)
