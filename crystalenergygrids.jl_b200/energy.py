"""Guest–guest pair energies for batched Monte-Carlo trial evaluation.

Mirrors /root/reference/src/energy.jl:
  compute_vdw / compute_vdw_noneighbour                 energy.jl:339-385
  single_contribution_vdw(_noneighbour)                 energy.jl:397-427
with the pair rules evaluated as `rule(d²)` (interactions.jl:367-409), i.e. including the real-space
Ewald term CoulombEwaldDirect that parse_forcefield_RASPA adds to charged pairs (raspa.jl:688-700).
The sums run on the device (ceg_guest_energy, csrc/guest.cuh); rigid molecules only (the reference's
energy_intra of flexible molecules requires an infinite cutoff, energy.jl:309-326).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .cell import flat, inv3, prepare_periodic_distance_computations
from .constants import COULOMBIC_CONVERSION_FACTOR
from .forcefields import (PAIR_BUCK, PAIR_HS, PAIR_LJ, ForceField, Kind, UndefinedInteractionError, _rules_of,
                          get_atom_name)

PAIR_EWALD = 8
PAIR_COULOMB = 16

pair_energy_dtype = np.dtype([
    ("flags", np.int32), ("_pad", np.int32), ("hs_r", np.float64),
    ("lj_eps", np.float64), ("lj_sigma", np.float64), ("lj_shift", np.float64),
    ("buck_A", np.float64), ("buck_B", np.float64), ("buck_C", np.float64), ("buck_shift", np.float64),
    ("ewald_alpha", np.float64), ("ewald_qq", np.float64), ("coulomb_qq", np.float64),
], align=True)
assert pair_energy_dtype.itemsize == 96


def flatten_pair_energy(rule) -> np.void:
    """One ForceField entry as the energy call rule(d²) sees it (interactions.jl:367-409)."""
    out = np.zeros((), dtype=pair_energy_dtype)
    for r in _rules_of(rule):
        k = r.kind
        if k is Kind.NoInteraction:
            continue
        if k is Kind.HardSphere:
            out["flags"] |= PAIR_HS
            out["hs_r"] = r.params[0] + r.params[1]
        elif k is Kind.LennardJones:
            out["flags"] |= PAIR_LJ
            out["lj_eps"], out["lj_sigma"] = r.params
            out["lj_shift"] = r.shift
        elif k is Kind.Buckingham:
            out["flags"] |= PAIR_BUCK
            out["buck_A"], out["buck_B"], out["buck_C"] = r.params
            out["buck_shift"] = r.shift
        elif k is Kind.CoulombEwaldDirect:
            out["flags"] |= PAIR_EWALD
            out["ewald_alpha"] = r.params[0]
            out["ewald_qq"] = COULOMBIC_CONVERSION_FACTOR * r.params[1] * r.params[2]
        elif k is Kind.Coulomb:
            out["flags"] |= PAIR_COULOMB
            out["coulomb_qq"] = COULOMBIC_CONVERSION_FACTOR * r.params[0] * r.params[1]
        elif k is Kind.UndefinedInteraction:
            raise UndefinedInteractionError()
        else:
            raise ValueError(f"guest pair energy not implemented for {k.name}")
    return out


def pair_energy_table(ff: ForceField, species) -> np.ndarray:
    """(n, n) table over the listed guest species names."""
    idx = [ff.sdict[get_atom_name(s)] for s in species]
    tab = np.zeros((len(idx), len(idx)), dtype=pair_energy_dtype)
    for a, i in enumerate(idx):
        for b, j in enumerate(idx):
            tab[a, b] = flatten_pair_energy(ff.interactions[i][j])
    return tab


@dataclass
class GuestEnvironment:
    """The guest atoms a trial molecule interacts with (SimulationStep.positions/atoms, energy.jl:43-60)."""
    mat: np.ndarray               # simulation cell (columns = vectors)
    cutoff: float
    table: np.ndarray             # pair_energy_dtype (n_species, n_species)
    positions: np.ndarray         # (n_env, 3)
    species: np.ndarray           # int32 (n_env,), index into table
    molecule: np.ndarray          # int32 (n_env,), molecule id of each atom
    device: int = 0

    def cell_struct(self):
        m = np.asarray(self.mat, dtype=np.float64)
        ortho, safemin = prepare_periodic_distance_computations(m)
        return _lib.make_cell(flat(m), flat(inv3(m)), ortho, safemin * safemin)

    # device-resident copy (ceg_guest_env_create), made on first use; never copied by dataclasses.replace
    _handle: object = field(default=None, init=False, repr=False, compare=False)

    @property
    def handle(self):
        if self._handle is None:
            self._handle = _lib.guest_env_create(self.device, self.cell_struct(), float(self.cutoff) ** 2, self.table,
                                                 self.positions, self.species, self.molecule)
        return self._handle

    def move(self, first: int, positions) -> None:
        """Accepted move: atoms [first, first + len(positions)) take new positions, on the host copy and on the device."""
        pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
        self.positions = np.array(self.positions, dtype=np.float64)
        self.positions[first:first + len(pos)] = pos
        if self._handle is not None:
            _lib.guest_env_update(self._handle, first, pos)

    def close(self):
        h, self._handle = self._handle, None
        if h is not None:
            _lib.guest_env_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def single_contribution_vdw(env: GuestEnvironment, mol_species, trial_positions, skip_molecule: int = -1) -> np.ndarray:
    """energy.jl:407-427, batched: energy of a rigid molecule (species indices `mol_species`) at each trial
    placement (n_trials, n_atoms, 3) with every environment atom not belonging to `skip_molecule`.  The environment
    stays on the device between calls (GuestEnvironment.handle / .move)."""
    return _lib.guest_env_energy(env.handle, mol_species, skip_molecule, trial_positions)


def compute_vdw(env: GuestEnvironment) -> float:
    """energy.jl:339-385 for rigid molecules: sum over pairs of atoms of different molecules.  Evaluated as
    half the sum of the single contributions of every molecule (same pairs, different summation order)."""
    total = 0.0
    for m in np.unique(env.molecule):
        sel = env.molecule == m
        total += float(single_contribution_vdw(env, env.species[sel], env.positions[sel][None], int(m))[0])
    return 0.5 * total


def movement_energies(setup, env: GuestEnvironment, ewald_rest, mol_species, trial_positions, skip_molecule: int = -1):
    """movement_energy (montecarlo.jl:563-606) for a batch of trial placements of one rigid molecule:
    returns (framework_vdw, framework_coulomb_direct, ewald_reciprocal, guest_guest), each (n_trials,).

      framework_*      framework_interactions: interpolated grids            -> ceg_energy_points
      ewald_reciprocal single_contribution_ewald with `ewald_rest` = framework + other adsorbates
                       (ewald.jl:704-738)                                     -> ceg_ewald_single_contribution
      guest_guest      single_contribution_vdw (energy.jl:397-427)            -> ceg_guest_energy
    The Monte-Carlo driver that consumes them (acceptance, bookkeeping) is out of scope."""
    from .ewald import single_contribution_ewald
    from .grids import energy_points
    pos = np.asarray(trial_positions, dtype=np.float64).reshape(-1, len(setup.atomsidx), 3)
    fv, fc = energy_points(setup, pos, reciprocal=False)
    rec = single_contribution_ewald(setup.ewald, ewald_rest, setup.charges, pos)
    gg = single_contribution_vdw(env, mol_species, pos, skip_molecule)
    return fv, fc, rec, gg
