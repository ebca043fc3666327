"""The grid / Ewald part of the reference's Monte-Carlo setup, and the batched framework term of its moves.

Mirrors /root/reference/src/montecarlo.jl:
  setup_montecarlo(framework, pff, systems; ...)    montecarlo.jl:269-334 — the species' VdW grids, the Coulomb grid,
                                                     the EwaldFramework and the blocking masks of ALL molecule kinds
  framework_interactions(mc, i, positions)          montecarlo.jl:490-520

The Markov chain itself (SimulationStep, MCMoves, acceptance, IncrementalEwaldContext bookkeeping) is serial host code
and stays in the reference (out of scope); what it needs from this path is gathered in MonteCarloGrids.
Grids that have to be created are built in fused passes and handed to the interpolator on the device
(raspa._create_missing_grids_fused), never re-read from disk.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import forcefields as FF
from .cell import find_supercell
from .coordinates import BlockFile, GridCoordinatesSetup
from .grids import CrystalEnergySetup, EnergyGrid, energy_points


@dataclass
class MonteCarloGrids:
    cell: np.ndarray                 # supercell matrix (montecarlo.jl:287)
    csetup: GridCoordinatesSetup
    forcefield: object
    eframework: object               # EwaldFramework
    coulomb: EnergyGrid
    grids: list                      # indexed by ff.sdict index; None for species no molecule contains
    blocks: list                     # BlockFile per system
    num_framework_atoms: np.ndarray  # per ff.sdict index, over the supercell
    systems: list                    # RASPASystem per molecule kind
    ffidx: list                      # per system: ff.sdict index of each atom
    charges: np.ndarray              # per ff.sdict index (e)
    _setups: dict = field(default_factory=dict)

    def setup_of(self, i: int) -> CrystalEnergySetup:
        """The CrystalEnergySetup view of molecule kind i (what energy_point evaluates)."""
        if i not in self._setups:
            idx = self.ffidx[i]
            uniq = sorted(set(idx))
            grids = [self.grids[k] if self.grids else EnergyGrid.zero() for k in uniq]
            atomsidx = [uniq.index(k) for k in idx]
            self._setups[i] = CrystalEnergySetup(None, self.systems[i], self.coulomb, np.array([self.charges[k] for k in idx]),
                                                 grids, atomsidx, self.eframework, self.forcefield, BlockFile(self.csetup))
        return self._setups[i]


def setup_montecarlo(framework, pff, systems, *, blockfiles=None, gridstep: float = 0.15, supercell=None, new: bool = False,
                     cutoff: float = 12.0, grids_root: str | None = None, devices=None, keep_on_device: bool = True):
    """montecarlo.jl:269-334, up to (not including) the SimulationStep: `systems` is a list of molecule names or
    RASPASystem objects (a (system, count) tuple is accepted and its count ignored)."""
    from . import raspa as R
    from .ewald import EwaldFramework, initialize_ewald
    if isinstance(framework, str) and framework.endswith(".cif"):
        framework = framework[:-4]
    syst_framework = R.load_framework_RASPA(framework, pff) if isinstance(framework, str) else framework
    ff = R._ff(pff, cutoff=cutoff)
    mat = syst_framework.mat
    if supercell is None:
        supercell = find_supercell(mat, cutoff)
    supercell = tuple(int(x) for x in supercell)
    cell = np.asarray(mat, dtype=np.float64) * np.asarray(supercell, dtype=np.float64)[None, :]
    csetup = GridCoordinatesSetup.from_cell(mat, gridstep)
    blockfiles = [None] * len(systems) if blockfiles is None else list(blockfiles)
    if len(blockfiles) != len(systems):
        raise ValueError("Please provide one blockfiles element per system")
    mols = [R.default_system(s[0] if isinstance(s, tuple) else s, pff) for s in systems]
    fname = framework if isinstance(framework, str) else ""
    blocks = [R.parse_block(b, fname, syst_framework, m, gridstep) if isinstance(framework, str) else BlockFile(csetup)
              for b, m in zip(blockfiles, mols)]
    needcoulomb = any(bool(np.any(np.asarray(m.atomic_charge) != 0)) for m in mols)
    encountered = {a for m in mols for a in m.atomic_symbol}
    atoms = sorted(((a, ff.sdict[a]) for a in encountered), key=lambda t: t[1])
    coulomb_grid_path, vdw_paths = R.grid_locations(framework, pff, ff, [a for a, _ in atoms], gridstep, supercell, grids_root)
    eframework = initialize_ewald(syst_framework, supercell) if needcoulomb else EwaldFramework.empty(mat)
    made = {}
    if isinstance(framework, str) and not math.isinf(cutoff):
        made = R._create_missing_grids_fused(syst_framework, ff, gridstep, cutoff, new, devices,
                                             coulomb_grid_path if needcoulomb else None, eframework,
                                             [(a, p) for (a, _), p in zip(atoms, vdw_paths) if FF.needsvdwgrid(ff, a)],
                                             keep_on_device)
        new = False
    if needcoulomb and isinstance(framework, str):
        coulomb = made.get(coulomb_grid_path) or R.retrieve_or_create_grid(coulomb_grid_path, syst_framework, ff, gridstep,
                                                                           eframework, mat, new, cutoff, devices)
    else:
        coulomb = EnergyGrid.zero()
    grids = []
    if vdw_paths and isinstance(framework, str):
        grids = [None] * len(ff.sdict)
        for (atom, i), p in zip(atoms, vdw_paths):
            grids[i] = made.get(p) or R.retrieve_or_create_grid(p, syst_framework, ff, gridstep, atom, mat, new, cutoff, devices)
    num_framework_atoms = np.zeros(len(ff.sdict), dtype=np.int64)
    for s in syst_framework.atomic_symbol:
        num_framework_atoms[ff.sdict[FF.get_atom_name(s)]] += int(np.prod(supercell))
    pal = R._ffpal(pff)

    def charge_of(sym):
        try:
            return float(pal[sym].charge)
        except KeyError:
            return 0.0
    charges = np.array([charge_of(s) for s in ff.symbols], dtype=np.float64)
    ffidx = [[ff.sdict[a] for a in m.atomic_symbol] for m in mols]
    return MonteCarloGrids(cell, csetup, ff, eframework, coulomb, grids, blocks, num_framework_atoms, mols, ffidx, charges)


def framework_interactions(mc: MonteCarloGrids, i: int, positions):
    """montecarlo.jl:490-520 for a BATCH of placements of molecule kind i: positions (n_pos, n_atoms, 3) (or one
    (n_atoms, 3) placement) -> (vdw_K, direct_K), the Van der Waals and real-space Ewald energies with the framework;
    the reciprocal part comes from compute_ewald / single_contribution_ewald."""
    if not mc.grids:
        p = np.asarray(positions, dtype=np.float64)
        z = np.zeros(1 if p.ndim == 2 else p.shape[0])
        return (0.0, 0.0) if p.ndim == 2 else (z, z.copy())
    p = np.asarray(positions, dtype=np.float64)
    v, c = energy_points(mc.setup_of(i), p.reshape(-1, len(mc.ffidx[i]), 3), reciprocal=False)
    return (float(v[0]), float(c[0])) if p.ndim == 2 else (v, c)
