"""ProbeSystem: framework replicated to the 24 Å-wide supercell + the device handle that
evaluates VdW / Ewald real-space derivatives on it.

Mirrors /root/reference/src/probes.jl:
  ProbeSystem ctor                      probes.jl:21-63
  compute_derivatives_vdw               probes.jl:71-92    -> ceg_compute_derivatives
  compute_derivatives_ewald             probes.jl:94-117   -> ceg_compute_derivatives

Unlike the reference (one ProbeSystem per grid, one pass over the atoms per grid) a handle can
fuse several VdW probes and the Coulomb grid into ONE pass over the atoms.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib
from .cell import find_supercell, flat, inv3, prepare_periodic_distance_computations
from .constants import CUTOFF_COULOMB
from .forcefields import ForceField, get_atom_name, pair_table


def replicate_framework(positions: np.ndarray, mat: np.ndarray, supercell) -> np.ndarray:
    """probes.jl:29-55 — image (ix,iy,iz) of atom i sits at index ix*n*ny*nz + iy*n*nz + iz*n + i
    and equals p_i + (ix*a + (iy*b + iz*c)) with that association."""
    nx, ny, nz = supercell
    n = positions.shape[0]
    if nx * ny * nz == 1:
        return positions.copy()
    wx, wy, wz = mat[:, 0], mat[:, 1], mat[:, 2]
    out = np.empty((nx * ny * nz * n, 3), dtype=np.float64)
    for iz in range(nz):
        stepz = iz * wz
        for iy in range(ny):
            stepyz = iy * wy + stepz
            for ix in range(nx):
                stepxyz = ix * wx + stepyz
                base = ix * (ny * nz * n) + iy * (nz * n) + iz * n
                if ix == iy == iz == 0:
                    out[base:base + n] = positions
                else:
                    out[base:base + n] = positions + stepxyz
    return out


@dataclass
class ProbeInputs:
    """Flattened inputs of a (fused) probe system, exactly what crosses the C ABI."""
    mat: np.ndarray            # supercell matrix (columns = vectors)
    invmat: np.ndarray
    ortho: bool
    safemin2: float
    positions: np.ndarray      # (N_sc, 3)
    types: np.ndarray          # int32 (N_sc,), 0-based force-field species index
    charges: np.ndarray | None
    n_types: int
    table: np.ndarray          # (n_probes, n_types) pair_dtype
    cutoff2: float
    want_coulomb: bool
    alpha: float
    supercell: tuple

    def cell_struct(self):
        return _lib.make_cell(flat(self.mat), flat(self.invmat), self.ortho, self.safemin2)


def probe_inputs(framework, forcefield: ForceField, atoms=(), coulomb: bool = False, alpha: float = 0.0) -> ProbeInputs:
    """probes.jl:21-63 for a list of probe species (`atoms`) and/or the Coulomb grid."""
    n = len(framework)
    kinds0 = np.array([forcefield.sdict[get_atom_name(s)] for s in framework.atomic_symbol], dtype=np.int32)
    supercell = find_supercell(framework.mat, CUTOFF_COULOMB)   # 12 Å hard-coded (probes.jl:24)
    num = supercell[0] * supercell[1] * supercell[2]
    mat = np.asarray(framework.mat, dtype=np.float64) * np.array(supercell, dtype=np.float64)[None, :]
    invmat = inv3(mat)
    positions = replicate_framework(np.asarray(framework.position, dtype=np.float64), framework.mat, supercell)
    types = np.tile(kinds0, num)
    charges = np.tile(np.asarray(framework.atomic_charge, dtype=np.float64), num) if coulomb else None
    table = pair_table(forcefield, list(atoms))
    present = np.unique(kinds0)
    for p in range(table.shape[0]):
        if (table[p, present]["flags"] < 0).any():
            from .forcefields import UndefinedInteractionError
            raise UndefinedInteractionError()
    ortho, safemin = prepare_periodic_distance_computations(mat)
    return ProbeInputs(mat, invmat, ortho, safemin * safemin, positions, types, charges, len(forcefield.sdict), table,
                       float(forcefield.cutoff) ** 2, bool(coulomb), float(alpha), supercell)


class ProbeSystem:
    """Device-resident probe system (probes.jl:12-20)."""

    def __init__(self, framework, forcefield: ForceField, atoms=(), coulomb: bool = False, alpha: float = 0.0,
                 device: int = 0, inputs: ProbeInputs | None = None):
        if isinstance(atoms, str):
            atoms = (atoms,)
        self.inputs = inputs if inputs is not None else probe_inputs(framework, forcefield, atoms, coulomb, alpha)
        i = self.inputs
        self.device = device
        self.handle = _lib.probe_create(device, i.cell_struct(), i.positions, i.types, i.charges, i.n_types, i.table,
                                        i.cutoff2, i.want_coulomb, i.alpha)
        self.n_grids = _lib.probe_num_grids(self.handle)

    def compute_derivatives(self, positions) -> np.ndarray:
        """(n_grids, n_pos, 8): value, d1[3], d2[3], d3 per grid (VdW probes first, Coulomb last)."""
        return _lib.compute_derivatives(self.handle, positions)

    def close(self):
        h, self.handle = self.handle, None
        if h is not None:
            _lib.probe_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compute_derivatives_vdw(s: ProbeSystem, pos):
    """probes.jl:71-92 — (value, d1, d2, d3) of the first VdW probe of `s` at `pos`."""
    d = s.compute_derivatives(np.asarray(pos, dtype=np.float64).reshape(1, 3))[0, 0]
    return d[0], d[1:4].copy(), d[4:7].copy(), d[7]


def compute_derivatives_ewald(s: ProbeSystem, pos):
    """probes.jl:94-117."""
    d = s.compute_derivatives(np.asarray(pos, dtype=np.float64).reshape(1, 3))[-1, 0]
    return d[0], d[1:4].copy(), d[4:7].copy(), d[7]
