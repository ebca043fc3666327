"""Grid coordinates and blocking masks (host side).

Mirrors /root/reference/src/coordinates.jl:
  GridCoordinatesSetup                  coordinates.jl:15-41
  wrap_atom / offsetpoint               coordinates.jl:58-66
  inverse_offsetpoint / abc_to_xyz      coordinates.jl:68-76
  BlockFile / parse_blockfile           coordinates.jl:83-167
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .cell import CellMatrix, prepare_periodic_distance_computations


@dataclass
class GridCoordinatesSetup:
    """coordinates.jl:15-23.  ``dims`` = number of nodes minus one along x, y, z."""
    cell: CellMatrix
    spacing: float
    dims: np.ndarray      # int32[3]
    size: np.ndarray      # float64[3]
    shift: np.ndarray     # float64[3]
    unitcell: np.ndarray  # float64[3] edge lengths
    delta: np.ndarray     # float64[3]  (Δ in the reference)

    @staticmethod
    def from_cell(cell: CellMatrix | np.ndarray, spacing: float) -> "GridCoordinatesSetup":
        """coordinates.jl:32-41 — axis-aligned bounding box of the cell."""
        if not isinstance(cell, CellMatrix):
            cell = CellMatrix.from_mat(cell)
        a, b, c = cell.mat[:, 0], cell.mat[:, 1], cell.mat[:, 2]
        size = np.abs(a) + np.abs(b) + np.abs(c)
        shift = np.minimum(a, 0.0) + np.minimum(b, 0.0) + np.minimum(c, 0.0)
        _dims = np.floor(size / spacing).astype(np.int32)
        dims = (_dims + (_dims % 2 == 0)).astype(np.int32)
        unitcell = np.array([math.sqrt(float(v @ v)) for v in (a, b, c)])
        delta = size / dims
        return GridCoordinatesSetup(cell, float(spacing), dims, size, shift, unitcell, delta)

    @property
    def nodes(self) -> tuple[int, int, int]:
        return (int(self.dims[0]) + 1, int(self.dims[1]) + 1, int(self.dims[2]) + 1)

    @property
    def npoints(self) -> int:
        nx, ny, nz = self.nodes
        return nx * ny * nz


def wrap_atom(point, cell: CellMatrix) -> np.ndarray:
    """coordinates.jl:58-61 (vectorised over leading axes)."""
    p = np.asarray(point, dtype=np.float64)
    abc = p @ cell.invmat.T
    return (abc - np.floor(abc)) @ cell.mat.T


def offsetpoint(point, csetup: GridCoordinatesSetup) -> np.ndarray:
    """coordinates.jl:63-66 — 1-based continuous grid coordinates."""
    newpoint = wrap_atom(point, csetup.cell)
    return (newpoint - csetup.shift) * csetup.dims / csetup.size + 1


def inverse_offsetpoint(ipoint, csetup: GridCoordinatesSetup) -> np.ndarray:
    """coordinates.jl:68-70."""
    return (np.asarray(ipoint, dtype=np.float64) - 1) * csetup.delta + csetup.shift


def abc_to_xyz(cset: GridCoordinatesSetup, i, j, k) -> np.ndarray:
    """coordinates.jl:72-76 — multiply, divide, add in that order."""
    return np.array([i * cset.size[0] / cset.dims[0] + cset.shift[0],
                     j * cset.size[1] / cset.dims[1] + cset.shift[1],
                     k * cset.size[2] / cset.dims[2] + cset.shift[2]])


def _round_half_even(x: np.ndarray) -> np.ndarray:
    # Julia's round(Int, x) is round-half-to-even, like numpy's rint
    return np.rint(x).astype(np.int64)


@dataclass
class BlockFile:
    """coordinates.jl:83-101.  ``block[a-1, b-1, c-1]`` (x, y, z node indices, 0-based here)."""
    csetup: GridCoordinatesSetup
    block: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 0), dtype=bool))
    empty: bool = True

    @staticmethod
    def from_mask(csetup: GridCoordinatesSetup, block: np.ndarray) -> "BlockFile":
        if block.size and block.any():
            return BlockFile(csetup, np.ascontiguousarray(block, dtype=bool), False)
        return BlockFile(csetup)

    def __getitem__(self, pos) -> bool:
        """coordinates.jl:97-101."""
        if self.empty:
            return False
        a, b, c = _round_half_even(offsetpoint(np.asarray(pos, dtype=np.float64), self.csetup))
        return bool(self.block[a - 1, b - 1, c - 1])


def read_blockfile(file, csetup: GridCoordinatesSetup):
    """Text part of parse_blockfile (coordinates.jl:113-135): sphere centres mapped through
    inverse_offsetpoint(offsetpoint(.)) and squared radii."""
    with open(file) as f:
        lines = f.read().splitlines()
    num = int(lines[0].split()[0])
    centers, radius2 = [], []
    for l in lines[1:1 + num]:
        sl = l.split()
        radius = float(sl[-1])
        _center = csetup.cell.mat @ np.array([float(x) for x in sl[:3]])
        centers.append(inverse_offsetpoint(offsetpoint(_center, csetup), csetup))
        radius2.append(radius * radius)
    return np.array(centers, dtype=np.float64).reshape(-1, 3), np.array(radius2, dtype=np.float64)


def parse_blockfile(file, csetup: GridCoordinatesSetup, device: int = 0) -> BlockFile:
    """coordinates.jl:112-167 — blocking spheres projected on the grid nodes (ceg_block_from_spheres)."""
    from . import _lib
    from .cell import flat
    centers, radius2 = read_blockfile(file, csetup)
    if len(radius2) == 0:
        return BlockFile(csetup)
    ortho, safemin = prepare_periodic_distance_computations(csetup.cell.mat)
    cell = _lib.make_cell(flat(csetup.cell.mat), flat(csetup.cell.invmat), ortho, safemin * safemin)
    mask = _lib.block_from_spheres(device, cell, _lib.make_gridspec(csetup), centers, radius2, csetup.nodes)
    return BlockFile.from_mask(csetup, mask)


def _periodic_distance2(buffer, mat, invmat, ortho, safemin2):
    """utils.jl:210-246, vectorised; returns squared distances only.  Host helper for the O(atoms^2)
    intramolecular terms of the EwaldContext constructor (ewald.jl:516-525); grid-sized work goes
    through the device."""
    f = buffer @ invmat.T
    f = f + 0.5
    f = f - np.floor(f) - 0.5
    v = f @ mat.T
    ref2 = (v * v).sum(-1)
    if ortho:
        return ref2
    out = ref2.copy()
    todo = ref2 > safemin2
    if todo.any():
        fs = f[todo]
        r2 = ref2[todo]
        res = r2.copy()
        done = np.zeros(len(r2), dtype=bool)
        for i in range(3):
            for s in (1.0, -1.0):
                g = fs.copy()
                g[:, i] += s
                w = g @ mat.T
                n2 = (w * w).sum(-1)
                hit = (~done) & (n2 < r2)
                res[hit] = n2[hit]
                done |= hit
        out[todo] = res
    return out


def blockfile_from_grid(g, device: int = 0) -> BlockFile:
    """grids.jl:188-204 — voxel-dilated mask of VdW nodes whose value exceeds 5e6 K
    (ceg_block_from_values).  ``g.grid`` has the reference layout [channel][x][y][z] (z fastest)."""
    from . import _lib
    mask = _lib.block_from_values(device, _lib.make_gridspec(g.csetup), g.grid[0], 5e6)
    return BlockFile.from_mask(g.csetup, mask)
