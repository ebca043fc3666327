"""Multi-GPU partitioning of the hot path: no data-path collective.

Grid construction shards the x index range 0:dims[1] (the reference's own @threads axis,
grids.jl:144,171) into contiguous slabs, one per rank/GPU; x is the slowest spatial index of the
.grid layout, so a slab is 8 contiguous runs (one per channel) of the host/file array.
Interpolation shards queries by index; every rank holds a replica of the grids it needs.
"""
from __future__ import annotations

import numpy as np


def split_range(n: int, parts: int) -> list[tuple[int, int]]:
    """`parts` contiguous [start, stop) ranges covering 0..n, sizes differing by at most one."""
    base, rem = divmod(int(n), int(parts))
    out, s = [], 0
    for p in range(parts):
        e = s + base + (1 if p < rem else 0)
        out.append((s, e))
        s = e
    return out


def slab_of_rank(n_planes: int, world: int, rank: int) -> tuple[int, int]:
    return split_range(n_planes, world)[rank]


def slab_file_runs(nodes: tuple[int, int, int], i0: int, i1: int, header_bytes: int = 128):
    """Byte ranges of x-slab [i0,i1) inside a .grid file (grids.jl:61-94 layout): one run per channel."""
    nx, ny, nz = nodes
    plane = ny * nz * 4
    return [(header_bytes + c * nx * plane + i0 * plane, (i1 - i0) * plane) for c in range(8)]


def place_slab(full: np.ndarray, slab: np.ndarray, i0: int) -> None:
    """Copy a [8][i1-i0][ny][nz] slab into the full [8][nx][ny][nz] array."""
    full[:, i0:i0 + slab.shape[1]] = slab
