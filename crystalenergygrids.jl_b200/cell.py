"""Unit-cell geometry helpers (host side, FP64, reference operation order).

Mirrors the hot-path parts of /root/reference/src/utils.jl:
  perpendicular_lengths / find_supercell            utils.jl:10-67
  cell_parameters / mat_from_parameters             utils.jl:128-144
  prepare_periodic_distance_computations            utils.jl:146-155
  CellMatrix                                        utils.jl:282-292

Convention: ``mat`` is a 3x3 float64 array whose COLUMNS are the cell vectors a, b, c
(same as the reference's SMatrix).  ``flat(mat)`` gives the 9 doubles in column-major order,
which is the memory layout the C ABI (include/ceg_cuda.h) and Julia both use.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


def flat(mat) -> np.ndarray:
    """9 doubles, column-major (a_x a_y a_z b_x ... c_z)."""
    return np.ascontiguousarray(np.asarray(mat, dtype=np.float64).T).reshape(9).copy()


def inv3(mat: np.ndarray) -> np.ndarray:
    """Closed-form 3x3 inverse (cross products / determinant), the form StaticArrays uses
    for `inv(::SMatrix{3,3})` (probes.jl:28, utils.jl:286)."""
    m = np.asarray(mat, dtype=np.float64)
    x0, x1, x2 = m[:, 0].copy(), m[:, 1].copy(), m[:, 2].copy()

    def cross(u, v):
        return np.array([u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]])

    y0 = cross(x1, x2)
    d = x0[0] * y0[0] + x0[1] * y0[1] + x0[2] * y0[2]
    x0 = x0 / d
    y0 = y0 / d
    y1 = cross(x2, x0)
    y2 = cross(x0, x1)
    return np.array([y0, y1, y2], dtype=np.float64)  # rows are y0,y1,y2


def mat_from_parameters(lengths, angles_deg) -> np.ndarray:
    """utils.jl:140-144 (a along x, b in the xy plane) — also the Chemfiles CIF convention."""
    a, b, c = lengths
    al, be, ga = (math.radians(x) for x in angles_deg)
    cosa, cosb = math.cos(al), math.cos(be)
    sing, cosg = math.sin(ga), math.cos(ga)
    # exact right angles (cosd(90) == 0 in Julia)
    if angles_deg[0] == 90:
        cosa = 0.0
    if angles_deg[1] == 90:
        cosb = 0.0
    if angles_deg[2] == 90:
        cosg, sing = 0.0, 1.0
    omega = math.sqrt(1 - cosa * cosa - cosb * cosb - cosg * cosg + 2 * cosa * cosb * cosg)
    return np.array([[a, b * cosg, c * cosb],
                     [0.0, b * sing, c * (cosa - cosb * cosg) / sing],
                     [0.0, 0.0, c * omega / sing]], dtype=np.float64)


def cell_parameters(mat):
    """utils.jl:128-138."""
    m = np.asarray(mat, dtype=np.float64)
    _a, _b, _c = m[:, 0], m[:, 1], m[:, 2]
    a = math.sqrt(float(_a @ _a))
    b = math.sqrt(float(_b @ _b))
    c = math.sqrt(float(_c @ _c))

    def acosd(x):
        return math.degrees(math.acos(max(-1.0, min(1.0, x))))

    return (a, b, c), (acosd(float(_b @ _c) / (b * c)), acosd(float(_c @ _a) / (c * a)), acosd(float(_a @ _b) / (a * b)))


def perpendicular_lengths(mat) -> tuple[float, float, float]:
    """utils.jl:10-29."""
    m = np.asarray(mat, dtype=np.float64)
    a, b, c = m[:, 0], m[:, 1], m[:, 2]
    axb = np.cross(a, b)
    bxc = np.cross(b, c)
    cxa = np.cross(c, a)
    volume = abs(float(a @ bxc))
    return (volume / float(np.linalg.norm(bxc)), volume / float(np.linalg.norm(cxa)), volume / float(np.linalg.norm(axb)))


def find_supercell(mat, cutoff: float) -> tuple[int, int, int]:
    """utils.jl:43-67 — smallest replication with perpendicular widths >= 2*cutoff."""
    cx, cy, cz = perpendicular_lengths(mat)
    return (math.ceil(2 * cutoff / cx), math.ceil(2 * cutoff / cy), math.ceil(2 * cutoff / cz))


def prepare_periodic_distance_computations(mat) -> tuple[bool, float]:
    """utils.jl:146-155 — returns (ortho, safemin).

    ortho: every cell angle, rounded to Float16, is within 2% of 90 degrees.
    safemin: half of min(V/(|b||c|), V/(|a||c|), V/(|a||b|))  (signed volume, as written).
    """
    m = np.asarray(mat, dtype=np.float64)
    (a, b, c), angles = cell_parameters(m)

    def approx90(x):
        x16 = np.float16(x)
        if not np.isfinite(x16):
            return False
        diff = abs(float(np.float16(x16 - np.float16(90))))
        return float(x16) == 90.0 or diff <= 0.02 * max(abs(float(x16)), 90.0)

    ortho = all(approx90(x) for x in angles)
    _a, _b, _c = m[:, 0], m[:, 1], m[:, 2]
    safemin = min(float(np.cross(_b, _c) @ _a) / (b * c),
                  float(np.cross(_c, _a) @ _b) / (a * c),
                  float(np.cross(_a, _b) @ _c) / (a * b)) / 2
    return ortho, safemin


@dataclass(frozen=True)
class CellMatrix:
    """utils.jl:282-292."""
    mat: np.ndarray
    invmat: np.ndarray

    @staticmethod
    def from_mat(mat) -> "CellMatrix":
        m = np.array(mat, dtype=np.float64).reshape(3, 3)
        return CellMatrix(m, inv3(m))
