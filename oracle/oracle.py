"""ORACLE — test infrastructure, not product code.

ctypes front-end of oracle/libceg_oracle.so (oracle/ceg_oracle.c: the plain-C CPU restatement of
the reference algorithm) plus the few numpy pieces that are easier to restate here
(the 64x64 tricubic matrix from its definition, sparse energy_point).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
It may import the product's HOST-side parsers (ceg_b200.*, no GPU involved); the product never
imports it.

Parity status: PINNED against the reference's own known-answer energies
(test/runtests.jl:24-56,165-169,269-272) — see oracle/pin_goldens.py and
tests/test_oracle_goldens.py.  The reference itself (Julia) cannot run here, so there is no
oracle/_ref build.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import ceg_b200 as ceg  # noqa: E402  (host-side parsers and structs only)
from ceg_b200._lib import ceg_cell_t, ceg_gridspec_t, make_cell, make_gridspec  # noqa: E402

LIB_PATH = os.path.join(_HERE, "libceg_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/ceg_oracle.c (gcc -O2 -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "ceg_oracle.c")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "ceg_cuda.h")
    if (not force and os.path.isfile(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    base = [cc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c11", "-D_GNU_SOURCE", "-shared",
            "-o", LIB_PATH, src, "-lm"]
    r = subprocess.run(base[:1] + ["-fopenmp"] + base[1:], capture_output=True, text=True)
    if r.returncode != 0:
        r = subprocess.run(base, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.ceg_o_ewald_reciprocal.restype = C.c_double
        _lib.ceg_o_periodic_distance2.restype = C.c_double
        _lib.ceg_o_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def num_threads() -> int:
    return int(lib().ceg_o_num_threads())


def use_all_cores() -> int:
    """Size the OpenMP team to the cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().ceg_o_set_num_threads(C.c_int(n))
    return num_threads()


# ---- tricubic matrix from its definition (constants.jl:24-89 is its tabulated form) ---------------
def tricubic_coeff() -> np.ndarray:
    """The 64x64 Lekien–Marsden matrix: inverse of the corner-constraint matrix B, where row
    (channel c, corner p) of B evaluates the derivative d^c of x^i y^j z^k at corner p of the unit
    cube.  Channel order: f, fx, fy, fz, fxy, fxz, fyz, fxyz; corner index ix + 2 iy + 4 iz;
    monomial index i + 4 j + 16 k (grids.jl:227-258).  The matrix is integer-valued."""
    ders = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]

    def d1(power, der, x):
        if der == 0:
            return float(x ** power) if power > 0 else 1.0
        if power == 0:
            return 0.0
        return float(power * (x ** (power - 1))) if power > 1 else 1.0

    B = np.zeros((64, 64))
    for c, (dx, dy, dz) in enumerate(ders):
        for iz in (0, 1):
            for iy in (0, 1):
                for ix in (0, 1):
                    row = 8 * c + ix + 2 * iy + 4 * iz
                    for k in range(4):
                        for j in range(4):
                            for i in range(4):
                                B[row, i + 4 * j + 16 * k] = d1(i, dx, ix) * d1(j, dy, iy) * d1(k, dz, iz)
    A = np.linalg.inv(B)
    Ai = np.rint(A)
    assert np.abs(A - Ai).max() < 1e-9
    return np.ascontiguousarray(Ai)


_COEFF = None


def coeff() -> np.ndarray:
    global _COEFF
    if _COEFF is None:
        _COEFF = tricubic_coeff()
    return _COEFF


# ---- probes ---------------------------------------------------------------------------------------
def compute_derivatives(inp, positions) -> np.ndarray:
    """probes.jl:71-117 at arbitrary positions: (n_grids, n_pos, 8)."""
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    cell = inp.cell_struct()
    ng = inp.table.shape[0] + int(inp.want_coulomb)
    out = np.empty((ng, pos.shape[0], 8))
    xyz = np.ascontiguousarray(inp.positions)
    types = np.ascontiguousarray(inp.types, dtype=np.int32)
    q = np.ascontiguousarray(inp.charges) if inp.charges is not None else np.zeros(1)
    tab = np.ascontiguousarray(inp.table)
    lib().ceg_o_compute_derivatives(C.byref(cell), C.c_int64(xyz.shape[0]), _p(xyz, C.c_double), _p(types, C.c_int32),
                                    _p(q, C.c_double), C.c_int32(inp.n_types), C.c_int32(inp.table.shape[0]),
                                    C.c_void_p(tab.ctypes.data), C.c_double(inp.cutoff2), C.c_int32(int(inp.want_coulomb)),
                                    C.c_double(inp.alpha), C.c_int64(pos.shape[0]), _p(pos, C.c_double), _p(out, C.c_double))
    return out


def grid_build(inp, cset, i0=0, i1=None, istride=1, want_f32=True, want_raw=False):
    """grids.jl:137-185 for x planes i0, i0+istride, … < i1.
    Returns (list of float32 [8][nx][ny][nz] arrays (planes not computed are left 0), raw or None)."""
    nx, ny, nz = cset.nodes
    i1 = nx if i1 is None else i1
    ng = inp.table.shape[0] + int(inp.want_coulomb)
    n_planes = (i1 - i0 + istride - 1) // istride
    outs = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(ng)] if want_f32 else None
    raw = np.empty((ng, 8, n_planes, ny, nz)) if want_raw else None
    ptrs = (C.POINTER(C.c_float) * ng)()
    if outs is not None:
        for g in range(ng):
            ptrs[g] = _p(outs[g], C.c_float)
    cell = inp.cell_struct()
    gs = make_gridspec(cset)
    lam_v, thr_v = ceg.vdw_scaling()
    lam_c, thr_c = ceg.coulomb_scaling()
    xyz = np.ascontiguousarray(inp.positions)
    types = np.ascontiguousarray(inp.types, dtype=np.int32)
    q = np.ascontiguousarray(inp.charges) if inp.charges is not None else np.zeros(1)
    tab = np.ascontiguousarray(inp.table)
    lib().ceg_o_grid_build(C.byref(cell), C.c_int64(xyz.shape[0]), _p(xyz, C.c_double), _p(types, C.c_int32),
                           _p(q, C.c_double), C.c_int32(inp.n_types), C.c_int32(inp.table.shape[0]),
                           C.c_void_p(tab.ctypes.data), C.c_double(inp.cutoff2), C.c_int32(int(inp.want_coulomb)),
                           C.c_double(inp.alpha), C.byref(gs), C.c_int32(i0), C.c_int32(i1), C.c_int32(istride),
                           C.c_double(lam_v), C.c_double(thr_v), C.c_double(lam_c), C.c_double(thr_c),
                           ptrs if outs is not None else None, _p(raw, C.c_double) if raw is not None else None)
    return outs, raw


def set_gridpoint(delta, lam, thr, der8) -> np.ndarray:
    d = np.ascontiguousarray(delta, dtype=np.float64)
    der = np.ascontiguousarray(der8, dtype=np.float64)
    out = np.empty(8, dtype=np.float32)
    lib().ceg_o_set_gridpoint(_p(d, C.c_double), C.c_double(lam), C.c_double(thr), _p(der, C.c_double), _p(out, C.c_float))
    return out


def scale_to_kelvin(grid_f32: np.ndarray) -> np.ndarray:
    """parse_grid's `grid .*= GRID_TO_KELVIN` on a Float32 array (grids.jl:78)."""
    return (grid_f32.astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32)


# ---- interpolation --------------------------------------------------------------------------------
def unitcell_struct(cell):
    ortho, safemin = ceg.prepare_periodic_distance_computations(cell.mat)
    return make_cell(ceg.flat(cell.mat), ceg.flat(cell.invmat), ortho, safemin * safemin)


def interpolate(cset, grid_K: np.ndarray, is_vdw: bool, positions) -> np.ndarray:
    """grids.jl:212-258 on a dense grid [8][nx][ny][nz] (already in K)."""
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    g = np.ascontiguousarray(grid_K, dtype=np.float32)
    out = np.empty(pos.shape[0])
    uc = unitcell_struct(cset.cell)
    gs = make_gridspec(cset)
    co = coeff()
    lib().ceg_o_interpolate(C.byref(uc), C.byref(gs), _p(g, C.c_float), C.c_int32(int(is_vdw)), _p(co, C.c_double),
                            C.c_int64(pos.shape[0]), _p(pos, C.c_double), _p(out, C.c_double))
    return out


def block_lookup(cset, mask, positions) -> np.ndarray:
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty(pos.shape[0], dtype=np.uint8)
    uc = unitcell_struct(cset.cell)
    gs = make_gridspec(cset)
    lib().ceg_o_block_lookup(C.byref(uc), C.byref(gs), _p(m, C.c_uint8), C.c_int64(pos.shape[0]), _p(pos, C.c_double),
                             _p(out, C.c_uint8))
    return out.astype(bool)


def interpolate_sparse(inp, g_index: int, cset, is_vdw: bool, point) -> float:
    """interpolate_grid at ONE point without building the whole grid: the 8 corner nodes are
    computed on demand (compute_derivatives -> _set_gridpoint! -> Float32 -> x GRID_TO_KELVIN), then
    grids.jl:227-258 is applied."""
    s = ceg.offsetpoint(np.asarray(point, dtype=np.float64), cset)
    nn = np.array(cset.nodes)
    p0 = np.floor(s).astype(np.int64)
    p1 = p0 + (p0 != nn)
    r = s - p0
    lam, thr = ceg.vdw_scaling() if is_vdw else ceg.coulomb_scaling()
    X = np.zeros(64)
    for iz in (0, 1):
        for iy in (0, 1):
            for ix in (0, 1):
                node = (int(p1[0] if ix else p0[0]) - 1, int(p1[1] if iy else p0[1]) - 1, int(p1[2] if iz else p0[2]) - 1)
                pos = ceg.abc_to_xyz(cset, *node)
                der = compute_derivatives(inp, pos)[g_index, 0]
                q8 = scale_to_kelvin(set_gridpoint(cset.delta, lam, thr, der))
                for c in range(8):
                    X[8 * c + ix + 2 * iy + 4 * iz] = float(q8[c])
    if is_vdw and (X[:8] > 5e6).any():
        return 1e100
    a = coeff() @ X
    rxs = [1.0, r[0], r[0] * r[0], r[0] * r[0] * r[0]]
    rys = [1.0, r[1], r[1] * r[1], r[1] * r[1] * r[1]]
    rzs = [1.0, r[2], r[2] * r[2], r[2] * r[2] * r[2]]
    ret = 0.0
    for k in range(4):
        for j in range(4):
            for i in range(4):
                ret += a[i + 4 * j + 16 * k] * rxs[i] * rys[j] * rzs[k]
    return ret


# ---- Ewald ------------------------------------------------------------------------------------------
class OracleEwald:
    """initialize_ewald (ewald.jl:195-281) with S_f from the C restatement."""

    def __init__(self, syst, supercell=None, precision=1e-6):
        if supercell is None:
            supercell = ceg.find_supercell(syst.mat, 12.0)
        self.supercell = tuple(int(x) for x in supercell)
        self.alpha, tol1 = ceg.ewald_parameters(precision)
        self.mat = np.asarray(syst.mat, dtype=np.float64) * np.array(self.supercell, dtype=np.float64)[None, :]
        self.invmat = ceg.inv3(self.mat)
        self.kspace = ceg.build_kspace(self.mat, self.alpha, tol1)
        self.kfactors = ceg.build_kfactors(self.kspace, self.mat, self.invmat, self.alpha)
        self.UIon = ceg.COULOMBIC_CONVERSION_FACTOR * self.alpha / math.sqrt(math.pi) - float(np.sum(self.kfactors))
        frac, charges = ceg.framework_sites(syst, self.supercell)
        self.net_charges_framework = float(np.sum(charges)) if len(charges) else 0.0
        self.precision = precision
        self.rows = self.kspace.rows_array()
        self.kmax = np.array(self.kspace.ks, dtype=np.int32)
        self.S = structure_factor(self.kmax, self.rows, self.kspace.num_kvecs, frac, charges)

    def as_framework(self):
        """An EwaldFramework (host fields only, no device handle) for ewald_context()."""
        return ceg.EwaldFramework(self.kspace, self.alpha, self.mat, self.invmat, self.kfactors, self.UIon, self.S,
                                  self.net_charges_framework, self.precision, None)

    def reciprocal(self, positions, q, energy_net_charges, static) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float64)
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, q.shape[0], 3)
        out = np.empty(pos.shape[0])
        inv9 = ceg.flat(self.invmat)
        Sri = np.ascontiguousarray(np.stack([self.S.real, self.S.imag], axis=1))
        lib().ceg_o_ewald_reciprocal_batch(_p(inv9, C.c_double), _p(self.kmax, C.c_int32), C.c_int64(self.rows.shape[0]),
                                           _p(self.rows, C.c_int32), C.c_int64(self.kspace.num_kvecs),
                                           _p(self.kfactors, C.c_double), _p(Sri, C.c_double), C.c_int64(pos.shape[0]),
                                           C.c_int32(q.shape[0]), _p(pos, C.c_double), _p(q, C.c_double),
                                           C.c_double(energy_net_charges), C.c_double(static), _p(out, C.c_double))
        return out


def structure_factor(kmax, rows, n_k, frac, charges) -> np.ndarray:
    frac = np.ascontiguousarray(frac, dtype=np.float64).reshape(-1, 3)
    charges = np.ascontiguousarray(charges, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    kmax = np.ascontiguousarray(kmax, dtype=np.int32)
    out = np.zeros((n_k, 2))
    lib().ceg_o_structure_factor(_p(kmax, C.c_int32), C.c_int64(rows.shape[0]), _p(rows, C.c_int32), C.c_int64(n_k),
                                 C.c_int64(frac.shape[0]), _p(frac, C.c_double), _p(charges, C.c_double),
                                 _p(out, C.c_double))
    return out[:, 0] + 1j * out[:, 1]


def compute_ewald(oe: OracleEwald, systems) -> float:
    """compute_ewald(eframework, systems) (ewald.jl:579-581)."""
    ctx = ceg.ewald_context(oe.as_framework(), systems)
    return float(oe.reciprocal(ctx.positions[None], ctx.charges, ctx.energy_net_charges, ctx.static_contribution)[0])


# ---- blocking masks ---------------------------------------------------------------------------------------
def parse_blockfile(file, cset) -> np.ndarray:
    """coordinates.jl:112-167 restated with numpy: boolean mask [nx][ny][nz]."""
    from ceg_b200.coordinates import _periodic_distance2, inverse_offsetpoint, read_blockfile
    centers, radius2 = read_blockfile(file, cset)
    a, b, c = cset.nodes
    mask = np.zeros((a, b, c), dtype=bool)
    if len(radius2) == 0:
        return mask
    ortho, safemin = ceg.prepare_periodic_distance_computations(cset.cell.mat)
    ii, jj, kk = np.meshgrid(np.arange(1, a + 1), np.arange(1, b + 1), np.arange(1, c + 1), indexing="ij")
    nodes = inverse_offsetpoint(np.stack([ii, jj, kk], axis=-1), cset)
    for ctr, r2 in zip(centers, radius2):
        mask |= _periodic_distance2(ctr - nodes, cset.cell.mat, cset.cell.invmat, ortho, safemin * safemin) < r2
    return mask


def blockfile_from_values(values: np.ndarray, threshold: float = 5e6) -> np.ndarray:
    """grids.jl:188-204 restated with numpy."""
    a, b, c = values.shape
    hot = np.zeros((a, b, c), dtype=bool)
    hot[:a - 1, :b - 1, :c - 1] = values[:a - 1, :b - 1, :c - 1] > threshold
    block = np.zeros((a, b, c), dtype=bool)
    for di in (0, 1):
        for dj in (0, 1):
            for dk in (0, 1):
                block[di:, dj:, dk:] |= hot[:a - di, :b - dj, :c - dk]
    return block


# ---- guest-guest pair energies ------------------------------------------------------------------------------
def guest_energy(env, mol_species, trial_positions, skip_molecule=-1) -> np.ndarray:
    """energy.jl:407-427 batched over trial placements (env: ceg_b200.GuestEnvironment)."""
    ms = np.ascontiguousarray(mol_species, dtype=np.int32)
    tp = np.ascontiguousarray(trial_positions, dtype=np.float64).reshape(-1, ms.shape[0], 3)
    out = np.empty(tp.shape[0])
    cell = env.cell_struct()
    tab = np.ascontiguousarray(env.table)
    exyz = np.ascontiguousarray(env.positions, dtype=np.float64)
    esp = np.ascontiguousarray(env.species, dtype=np.int32)
    emol = np.ascontiguousarray(env.molecule, dtype=np.int32)
    lib().ceg_o_guest_energy(C.byref(cell), C.c_double(float(env.cutoff) ** 2), C.c_int32(tab.shape[0]),
                             C.c_void_p(tab.ctypes.data), C.c_int64(exyz.shape[0]), _p(exyz, C.c_double), _p(esp, C.c_int32),
                             _p(emol, C.c_int32), C.c_int64(tp.shape[0]), C.c_int32(ms.shape[0]), _p(ms, C.c_int32),
                             C.c_int32(skip_molecule), _p(tp, C.c_double), _p(out, C.c_double))
    return out
