"""ctypes binding of libceg_cuda.so (include/ceg_cuda.h) — the only way the host layer computes.

There is deliberately no fallback: if the shared library is missing or no CUDA device is usable,
every compute entry point raises.  (The CPU oracle under oracle/ is test infrastructure and is
never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CEG_LIB_PATH: a differently-compiled libceg_cuda.so for kernel experiments (tools/); same ABI, still CUDA only
LIB_PATH = os.environ.get("CEG_LIB_PATH") or os.path.join(_HERE, "libceg_cuda.so")

c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_uint8_p = C.POINTER(C.c_uint8)


class CegError(RuntimeError):
    pass


class ceg_cell_t(C.Structure):
    _fields_ = [("mat", C.c_double * 9), ("invmat", C.c_double * 9), ("ortho", C.c_int32), ("_pad", C.c_int32),
                ("safemin2", C.c_double)]


class ceg_gridspec_t(C.Structure):
    _fields_ = [("dims", C.c_int32 * 3), ("_pad", C.c_int32), ("size", C.c_double * 3), ("shift", C.c_double * 3),
                ("delta", C.c_double * 3)]


EXPORTS = [
    "ceg_last_error", "ceg_abi_version", "ceg_device_count", "ceg_device_info", "ceg_trim_device_cache", "ceg_kernel_launch_count",
    "ceg_measure_fp64_peak", "ceg_measure_fp64_peak_regops", "ceg_selftest_math", "ceg_host_register",
    "ceg_host_unregister", "ceg_probe_create", "ceg_probe_free", "ceg_probe_num_grids", "ceg_compute_derivatives",
    "ceg_grid_build", "ceg_grid_build_dev", "ceg_probe_last_stats", "ceg_grid_upload", "ceg_grid_upload_dev", "ceg_grid_free",
    "ceg_interpolate", "ceg_interpolate_dev", "ceg_interpolate_multi", "ceg_interpolate_sweep", "ceg_energy_lattice", "ceg_set_interp_layout", "ceg_grid_layout", "ceg_ewald_create", "ceg_ewald_free", "ceg_ewald_reciprocal",
    "ceg_ewald_single_contribution", "ceg_ewald_reciprocal_dev", "ceg_guest_energy", "ceg_guest_env_create", "ceg_guest_env_free", "ceg_guest_env_update", "ceg_guest_env_energy", "ceg_guest_env_last_kernel_ms", "ceg_block_upload", "ceg_block_free", "ceg_block_from_spheres", "ceg_block_from_values", "ceg_energy_points",
    "ceg_probe_last_timing", "ceg_grid_gather", "ceg_kernel_registers", "ceg_device_malloc", "ceg_device_free",
    "ceg_device_download", "ceg_device_synchronize", "ceg_grid_pair_create", "ceg_grid_pair_free", "ceg_interpolate_pair",
    "ceg_interpolate_pair_dev", "ceg_ewald_create_from_sf", "ceg_grid_upload_parts_dev",
    "ceg_measure_launch_roundtrip", "ceg_grid_create_empty", "ceg_grid_records", "ceg_grid_build_records", "ceg_ipc_export", "ceg_ipc_open", "ceg_ipc_close",
]

_lib = None
_lock = threading.Lock()


def load():
    """Load libceg_cuda.so (once).  Raises CegError with build instructions if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise CegError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.ceg_last_error.restype = C.c_char_p
        for name in EXPORTS:
            if not hasattr(lib, name):
                raise CegError(f"libceg_cuda.so does not export {name}")
            if name != "ceg_last_error":
                getattr(lib, name).restype = C.c_int
        _lib = lib
        return lib


def _check(rc: int):
    if rc != 0:
        msg = load().ceg_last_error()
        raise CegError(msg.decode("utf-8", "replace") if msg else f"libceg_cuda error {rc}")


def _p(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _f64(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    return out if shape is None else out.reshape(shape)


def make_cell(mat9, invmat9, ortho: bool, safemin2: float) -> ceg_cell_t:
    c = ceg_cell_t()
    c.mat[:] = [float(x) for x in mat9]
    c.invmat[:] = [float(x) for x in invmat9]
    c.ortho = int(bool(ortho))
    c.safemin2 = float(safemin2)
    return c


def make_gridspec(csetup) -> ceg_gridspec_t:
    g = ceg_gridspec_t()
    g.dims[:] = [int(x) for x in csetup.dims]
    g.size[:] = [float(x) for x in csetup.size]
    g.shift[:] = [float(x) for x in csetup.shift]
    g.delta[:] = [float(x) for x in csetup.delta]
    return g


# ---------------------------------------------------------------------------------------------
def device_count() -> int:
    n = C.c_int32(0)
    _check(load().ceg_device_count(C.byref(n)))
    return n.value


def device_info(device: int = 0) -> dict:
    name = C.create_string_buffer(256)
    sm, mem, clk = C.c_int32(0), C.c_int64(0), C.c_int32(0)
    _check(load().ceg_device_info(C.c_int32(device), name, C.byref(sm), C.byref(mem), C.byref(clk)))
    return {"name": name.value.decode(), "sm_count": sm.value, "mem_bytes": mem.value, "clock_khz": clk.value}


def trim_device_cache():
    _check(load().ceg_trim_device_cache())


def kernel_launch_count() -> int:
    n = C.c_int64(0)
    _check(load().ceg_kernel_launch_count(C.byref(n)))
    return n.value


def measure_fp64_peak(device: int = 0) -> float:
    t = C.c_double(0.0)
    _check(load().ceg_measure_fp64_peak(C.c_int32(device), C.byref(t)))
    return t.value


def measure_fp64_peak_regops(device: int = 0) -> float:
    t = C.c_double(0.0)
    _check(load().ceg_measure_fp64_peak_regops(C.c_int32(device), C.byref(t)))
    return t.value


def selftest_math(x, device: int = 0) -> np.ndarray:
    x = _f64(x, (-1,))
    out = np.empty((x.shape[0], 6), dtype=np.float64)
    _check(load().ceg_selftest_math(C.c_int32(device), C.c_int64(x.shape[0]), _p(x, C.c_double), _p(out, C.c_double)))
    return out


def host_register(arr: np.ndarray):
    _check(load().ceg_host_register(C.c_void_p(arr.ctypes.data), C.c_int64(arr.nbytes)))


def host_unregister(arr: np.ndarray):
    _check(load().ceg_host_unregister(C.c_void_p(arr.ctypes.data)))


# ---- probes / grid construction --------------------------------------------------------------
def probe_create(device: int, cell: ceg_cell_t, xyz, types, charges, n_types: int, table: np.ndarray, cutoff2: float,
                 want_coulomb: bool, alpha: float):
    xyz = _f64(xyz, (-1, 3))
    types = np.ascontiguousarray(types, dtype=np.int32)
    n = xyz.shape[0]
    assert types.shape == (n,)
    q = None
    if charges is not None:
        q = _f64(charges, (n,))
    table = np.ascontiguousarray(table)
    assert table.dtype.itemsize == 72
    n_probes = table.shape[0] if table.size else 0
    h = C.c_void_p()
    _check(load().ceg_probe_create(C.c_int32(device), C.byref(cell), C.c_int64(n), _p(xyz, C.c_double),
                                   _p(types, C.c_int32), _p(q, C.c_double) if q is not None else None,
                                   C.c_int32(n_types), C.c_int32(n_probes), C.c_void_p(table.ctypes.data if table.size else 0),
                                   C.c_double(cutoff2), C.c_int32(int(want_coulomb)), C.c_double(alpha), C.byref(h)))
    return h


def probe_free(h):
    _check(load().ceg_probe_free(h))


def probe_num_grids(h) -> int:
    n = C.c_int32(0)
    _check(load().ceg_probe_num_grids(h, C.byref(n)))
    return n.value


def compute_derivatives(h, xyz) -> np.ndarray:
    xyz = _f64(xyz, (-1, 3))
    ng = probe_num_grids(h)
    out = np.empty((ng, xyz.shape[0], 8), dtype=np.float64)
    _check(load().ceg_compute_derivatives(h, C.c_int64(xyz.shape[0]), _p(xyz, C.c_double), _p(out, C.c_double)))
    return out


def grid_build(h, gs: ceg_gridspec_t, i0: int, i1: int, lam_vdw, thr_vdw, lam_coul, thr_coul, outs, raw=None):
    """outs: list of n_grids float32 arrays [8][nx][ny][nz] (or None)."""
    ng = probe_num_grids(h)
    assert len(outs) == ng
    ptrs = (c_float_p * ng)()
    for g, o in enumerate(outs):
        if o is not None:
            assert o.dtype == np.float32 and o.flags.c_contiguous
            ptrs[g] = _p(o, C.c_float)
    rawp = None
    if raw is not None:
        assert raw.dtype == np.float64 and raw.flags.c_contiguous
        rawp = _p(raw, C.c_double)
    _check(load().ceg_grid_build(h, C.byref(gs), C.c_int32(i0), C.c_int32(i1), C.c_double(lam_vdw), C.c_double(thr_vdw),
                                 C.c_double(lam_coul), C.c_double(thr_coul), ptrs, rawp))


def grid_build_dev(h, gs: ceg_gridspec_t, i0: int, i1: int, lam_vdw, thr_vdw, lam_coul, thr_coul, dev_ptrs, stream=0):
    ng = probe_num_grids(h)
    assert len(dev_ptrs) == ng
    ptrs = (C.c_void_p * ng)(*[C.c_void_p(int(p)) for p in dev_ptrs])
    _check(load().ceg_grid_build_dev(h, C.byref(gs), C.c_int32(i0), C.c_int32(i1), C.c_double(lam_vdw),
                                     C.c_double(thr_vdw), C.c_double(lam_coul), C.c_double(thr_coul), ptrs,
                                     C.c_void_p(int(stream))))


def probe_last_stats(h) -> dict:
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    _check(load().ceg_probe_last_stats(h, C.byref(a), C.byref(b), C.byref(c)))
    return {"n_points": a.value, "pairs_tested": b.value, "pairs_in_cutoff": c.value}


def probe_last_timing(h) -> dict:
    """Phases of the last ceg_grid_build on this handle (ms)."""
    t = (C.c_double * 4)()
    _check(load().ceg_probe_last_timing(h, t))
    return {"image_list_host_ms": t[0], "kernel_span_ms": t[1], "copy_tail_ms": t[2], "call_ms": t[3]}


def kernel_registers(which: int, device: int = 0) -> int:
    r = C.c_int32(0)
    _check(load().ceg_kernel_registers(C.c_int32(device), C.c_int32(which), C.byref(r)))
    return r.value


# ---- plain device buffers (chaining the _dev entry points from the host layer) ---------------------------
def device_malloc(device: int, nbytes: int) -> int:
    p = C.c_void_p()
    _check(load().ceg_device_malloc(C.c_int32(device), C.c_int64(int(nbytes)), C.byref(p)))
    return int(p.value)


def device_free(ptr: int):
    _check(load().ceg_device_free(C.c_void_p(int(ptr))))


def device_download(device: int, ptr: int, out: np.ndarray):
    assert out.flags.c_contiguous
    _check(load().ceg_device_download(C.c_int32(device), C.c_void_p(int(ptr)), C.c_int64(out.nbytes), C.c_void_p(out.ctypes.data)))


def device_synchronize(device: int):
    _check(load().ceg_device_synchronize(C.c_int32(device)))


def grid_gather(devices, d_slabs, i0s, i1s, gs: ceg_gridspec_t, dst_device: int, d_full: int):
    """x-slabs [8][i1-i0][ny][nz] on several devices of this process -> full [8][nx][ny][nz] grid on dst_device."""
    n = len(devices)
    dv = np.ascontiguousarray(devices, dtype=np.int32)
    a = np.ascontiguousarray(i0s, dtype=np.int32)
    b = np.ascontiguousarray(i1s, dtype=np.int32)
    ptrs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in d_slabs])
    _check(load().ceg_grid_gather(C.c_int32(n), _p(dv, C.c_int32), ptrs, _p(a, C.c_int32), _p(b, C.c_int32), C.byref(gs),
                                  C.c_int32(dst_device), C.c_void_p(int(d_full))))


# ---- interpolation -----------------------------------------------------------------------------
def grid_upload(device: int, cell: ceg_cell_t, gs: ceg_gridspec_t, grid: np.ndarray, is_vdw: bool):
    assert grid.dtype == np.float32 and grid.flags.c_contiguous
    h = C.c_void_p()
    _check(load().ceg_grid_upload(C.c_int32(device), C.byref(cell), C.byref(gs), _p(grid, C.c_float),
                                  C.c_int32(int(is_vdw)), C.byref(h)))
    return h


def grid_upload_dev(device: int, cell: ceg_cell_t, gs: ceg_gridspec_t, d_grid: int, scale: float, is_vdw: bool):
    """Device-resident [8][nx][ny][nz] f32 grid -> grid handle; values become Float32(value * scale)."""
    h = C.c_void_p()
    _check(load().ceg_grid_upload_dev(C.c_int32(device), C.byref(cell), C.byref(gs), C.c_void_p(int(d_grid)),
                                      C.c_double(scale), C.c_int32(int(is_vdw)), C.byref(h)))
    return h


def grid_upload_parts_dev(device: int, cell: ceg_cell_t, gs: ceg_gridspec_t, d_parts, i0s, i1s, strides, scale: float,
                          is_vdw: bool):
    """x-slabs on this device (after an all-gather) -> grid handle, without assembling the full grid."""
    n = len(d_parts)
    ptrs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in d_parts])
    a = np.ascontiguousarray(i0s, dtype=np.int32)
    b = np.ascontiguousarray(i1s, dtype=np.int32)
    st = np.ascontiguousarray(strides, dtype=np.int64) if strides is not None else None
    h = C.c_void_p()
    _check(load().ceg_grid_upload_parts_dev(C.c_int32(device), C.byref(cell), C.byref(gs), C.c_int32(n), ptrs, _p(a, C.c_int32),
                                            _p(b, C.c_int32), _p(st, C.c_int64) if st is not None else None, C.c_double(scale),
                                            C.c_int32(int(is_vdw)), C.byref(h)))
    return h


def measure_launch_roundtrip(device: int = 0, reps: int = 300) -> float:
    """Median µs of one empty launch through the latency path (launch -> mapped completion flag seen by the CPU)."""
    t = C.c_double(0.0)
    _check(load().ceg_measure_launch_roundtrip(C.c_int32(device), C.c_int32(reps), C.byref(t)))
    return t.value


def grid_create_empty(device: int, cell: ceg_cell_t, gs: ceg_gridspec_t, is_vdw: bool):
    """Grid handle whose records a build will write (ceg_grid_build_records)."""
    h = C.c_void_p()
    _check(load().ceg_grid_create_empty(C.c_int32(device), C.byref(cell), C.byref(gs), C.c_int32(int(is_vdw)), C.byref(h)))
    return h


def grid_records(h):
    """(device pointer, bytes, layout) of a grid's record storage."""
    p, n, lay = C.c_void_p(), C.c_int64(0), C.c_int32(0)
    _check(load().ceg_grid_records(h, C.byref(p), C.byref(n), C.byref(lay)))
    return int(p.value), n.value, lay.value


def grid_build_records(h, gs: ceg_gridspec_t, i0: int, i1: int, lam_v: float, thr_v: float, lam_c: float, thr_c: float,
                       dest_records, layout: int, scales, d_out=None, stream: int = 0):
    """Build planes [i0, i1) straight into interpolation records: dest_records[t][g] = record pointer of grid g at
    destination t (local handle, peer device or IPC mapping); d_out: optional `.grid`-order device buffers."""
    n_dest = len(dest_records)
    flat = [int(p) for dest in dest_records for p in dest]
    ptrs = (C.c_void_p * len(flat))(*[C.c_void_p(p) for p in flat])
    sc = np.ascontiguousarray(scales, dtype=np.float64)
    outs = (C.c_void_p * len(d_out))(*[C.c_void_p(int(p)) for p in d_out]) if d_out is not None else None
    _check(load().ceg_grid_build_records(h, C.byref(gs), C.c_int32(i0), C.c_int32(i1), C.c_double(lam_v), C.c_double(thr_v),
                                         C.c_double(lam_c), C.c_double(thr_c), C.c_int32(n_dest), ptrs, C.c_int32(layout),
                                         _p(sc, C.c_double), outs, C.c_void_p(stream)))


def ipc_export(device: int, d_ptr: int):
    """(64-byte handle, offset) of a device buffer for another process (CUDA IPC)."""
    hd = (C.c_uint8 * 64)()
    off = C.c_int64(0)
    _check(load().ceg_ipc_export(C.c_int32(device), C.c_void_p(int(d_ptr)), hd, C.byref(off)))
    return bytes(hd), off.value


def ipc_open(device: int, handle: bytes, offset: int) -> int:
    hd = (C.c_uint8 * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    _check(load().ceg_ipc_open(C.c_int32(device), hd, C.c_int64(offset), C.byref(p)))
    return int(p.value)


def ipc_close(device: int, d_ptr: int):
    _check(load().ceg_ipc_close(C.c_int32(device), C.c_void_p(int(d_ptr))))


def set_interp_layout(layout: int):
    """-1 auto, 0 node records, 1 quads, 2 cells (include/ceg_cuda.h)."""
    _check(load().ceg_set_interp_layout(C.c_int32(layout)))


def grid_layout(h) -> int:
    n = C.c_int32(0)
    _check(load().ceg_grid_layout(h, C.byref(n)))
    return n.value


def grid_free(h):
    _check(load().ceg_grid_free(h))


def interpolate(h, xyz) -> np.ndarray:
    xyz = _f64(xyz, (-1, 3))
    out = np.empty(xyz.shape[0], dtype=np.float64)
    _check(load().ceg_interpolate(h, C.c_int64(xyz.shape[0]), _p(xyz, C.c_double), _p(out, C.c_double)))
    return out


def interpolate_multi(handles, xyz, outs=None) -> np.ndarray:
    """(n_grids, n) energies of several grids at the same positions (uploaded once)."""
    xyz = _f64(xyz, (-1, 3))
    G, n = len(handles), xyz.shape[0]
    if outs is None:
        outs = np.empty((G, n), dtype=np.float64)
    assert outs.shape == (G, n) and outs.dtype == np.float64 and outs.flags.c_contiguous
    hs = (C.c_void_p * G)(*handles)
    ptrs = (c_double_p * G)(*[_p(outs[g], C.c_double) for g in range(G)])
    _check(load().ceg_interpolate_multi(C.c_int32(G), hs, C.c_int64(n), _p(xyz, C.c_double), ptrs))
    return outs


def interpolate_sweep(h, first: int, n: int, seed: int, mat, n_keep: int = 0):
    """-> dict(sum, min, n_blocked, keep[n_keep], kernel_ms): energies at n device-generated positions."""
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.float64).T).ravel()     # column-major
    s, mn, nb, ms = C.c_double(), C.c_double(), C.c_int64(), C.c_float()
    n_keep = int(min(n_keep, n))
    keep = np.empty(n_keep, dtype=np.float64)
    _check(load().ceg_interpolate_sweep(h, C.c_int64(first), C.c_int64(n), C.c_uint64(seed), _p(m, C.c_double),
                                        C.byref(s), C.byref(mn), C.byref(nb), C.c_int64(n_keep),
                                        _p(keep, C.c_double) if n_keep else None, C.byref(ms)))
    return {"sum": s.value, "min": mn.value, "n_blocked": nb.value, "keep": keep, "kernel_ms": ms.value}


def grid_pair_create(vdw_handle, coulomb_handle):
    h = C.c_void_p()
    _check(load().ceg_grid_pair_create(vdw_handle, coulomb_handle, C.byref(h)))
    return h


def grid_pair_free(h):
    _check(load().ceg_grid_pair_free(h))


def interpolate_pair(h, xyz, outs=None) -> np.ndarray:
    """(2, n): VdW and Coulomb grid values at the same positions through the fused pair records."""
    xyz = _f64(xyz, (-1, 3))
    n = xyz.shape[0]
    if outs is None:
        outs = np.empty((2, n), dtype=np.float64)
    assert outs.shape == (2, n) and outs.dtype == np.float64 and outs.flags.c_contiguous
    _check(load().ceg_interpolate_pair(h, C.c_int64(n), _p(xyz, C.c_double), _p(outs[0], C.c_double), _p(outs[1], C.c_double)))
    return outs


def interpolate_pair_dev(h, n: int, d_xyz: int, d_out_vdw: int, d_out_coul: int, stream: int = 0):
    _check(load().ceg_interpolate_pair_dev(h, C.c_int64(n), C.c_void_p(int(d_xyz)), C.c_void_p(int(d_out_vdw)),
                                           C.c_void_p(int(d_out_coul)), C.c_void_p(int(stream))))


def interpolate_dev(h, n: int, d_xyz: int, d_out: int, stream: int = 0):
    _check(load().ceg_interpolate_dev(h, C.c_int64(n), C.c_void_p(int(d_xyz)), C.c_void_p(int(d_out)),
                                      C.c_void_p(int(stream))))


# ---- Ewald -------------------------------------------------------------------------------------
def ewald_create(device: int, invmat9, kmax, kvec, kfactor, frac, charge):
    invmat9 = _f64(invmat9, (9,))
    kmax = np.ascontiguousarray(kmax, dtype=np.int32)
    kvec = np.ascontiguousarray(kvec, dtype=np.int32).reshape(-1, 3)
    kfactor = _f64(kfactor, (-1,))
    frac = _f64(frac, (-1, 3))
    charge = _f64(charge, (-1,))
    n_k = kvec.shape[0]
    assert kfactor.shape[0] == n_k and charge.shape[0] == frac.shape[0]
    S = np.zeros((n_k, 2), dtype=np.float64)
    h = C.c_void_p()
    _check(load().ceg_ewald_create(C.c_int32(device), _p(invmat9, C.c_double), _p(kmax, C.c_int32), C.c_int64(n_k),
                                   _p(kvec, C.c_int32), _p(kfactor, C.c_double), C.c_int64(frac.shape[0]),
                                   _p(frac, C.c_double), _p(charge, C.c_double), C.byref(h), _p(S, C.c_double)))
    return h, S[:, 0] + 1j * S[:, 1]


def ewald_create_from_sf(device: int, invmat9, kmax, kvec, kfactor, S) -> object:
    """Ewald handle from an existing structure factor S (complex, n_k)."""
    invmat9 = _f64(invmat9, (9,))
    kmax = np.ascontiguousarray(kmax, dtype=np.int32)
    kvec = np.ascontiguousarray(kvec, dtype=np.int32).reshape(-1, 3)
    kfactor = _f64(kfactor, (-1,))
    Sr = np.ascontiguousarray(np.stack([np.real(S), np.imag(S)], axis=1), dtype=np.float64)
    assert kfactor.shape[0] == kvec.shape[0] == Sr.shape[0]
    h = C.c_void_p()
    _check(load().ceg_ewald_create_from_sf(C.c_int32(device), _p(invmat9, C.c_double), _p(kmax, C.c_int32), C.c_int64(kvec.shape[0]),
                                           _p(kvec, C.c_int32), _p(kfactor, C.c_double), _p(Sr, C.c_double), C.byref(h)))
    return h


def ewald_free(h):
    _check(load().ceg_ewald_free(h))


def ewald_reciprocal(h, xyz, q, energy_net_charges: float, static_contribution: float) -> np.ndarray:
    q = _f64(q, (-1,))
    xyz = _f64(xyz, (-1, q.shape[0], 3))
    out = np.empty(xyz.shape[0], dtype=np.float64)
    _check(load().ceg_ewald_reciprocal(h, C.c_int64(xyz.shape[0]), C.c_int32(q.shape[0]), _p(xyz, C.c_double),
                                       _p(q, C.c_double), C.c_double(energy_net_charges),
                                       C.c_double(static_contribution), _p(out, C.c_double)))
    return out


def ewald_single_contribution(h, rest: np.ndarray, xyz, q) -> np.ndarray:
    q = _f64(q, (-1,))
    xyz = _f64(xyz, (-1, q.shape[0], 3))
    r = np.ascontiguousarray(np.stack([np.real(rest), np.imag(rest)], axis=1), dtype=np.float64)
    out = np.empty(xyz.shape[0], dtype=np.float64)
    _check(load().ceg_ewald_single_contribution(h, _p(r, C.c_double), C.c_int64(xyz.shape[0]), C.c_int32(q.shape[0]),
                                                _p(xyz, C.c_double), _p(q, C.c_double), _p(out, C.c_double)))
    return out


def ewald_reciprocal_dev(h, n_pos: int, q, d_xyz: int, energy_net_charges, static_contribution, d_out: int, stream=0):
    q = _f64(q, (-1,))
    _check(load().ceg_ewald_reciprocal_dev(h, C.c_int64(n_pos), C.c_int32(q.shape[0]), C.c_void_p(int(d_xyz)),
                                           _p(q, C.c_double), C.c_double(energy_net_charges),
                                           C.c_double(static_contribution), C.c_void_p(int(d_out)),
                                           C.c_void_p(int(stream))))


# ---- guest-guest pair energies -----------------------------------------------------------------------
def guest_env_create(device: int, cell: ceg_cell_t, cutoff2: float, table: np.ndarray, env_xyz, env_species, env_mol):
    table = np.ascontiguousarray(table)
    assert table.ndim == 2 and table.shape[0] == table.shape[1] and table.dtype.itemsize == 96
    env_xyz = _f64(env_xyz, (-1, 3))
    env_species = np.ascontiguousarray(env_species, dtype=np.int32)
    env_mol = np.ascontiguousarray(env_mol, dtype=np.int32)
    h = C.c_void_p()
    _check(load().ceg_guest_env_create(C.c_int32(device), C.byref(cell), C.c_double(cutoff2), C.c_int32(table.shape[0]),
                                       C.c_void_p(table.ctypes.data), C.c_int64(env_xyz.shape[0]), _p(env_xyz, C.c_double),
                                       _p(env_species, C.c_int32), _p(env_mol, C.c_int32), C.byref(h)))
    return h


def guest_env_free(h):
    _check(load().ceg_guest_env_free(h))


def guest_env_update(h, first: int, xyz):
    xyz = _f64(xyz, (-1, 3))
    _check(load().ceg_guest_env_update(h, C.c_int64(first), C.c_int64(xyz.shape[0]), _p(xyz, C.c_double)))


def guest_env_energy(h, mol_species, skip_mol: int, trial_xyz) -> np.ndarray:
    mol_species = np.ascontiguousarray(mol_species, dtype=np.int32)
    na = mol_species.shape[0]
    trial_xyz = _f64(trial_xyz, (-1, na, 3))
    out = np.empty(trial_xyz.shape[0], dtype=np.float64)
    _check(load().ceg_guest_env_energy(h, C.c_int64(trial_xyz.shape[0]), C.c_int32(na), _p(mol_species, C.c_int32),
                                       C.c_int32(skip_mol), _p(trial_xyz, C.c_double), _p(out, C.c_double)))
    return out



def guest_env_last_kernel_ms(h) -> float:
    t = C.c_double(0.0)
    _check(load().ceg_guest_env_last_kernel_ms(h, C.byref(t)))
    return t.value


def guest_energy(device: int, cell: ceg_cell_t, cutoff2: float, table: np.ndarray, env_xyz, env_species, env_mol,
                 mol_species, skip_mol: int, trial_xyz) -> np.ndarray:
    table = np.ascontiguousarray(table)
    assert table.ndim == 2 and table.shape[0] == table.shape[1] and table.dtype.itemsize == 96
    env_xyz = _f64(env_xyz, (-1, 3))
    env_species = np.ascontiguousarray(env_species, dtype=np.int32)
    env_mol = np.ascontiguousarray(env_mol, dtype=np.int32)
    mol_species = np.ascontiguousarray(mol_species, dtype=np.int32)
    na = mol_species.shape[0]
    trial_xyz = _f64(trial_xyz, (-1, na, 3))
    out = np.empty(trial_xyz.shape[0], dtype=np.float64)
    _check(load().ceg_guest_energy(C.c_int32(device), C.byref(cell), C.c_double(cutoff2), C.c_int32(table.shape[0]),
                                   C.c_void_p(table.ctypes.data), C.c_int64(env_xyz.shape[0]), _p(env_xyz, C.c_double),
                                   _p(env_species, C.c_int32), _p(env_mol, C.c_int32), C.c_int64(trial_xyz.shape[0]),
                                   C.c_int32(na), _p(mol_species, C.c_int32), C.c_int32(skip_mol),
                                   _p(trial_xyz, C.c_double), _p(out, C.c_double)))
    return out


# ---- block mask + fused evaluation -----------------------------------------------------------------
def block_upload(device: int, cell: ceg_cell_t, gs: ceg_gridspec_t, mask: np.ndarray):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    h = C.c_void_p()
    _check(load().ceg_block_upload(C.c_int32(device), C.byref(cell), C.byref(gs), _p(m, C.c_uint8), C.byref(h)))
    return h


def block_from_spheres(device: int, cell: ceg_cell_t, gs: ceg_gridspec_t, centers, radius2, nodes) -> np.ndarray:
    centers = _f64(centers, (-1, 3))
    radius2 = _f64(radius2, (-1,))
    mask = np.empty(nodes, dtype=np.uint8)
    _check(load().ceg_block_from_spheres(C.c_int32(device), C.byref(cell), C.byref(gs), C.c_int32(centers.shape[0]),
                                         _p(centers, C.c_double), _p(radius2, C.c_double), _p(mask, C.c_uint8)))
    return mask.astype(bool)


def block_from_values(device: int, gs: ceg_gridspec_t, values: np.ndarray, threshold: float) -> np.ndarray:
    v = np.ascontiguousarray(values, dtype=np.float32)
    mask = np.empty(v.shape, dtype=np.uint8)
    _check(load().ceg_block_from_values(C.c_int32(device), C.byref(gs), _p(v, C.c_float), C.c_float(threshold),
                                        _p(mask, C.c_uint8)))
    return mask.astype(bool)


def block_free(h):
    _check(load().ceg_block_free(h))


def energy_points(vdw_handles, coulomb, ewald, block, atomsidx, charges, xyz, energy_net_charges, static_contribution):
    atomsidx = np.ascontiguousarray(atomsidx, dtype=np.int32)
    na = atomsidx.shape[0]
    charges = _f64(charges, (na,))
    xyz = _f64(xyz, (-1, na, 3))
    n = xyz.shape[0]
    nv = len(vdw_handles)
    arr = (C.c_void_p * max(nv, 1))(*[h if h is not None else C.c_void_p(0) for h in vdw_handles])
    out_v = np.empty(n, dtype=np.float64)
    out_c = np.empty(n, dtype=np.float64)
    _check(load().ceg_energy_points(C.c_int32(nv), arr, coulomb if coulomb is not None else C.c_void_p(0),
                                    ewald if ewald is not None else C.c_void_p(0),
                                    block if block is not None else C.c_void_p(0), C.c_int64(n), C.c_int32(na),
                                    _p(atomsidx, C.c_int32), _p(charges, C.c_double), _p(xyz, C.c_double),
                                    C.c_double(energy_net_charges), C.c_double(static_contribution),
                                    _p(out_v, C.c_double), _p(out_c, C.c_double)))
    return out_v, out_c


class EnergyPointsCall:
    """ceg_energy_points with everything but the positions marshalled once: a Monte-Carlo loop calls it with the same
    setup at every step, and for one placement the ctypes conversions cost as much as the library call itself."""

    def __init__(self, vdw_handles, coulomb, ewald, block, atomsidx, charges, energy_net_charges, static_contribution):
        self.atomsidx = np.ascontiguousarray(atomsidx, dtype=np.int32)
        self.na = int(self.atomsidx.shape[0])
        self.charges = _f64(charges, (self.na,))
        nv = len(vdw_handles)
        self.arr = (C.c_void_p * max(nv, 1))(*[h if h is not None else C.c_void_p(0) for h in vdw_handles])
        self.f = load().ceg_energy_points
        null = C.c_void_p(0)
        self.head = (C.c_int32(nv), self.arr, coulomb if coulomb is not None else null, ewald if ewald is not None else null,
                     block if block is not None else null)
        self.mid = (C.c_int32(self.na), _p(self.atomsidx, C.c_int32), _p(self.charges, C.c_double))
        self.tail = (C.c_double(energy_net_charges), C.c_double(static_contribution))

    def __call__(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        n = xyz.size // (3 * self.na)
        if n * 3 * self.na != xyz.size:
            raise ValueError("positions must have shape (n_pos, n_atoms, 3)")
        out = np.empty((2, n), dtype=np.float64)
        po = out.ctypes.data
        rc = self.f(*self.head, C.c_int64(n), *self.mid, C.c_void_p(xyz.ctypes.data), *self.tail, C.c_void_p(po),
                    C.c_void_p(po + 8 * n))
        if rc:
            _check(rc)
        return out[0], out[1]


def energy_points_prepared(vdw_handles, coulomb, ewald, block, atomsidx, charges, xyz, energy_net_charges, static_contribution):
    """-> (call, out_vdw, out_coulomb): `call()` re-issues ceg_energy_points on the SAME argument buffers with no Python
    marshalling — what a compiled host (the Julia shim's ccall) pays per call; used to time the library's latency path."""
    atomsidx = np.ascontiguousarray(atomsidx, dtype=np.int32)
    na = atomsidx.shape[0]
    charges = _f64(charges, (na,))
    xyz = _f64(xyz, (-1, na, 3))
    n = xyz.shape[0]
    nv = len(vdw_handles)
    arr = (C.c_void_p * max(nv, 1))(*[h if h is not None else C.c_void_p(0) for h in vdw_handles])
    out_v = np.empty(n, dtype=np.float64)
    out_c = np.empty(n, dtype=np.float64)
    f = load().ceg_energy_points
    args = (C.c_int32(nv), arr, coulomb if coulomb is not None else C.c_void_p(0), ewald if ewald is not None else C.c_void_p(0),
            block if block is not None else C.c_void_p(0), C.c_int64(n), C.c_int32(na), _p(atomsidx, C.c_int32),
            _p(charges, C.c_double), _p(xyz, C.c_double), C.c_double(energy_net_charges), C.c_double(static_contribution),
            _p(out_v, C.c_double), _p(out_c, C.c_double))
    keep = (atomsidx, charges, xyz, arr)

    def call():
        rc = f(*args)
        if rc:
            _check(rc)
        return keep and rc
    return call, out_v, out_c


def energy_lattice(vdw_handles, coulomb, ewald, block, nums, steps, rotpos, atomsidx, charges, energy_net_charges,
                   static_contribution) -> np.ndarray:
    """(n_rot, numA, numB, numC) vdw + coulomb on the lattice of energy_grid, generated on the device."""
    atomsidx = np.ascontiguousarray(atomsidx, dtype=np.int32)
    na = atomsidx.shape[0]
    charges = _f64(charges, (na,))
    rotpos = _f64(rotpos, (-1, na, 3))
    nums = np.ascontiguousarray(nums, dtype=np.int32)
    steps = _f64(steps, (3, 3))
    nv = len(vdw_handles)
    arr = (C.c_void_p * max(nv, 1))(*[h if h is not None else C.c_void_p(0) for h in vdw_handles])
    out = np.empty((rotpos.shape[0], int(nums[0]), int(nums[1]), int(nums[2])), dtype=np.float64)
    _check(load().ceg_energy_lattice(C.c_int32(nv), arr, coulomb if coulomb is not None else C.c_void_p(0),
                                     ewald if ewald is not None else C.c_void_p(0),
                                     block if block is not None else C.c_void_p(0), _p(nums, C.c_int32),
                                     _p(steps, C.c_double), C.c_int32(rotpos.shape[0]), C.c_int32(na),
                                     _p(rotpos, C.c_double), _p(atomsidx, C.c_int32), _p(charges, C.c_double),
                                     C.c_double(energy_net_charges), C.c_double(static_contribution), _p(out, C.c_double)))
    return out
