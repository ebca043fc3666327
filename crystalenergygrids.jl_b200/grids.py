"""Energy grids: construction, RASPA `.grid` I/O, tricubic interpolation, batched evaluation.

Mirrors /root/reference/src/grids.jl (same names, argument meaning and error behaviour):
  EnergyGrid                          grids.jl:10-19
  parse_grid                          grids.jl:61-94
  _create_grid_common (header)        grids.jl:108-116
  create_grid_vdw / create_grid_coulomb   grids.jl:137-185   -> ceg_grid_build
  BlockFile(g)                        grids.jl:188-204
  interpolate_grid                    grids.jl:212-273       -> ceg_interpolate
  CrystalEnergySetup                  grids.jl:284-294
  energy_point / energy_grid          grids.jl:311-424       -> ceg_energy_points

In-memory layout of ``EnergyGrid.grid``: float32 array of shape (8, nx, ny, nz), C order, which is
byte-for-byte the reference's `Array{Cfloat,4}(nz, ny, nx, 8)` (z fastest, channel slowest).
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .cell import CellMatrix, find_supercell, flat, prepare_periodic_distance_computations
from .constants import (BLOCKED_ENERGY, COULOMBIC_CONVERSION_FACTOR, CUTOFF_COULOMB, GRID_TO_KELVIN)
from .coordinates import BlockFile, GridCoordinatesSetup
from .forcefields import ForceField
from .probes import ProbeSystem


PAIR_AUTO_BYTES = 256 << 20      # fused pair records above this size lose to two separate lookups (see CrystalEnergySetup)


def unitcell_struct(cell: CellMatrix):
    ortho, safemin = prepare_periodic_distance_computations(cell.mat)
    return _lib.make_cell(flat(cell.mat), flat(cell.invmat), ortho, safemin * safemin)


class EnergyGrid:
    """grids.jl:10-16.  ewald_precision: -inf zero grid, +inf VdW grid, nan invalid grid.

    `grid` is the host copy (float32, K, shape (8, nx, ny, nz)).  A grid that came out of the device-resident build
    (build_grids_device, setup_RASPA) is usable on the device at once; its host copy and its `.grid` file are
    produced by a background thread, and reading `grid` waits for it."""

    def __init__(self, csetup, num_unitcell, ewald_precision, higherorder, grid, device: int = 0):
        self.csetup = csetup
        self.num_unitcell = num_unitcell
        self.ewald_precision = ewald_precision
        self.higherorder = higherorder
        self._grid = grid
        self.device = device
        self._handle = None
        self._replicas = {}        # device ordinal -> handle of a replica on another GPU
        self._pair = None          # fused {this VdW grid, a Coulomb grid} records
        self._pair_with = None
        self._pending = None       # background thread still writing the file / filling the host copy

    @property
    def grid(self) -> np.ndarray:
        self.wait()
        return self._grid

    @grid.setter
    def grid(self, value):
        self._grid = value

    def wait(self):
        t, self._pending = self._pending, None
        if t is not None:
            t.join()

    @staticmethod
    def zero() -> "EnergyGrid":
        return EnergyGrid(None, (0, 0, 0), -math.inf, False, np.zeros((0, 0, 0, 0), dtype=np.float32))

    @staticmethod
    def invalid() -> "EnergyGrid":
        return EnergyGrid(None, (0, 0, 0), math.nan, False, np.zeros((0, 0, 0, 0), dtype=np.float32))

    @property
    def is_zero(self) -> bool:
        return self.ewald_precision == -math.inf

    @property
    def is_vdw(self) -> bool:
        return self.ewald_precision == math.inf

    @property
    def handle(self):
        """Device-resident copy (node-major re-layout), uploaded on first use."""
        if self._handle is None:
            if self.is_zero or math.isnan(self.ewald_precision):
                return None
            self._handle = _lib.grid_upload(self.device, unitcell_struct(self.csetup.cell), _lib.make_gridspec(self.csetup),
                                            self.grid, self.is_vdw)
        return self._handle

    @staticmethod
    def from_device(csetup, num_unitcell, ewald_precision: float, d_grid: int, device: int = 0, in_kelvin: bool = False):
        """EnergyGrid around a grid that is already on the device in `.grid` order (ceg_grid_build_dev output, or
        slabs gathered from other GPUs): equivalent to writing the file and parse_grid-ing it (grids.jl:61-94) without
        the trip through host memory.  `.grid` stays empty: only device-side evaluation is available."""
        g = EnergyGrid(csetup, tuple(num_unitcell), float(ewald_precision), True, np.zeros((0, 0, 0, 0), dtype=np.float32), device)
        g._handle = _lib.grid_upload_dev(device, unitcell_struct(csetup.cell), _lib.make_gridspec(csetup), d_grid,
                                         1.0 if in_kelvin else GRID_TO_KELVIN, g.is_vdw)
        return g

    def handle_on(self, device: int):
        """Handle of this grid on `device`: the primary one, a replica made by the device-resident build, or a
        replica uploaded from the host copy on first use."""
        if device == self.device:
            return self.handle
        if device not in self._replicas:
            if self.is_zero or math.isnan(self.ewald_precision):
                return None
            if self.grid.size == 0:
                raise _lib.CegError(f"EnergyGrid has no replica on device {device} and no host copy to upload")
            self._replicas[device] = _lib.grid_upload(device, unitcell_struct(self.csetup.cell), _lib.make_gridspec(self.csetup),
                                                      self.grid, self.is_vdw)
        return self._replicas[device]

    def pair_with(self, coulomb: "EnergyGrid"):
        """Fuse this VdW grid with `coulomb` on the device (ceg_grid_pair_create): energy_point(s) and
        interpolate_pair then read 64-byte {VdW, Coulomb} node records with one stencil per atom.  Costs the memory of
        two more grids; results are bit-identical to the separate lookups."""
        if self._pair is not None and self._pair_with is coulomb:
            return self._pair
        self.unpair()
        if not self.is_vdw or coulomb.is_zero or coulomb.is_vdw or self.device != coulomb.device:
            raise ValueError("pair_with expects a VdW grid and a Coulomb grid on the same device")
        self._pair = _lib.grid_pair_create(self.handle, coulomb.handle)
        self._pair_with = coulomb
        return self._pair

    def unpair(self):
        h, self._pair, self._pair_with = self._pair, None, None
        if h is not None:
            _lib.grid_pair_free(h)

    @staticmethod
    def from_device_parts(csetup, num_unitcell, ewald_precision: float, d_parts, slabs, strides=None, device: int = 0,
                          in_kelvin: bool = False):
        """As from_device, from the x-slabs of a sharded build gathered onto this device (part r = planes
        slabs[r] = (i0, i1) in a [8][strides[r]][ny][nz] buffer): the full `.grid`-order array is never assembled."""
        g = EnergyGrid(csetup, tuple(num_unitcell), float(ewald_precision), True, np.zeros((0, 0, 0, 0), dtype=np.float32), device)
        g._handle = _lib.grid_upload_parts_dev(device, unitcell_struct(csetup.cell), _lib.make_gridspec(csetup), d_parts,
                                               [a for a, _ in slabs], [b for _, b in slabs], strides,
                                               1.0 if in_kelvin else GRID_TO_KELVIN, g.is_vdw)
        return g

    def close(self):
        self.wait()
        self.unpair()
        h, self._handle = self._handle, None
        if h is not None:
            _lib.grid_free(h)
        reps, self._replicas = self._replicas, {}
        for r in reps.values():
            _lib.grid_free(r)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- .grid files -------------------------------------------------------------------------------
def _write_header(f, cset: GridCoordinatesSetup, num_unitcell):
    """grids.jl:108-116."""
    f.write(struct.pack("<d", cset.spacing))
    f.write(struct.pack("<3i", *(int(x) for x in cset.dims)))
    for v in (cset.size, cset.shift, cset.delta, cset.unitcell):
        f.write(struct.pack("<3d", *(float(x) for x in v)))
    f.write(struct.pack("<3i", *(int(x) for x in num_unitcell)))


def write_grid(file, cset: GridCoordinatesSetup, num_unitcell, grid: np.ndarray, ewald_precision: float | None = None):
    """grids.jl:151-155, 178-183 — header, [precision], grid, cell matrix (column-major)."""
    with open(file, "wb") as f:
        _write_header(f, cset, num_unitcell)
        if ewald_precision is not None:
            f.write(struct.pack("<d", ewald_precision))
        f.write(np.ascontiguousarray(grid, dtype="<f4").tobytes())
        f.write(flat(cset.cell.mat).astype("<f8").tobytes())


def parse_grid(file, iscoulomb: bool, mat=None, device: int = 0) -> EnergyGrid:
    """grids.jl:61-94."""
    with open(file, "rb") as f:
        spacing, = struct.unpack("<d", f.read(8))
        dims = np.array(struct.unpack("<3i", f.read(12)), dtype=np.int32)
        size = np.array(struct.unpack("<3d", f.read(24)))
        shift = np.array(struct.unpack("<3d", f.read(24)))
        delta = np.array(struct.unpack("<3d", f.read(24)))
        unitcell = np.array(struct.unpack("<3d", f.read(24)))
        num_unitcell = struct.unpack("<3i", f.read(12))
        ewald_precision = struct.unpack("<d", f.read(8))[0] if iscoulomb else math.inf
        n = 8 * int(dims[0] + 1) * int(dims[1] + 1) * int(dims[2] + 1)
        raw = np.fromfile(f, dtype="<f4", count=n)
        if raw.size != n:
            raise EOFError(f"{file}: truncated grid")
        grid = (raw.astype(np.float64) * GRID_TO_KELVIN).astype(np.float32)   # grids.jl:78
        grid = grid.reshape(8, int(dims[0]) + 1, int(dims[1]) + 1, int(dims[2]) + 1)
        if mat is None:
            tail = f.read(72)
            if len(tail) != 72:
                raise ValueError("Missing `mat` argument to `parse_grid` not provided by the grid.")
            mat = np.frombuffer(tail, dtype="<f8").reshape(3, 3).T.copy()
        cell = mat if isinstance(mat, CellMatrix) else CellMatrix.from_mat(mat)
    cs = GridCoordinatesSetup(cell, spacing, dims, size, shift, unitcell, delta)
    return EnergyGrid(cs, tuple(int(x) for x in num_unitcell), float(ewald_precision), True, grid, device)


# ---- construction ------------------------------------------------------------------------------
def vdw_scaling():
    """grids.jl:142-143,148 — (lambda, threshold)."""
    lam_inv = GRID_TO_KELVIN
    return 1.0 / lam_inv, lam_inv * 1e7


def coulomb_scaling():
    """grids.jl:169-170."""
    lam = COULOMBIC_CONVERSION_FACTOR / GRID_TO_KELVIN
    return lam, (1.0 / lam) * 1e7


def _slabs(n_planes: int, n_parts: int):
    """Contiguous x-slabs balanced by plane count."""
    base, rem = divmod(n_planes, n_parts)
    out, s = [], 0
    for p in range(n_parts):
        e = s + base + (1 if p < rem else 0)
        out.append((s, e))
        s = e
    return out


def set_gridpoint(raw: np.ndarray, delta, lam: float, thr: float) -> np.ndarray:
    """grids.jl:118-135 _set_gridpoint! on an array of raw derivatives (..., 8) -> float32 (..., 8)."""
    r = np.array(raw, dtype=np.float64, copy=True)
    over = r[..., 0] > thr                      # also catches Inf; NaN falls through
    r[over, 0] = 2.0 * thr
    r[over, 1:4] = np.clip(r[over, 1:4], -thr, thr)
    r[over, 4:] = 0.0
    d = np.asarray(delta, dtype=np.float64)
    scale = np.array([1.0, d[0], d[1], d[2], d[0] * d[1], d[0] * d[2], d[1] * d[2], (d[0] * d[1]) * d[2]])
    return ((r * scale) * lam).astype(np.float32)


def _build_grids_literal(inputs, cset, outs, i0, i1, device, raw):
    """Slab [i0, i1) through the reference's per-pair algorithm on the device (ceg_compute_derivatives: minimum-image
    search for every supercell atom) — for force fields whose cutoff exceeds half the width of the 12 Å probe supercell
    (probes.jl:24), which the tiled kernel's unique-image argument does not cover."""
    nx, ny, nz = cset.nodes
    lam_v, thr_v = vdw_scaling()
    lam_c, thr_c = coulomb_scaling()
    n_grids = len(outs)
    ps = ProbeSystem(None, None, inputs=inputs, device=device)
    try:
        jj, kk = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
        for i in range(i0, i1):
            # abc_to_xyz (coordinates.jl:72-76): multiply, divide, add
            pos = np.stack([np.full(jj.shape, i * cset.size[0] / cset.dims[0] + cset.shift[0]),
                            jj * cset.size[1] / cset.dims[1] + cset.shift[1],
                            kk * cset.size[2] / cset.dims[2] + cset.shift[2]], axis=-1).reshape(-1, 3)
            der = ps.compute_derivatives(pos)                                  # (n_grids, ny*nz, 8)
            for g in range(n_grids):
                coul = inputs.want_coulomb and g == n_grids - 1
                f32 = set_gridpoint(der[g], cset.delta, lam_c if coul else lam_v, thr_c if coul else thr_v)
                if outs[g] is not None:
                    outs[g][:, i] = f32.reshape(ny, nz, 8).transpose(2, 0, 1)
                if raw is not None:
                    raw[g, :, i - i0] = der[g].reshape(ny, nz, 8).transpose(2, 0, 1)
    finally:
        ps.close()


def build_grids(framework, forcefield: ForceField, spacing: float, atoms=(), coulomb: bool = False, alpha: float = 0.0,
                devices=None, xrange=None, want_raw: bool = False, outs=None):
    """Fused construction of the VdW grids of `atoms` and (optionally) the Coulomb real-space
    grid in one pass over the framework atoms.  Returns (csetup, [grids...], raw) with grids in
    the un-scaled `.grid` units, exactly what create_grid_* return (grids.jl:156,184).

    devices: list of CUDA device ordinals; the x index range is split into contiguous slabs, one
    per device, driven from one host thread each.  xrange=(i0,i1) restricts the build to a slab
    (used by one-process-per-GPU drivers and by the parity tests).
    """
    import threading
    if isinstance(atoms, str):
        atoms = (atoms,)
    cset = GridCoordinatesSetup.from_cell(framework.mat, spacing)
    nx, ny, nz = cset.nodes
    n_grids = len(atoms) + int(coulomb)
    if outs is None:
        outs = [np.empty((8, nx, ny, nz), dtype=np.float32) for _ in range(n_grids)]
    i0, i1 = (0, nx) if xrange is None else xrange
    raw = np.empty((n_grids, 8, i1 - i0, ny, nz), dtype=np.float64) if want_raw else None
    devices = [0] if devices is None else list(devices)
    lam_v, thr_v = vdw_scaling()
    lam_c, thr_c = coulomb_scaling()
    gs = _lib.make_gridspec(cset)
    from .probes import probe_inputs
    inputs = probe_inputs(framework, forcefield, atoms, coulomb, alpha)
    slabs = [(i0 + a, i0 + b) for a, b in _slabs(i1 - i0, len(devices))]
    errors = []

    # the tiled kernel needs 2 * cutoff <= the perpendicular widths of the probe supercell, which is sized for 12 Å
    # (probes.jl:24) whatever the force field's cutoff: larger cutoffs go through the literal kernel (slow, exact)
    from .cell import perpendicular_lengths
    literal = 2.0 * math.sqrt(inputs.cutoff2) > min(perpendicular_lengths(np.asarray(inputs.mat))) * (1.0 + 1e-12)

    def work(dev, a, b):
        try:
            if b <= a:
                return
            if literal:
                r = np.empty((n_grids, 8, b - a, ny, nz), dtype=np.float64) if raw is not None else None
                _build_grids_literal(inputs, cset, outs, a, b, dev, r)
                if raw is not None:
                    raw[:, :, a - i0:b - i0] = r
                return
            ps = ProbeSystem(None, None, inputs=inputs, device=dev)
            try:
                r = None
                if raw is not None:
                    r = np.empty((n_grids, 8, b - a, ny, nz), dtype=np.float64)
                _lib.grid_build(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, outs, r)
                if raw is not None:
                    raw[:, :, a - i0:b - i0] = r
            finally:
                ps.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    if len(devices) == 1:
        work(devices[0], *slabs[0])
    else:
        ths = [threading.Thread(target=work, args=(d, a, b)) for d, (a, b) in zip(devices, slabs)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    if errors:
        raise errors[0]
    return cset, outs, raw


def build_grids_device(framework, forcefield: ForceField, spacing: float, atoms=(), coulomb: bool = False, alpha: float = 0.0,
                       devices=None, ewald_precision: float = 1e-6, download: bool = False, replicate: bool = True):
    """The same fused construction with the result LEFT ON THE DEVICES as interpolable grids: every GPU of `devices`
    builds its x-slab and the build kernel itself writes each node as a finished interpolation record (parse_grid's
    unit conversion applied, grids.jl:78) into the grid handle of EVERY target GPU over NVLink peer stores
    (ceg_grid_build_records) — the hand-off the reference makes by writing `.grid` files and parsing them back
    (raspa.jl:430-438), without the file system, host memory, an all-gather or a re-layout pass.

    -> (csetup, [EnergyGrid ...] (VdW grids of `atoms`, then the Coulomb grid), host) where every EnergyGrid is
    resident on devices[0] with replicas on the other devices (replicate=True), and host = list of float32 arrays in
    `.grid` units (download=True; for writing the files) or None."""
    import threading
    if isinstance(atoms, str):
        atoms = (atoms,)
    devices = [0] if devices is None else list(devices)
    cset = GridCoordinatesSetup.from_cell(framework.mat, spacing)
    nx, ny, nz = cset.nodes
    n_grids = len(atoms) + int(coulomb)
    lam_v, thr_v = vdw_scaling()
    lam_c, thr_c = coulomb_scaling()
    gs = _lib.make_gridspec(cset)
    from .probes import probe_inputs
    inputs = probe_inputs(framework, forcefield, atoms, coulomb, alpha)
    slabs = [(a, b) for a, b in _slabs(nx, len(devices)) if b > a]
    devs = devices[:len(slabs)]
    targets = devices if replicate else devices[:1]
    uc = unitcell_struct(cset.cell)
    handles = [[None] * n_grids for _ in targets]
    host = [np.empty((8, nx, ny, nz), dtype=np.float32) for _ in range(n_grids)] if download else None
    errors = []

    def build(k):
        d_slab = []
        try:
            dev, (a, b) = devs[k], slabs[k]
            ps = ProbeSystem(None, None, inputs=inputs, device=dev)
            try:
                slab_bytes = (b - a) * ny * nz * 4
                if download:
                    d_slab = [_lib.device_malloc(dev, 8 * slab_bytes) for _ in range(n_grids)]
                _lib.grid_build_records(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, dest, layout,
                                        [GRID_TO_KELVIN] * n_grids, d_slab if download else None, 0)
                _lib.device_synchronize(dev)
                for g in range(n_grids if download else 0):
                    for c in range(8):              # the slab of channel c is one contiguous run of the `.grid` array
                        _lib.device_download(dev, d_slab[g] + c * slab_bytes, host[g][c, a:b])
            finally:
                ps.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)
        finally:
            for p in d_slab:
                _lib.device_free(p)

    try:
        for t, dev in enumerate(targets):
            for g in range(n_grids):
                handles[t][g] = _lib.grid_create_empty(dev, uc, gs, g < len(atoms))
        recs = [[_lib.grid_records(h) for h in hs] for hs in handles]
        layout = recs[0][0][2]
        if any(r[2] != layout for rs in recs for r in rs):
            raise _lib.CegError("build_grids_device: the replicas disagree on the record layout")
        dest = [[r[0] for r in rs] for rs in recs]
        if len(devs) == 1:
            build(0)
        else:
            ths = [threading.Thread(target=build, args=(k,)) for k in range(len(devs))]
            [t.start() for t in ths]
            [t.join() for t in ths]
        if errors:
            raise errors[0]
    except BaseException:
        for hs in handles:
            for h in hs:
                if h is not None:
                    _lib.grid_free(h)
        raise
    grids = []
    num_vdw = find_supercell(framework.mat, forcefield.cutoff)
    num_coul = find_supercell(framework.mat, CUTOFF_COULOMB)
    for g in range(n_grids):
        is_vdw = g < len(atoms)
        eg = EnergyGrid(cset, tuple(num_vdw if is_vdw else num_coul), math.inf if is_vdw else float(ewald_precision), True,
                        np.zeros((0, 0, 0, 0), dtype=np.float32), targets[0])
        eg._handle = handles[0][g]
        eg._replicas = {targets[t]: handles[t][g] for t in range(1, len(targets))}
        grids.append(eg)
    return cset, grids, host


class ShardedGridBuild:
    """One process per GPU (torchrun, MPI): the grids of one probe pass built x-slab by x-slab across `world` ranks so
    that EVERY rank ends up with the complete, interpolable grids in its own HBM.  Each rank allocates empty grids
    (ceg_grid_create_empty), the record buffers are exchanged once as CUDA IPC handles, and `build()` launches this
    rank's slab of the build kernel, which writes every node record into all `world` replicas over NVLink
    (ceg_grid_build_records) — construction and hand-off in one kernel; the reference exchanges `.grid` files
    (raspa.jl:430-438).  After build(): synchronise the device, then the ranks (barrier), then interpolate.

    exchange(obj) -> list: all-gather of a picklable object over the ranks, indexed by rank (for torch.distributed:
    `lambda o: (lambda L: (dist.all_gather_object(L, o), L)[1])([None] * world)`).
    precisions: EnergyGrid.ewald_precision per grid (inf for VdW grids, probes in order, Coulomb last); supercell:
    EnergyGrid.num_unitcell, one triple for all grids or one per grid."""

    def __init__(self, inputs, cset, supercell, precisions, rank: int, world: int, exchange, device: int = 0):
        self.cset, self.rank, self.world, self.device = cset, rank, world, device
        self.gs = _lib.make_gridspec(cset)
        self.slabs = _slabs(cset.nodes[0], world)
        self.ps = ProbeSystem(None, None, inputs=inputs, device=device)
        uc = unitcell_struct(cset.cell)
        self.grids = []
        per_grid = len(supercell) == len(precisions) and all(np.ndim(sc) == 1 for sc in supercell)
        for g, prec in enumerate(precisions):
            eg = EnergyGrid(cset, tuple(supercell[g] if per_grid else supercell), float(prec), True,
                            np.zeros((0, 0, 0, 0), dtype=np.float32), device)
            eg._handle = _lib.grid_create_empty(device, uc, self.gs, eg.is_vdw)
            self.grids.append(eg)
        recs = [_lib.grid_records(eg._handle) for eg in self.grids]
        self.layout = recs[0][2]
        mine = [_lib.ipc_export(device, r[0]) + (r[2],) for r in recs] if world > 1 else None
        self._peer_ptrs = []
        self.dest = []
        for r, theirs in enumerate(exchange(mine) if world > 1 else [None]):
            if r == rank:
                self.dest.append([rec[0] for rec in recs])
                continue
            if any(lay != self.layout for _, _, lay in theirs):
                raise _lib.CegError("ShardedGridBuild: the ranks disagree on the record layout")
            ptrs = [_lib.ipc_open(device, hd, off) for hd, off, _ in theirs]
            self._peer_ptrs += ptrs
            self.dest.append(ptrs)
        self.lam = vdw_scaling() + coulomb_scaling()

    def build(self, stream: int = 0, i0: int | None = None, i1: int | None = None):
        """Launch (asynchronously, on `stream`) the build of this rank's x-slab — or of planes [i0, i1) — into every replica."""
        a, b = self.slabs[self.rank]
        a, b = (a if i0 is None else i0), (b if i1 is None else i1)
        if b > a:
            _lib.grid_build_records(self.ps.handle, self.gs, a, b, *self.lam, self.dest, self.layout,
                                    [GRID_TO_KELVIN] * len(self.grids), None, stream)

    def close(self, barrier=None):
        """Unmap the peers' buffers, (barrier: every rank has unmapped), release this rank's probe system.  The grids
        stay valid and are owned by the caller (EnergyGrid.close)."""
        _lib.device_synchronize(self.device)
        ptrs, self._peer_ptrs = self._peer_ptrs, []
        for p in ptrs:
            _lib.ipc_close(self.device, p)
        if barrier is not None:
            barrier()
        ps, self.ps = self.ps, None
        if ps is not None:
            ps.close()


def create_grid_vdw(file, framework, forcefield: ForceField, spacing: float, atom: str, devices=None) -> np.ndarray:
    """grids.jl:137-157."""
    cset, (grid,), _ = build_grids(framework, forcefield, spacing, atoms=(atom,), devices=devices)
    num_unitcell = find_supercell(framework.mat, forcefield.cutoff)
    write_grid(file, cset, num_unitcell, grid)
    return grid


def create_grid_coulomb(file, framework, forcefield: ForceField, spacing: float, _ewald=None, devices=None) -> np.ndarray:
    """grids.jl:159-185."""
    from .ewald import EwaldFramework, initialize_ewald
    num_unitcell = find_supercell(framework.mat, CUTOFF_COULOMB)
    ewald = _ewald if isinstance(_ewald, EwaldFramework) else initialize_ewald(framework, num_unitcell)
    cset, (grid,), _ = build_grids(framework, forcefield, spacing, coulomb=True, alpha=ewald.alpha, devices=devices)
    write_grid(file, cset, num_unitcell, grid, ewald_precision=ewald.precision)
    return grid


# ---- interpolation -----------------------------------------------------------------------------
def interpolate_grid(g: EnergyGrid, point):
    """grids.jl:212-273.  `point`: one Cartesian triplet (returns a float) or an (n,3) array
    (batched method, returns an array).  1e100 marks a blocked VdW position."""
    if g.ewald_precision == -math.inf:
        p = np.asarray(point, dtype=np.float64)
        return 0.0 if p.ndim == 1 else np.zeros(p.reshape(-1, 3).shape[0])
    if math.isnan(g.ewald_precision):
        raise ValueError("Invalid grid cannot be interpolated!")
    p = np.asarray(point, dtype=np.float64)
    out = _lib.interpolate(g.handle, p.reshape(-1, 3))
    return float(out[0]) if p.ndim == 1 else out


def interpolate_pair(vdw: EnergyGrid, coulomb: EnergyGrid, points):
    """interpolate_grid(vdw, p), interpolate_grid(coulomb, p) for an (n,3) array of positions through the fused pair
    records (what energy_point does per atom, grids.jl:316-318) -> (vdw_K[n], coulomb_K[n])."""
    for g in (vdw, coulomb):
        if math.isnan(g.ewald_precision):
            raise ValueError("Invalid grid cannot be interpolated!")
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    out = _lib.interpolate_pair(vdw.pair_with(coulomb), p)
    return out[0], out[1]


def interpolate_sweep(g: EnergyGrid, n: int, seed: int, first: int = 0, n_keep: int = 0):
    """interpolate_grid over `n` pseudo-random positions of the unit cell generated ON the device
    (workloads.random_positions(mat, n, seed, first) bit for bit), reduced there to the sum and minimum of the
    unblocked energies and the blocked count (the 10^9-query sweep).  -> dict."""
    if math.isnan(g.ewald_precision):
        raise ValueError("Invalid grid cannot be interpolated!")
    if g.ewald_precision == -math.inf:
        return {"sum": 0.0, "min": 0.0 if n else math.inf, "n_blocked": 0, "keep": np.zeros(min(n, n_keep)), "kernel_ms": 0.0}
    return _lib.interpolate_sweep(g.handle, first, n, seed, g.csetup.cell.mat, n_keep)


# ---- batched evaluation ------------------------------------------------------------------------
@dataclass
class CrystalEnergySetup:
    """grids.jl:284-294."""
    framework: object
    molecule: object
    coulomb: EnergyGrid
    charges: np.ndarray
    grids: list
    atomsidx: list
    ewald: object
    forcefield: ForceField
    block: BlockFile
    _block_handle: object = None
    _ctx: object = None
    # Evaluate each atom's (VdW, Coulomb) couple through fused 64-byte node records (ceg_grid_pair_create; 2x the grid memory
    # on top of the separate handles): True / False / "auto" = when the fused records of a couple stay below PAIR_AUTO_BYTES.
    # Measured on B200, 10^7 random queries on both grids: 34 MB grids 0.65 vs 0.77 ms, 115 MB 0.80 vs 0.92 ms, but 186 MB
    # 1.69 vs 1.06 ms and 270 MB 2.25 vs 1.26 ms — one launch then sweeps a working set of twice the grid size, and beyond
    # about twice the 126 MB L2 the random-sector rate of the memory system collapses.
    fused_pairs: object = "auto"
    _ep_cache: object = None        # marshalled arguments of the last energy_points configuration

    def block_handle(self, device=0):
        if self.block.empty:
            return None
        if self._block_handle is None:
            cs = self.block.csetup
            self._block_handle = _lib.block_upload(device, unitcell_struct(cs.cell), _lib.make_gridspec(cs), self.block.block)
        return self._block_handle

    def wait(self):
        """Block until the `.grid` files of freshly created grids are on disk and their host copies are filled."""
        for g in list(self.grids) + [self.coulomb]:
            g.wait()

    def ewald_ctx(self):
        if self._ctx is None:
            from .ewald import ewald_context
            self._ctx = ewald_context(self.ewald, ((self.molecule,),))
        return self._ctx

    def __del__(self):
        h, self._block_handle = self._block_handle, None
        if h is not None:
            try:
                _lib.block_free(h)
            except Exception:
                pass


def _energy_args(setup: CrystalEnergySetup, reciprocal: bool):
    for g in setup.grids:
        if math.isnan(g.ewald_precision):
            raise ValueError("Invalid grid cannot be interpolated!")
    vdw_handles = [g.handle for g in setup.grids]
    has_coulomb = setup.coulomb.ewald_precision != -math.inf
    coul = setup.coulomb.handle if has_coulomb else None
    if has_coulomb and setup.fused_pairs:
        for g in setup.grids:      # the library picks the fused records up from the (VdW, Coulomb) couple of handles
            if g.is_vdw and g.csetup is not None and np.array_equal(g.csetup.dims, setup.coulomb.csetup.dims):
                nx, ny, nz = (int(x) for x in g.csetup.nodes)
                if setup.fused_pairs == "auto" and 2 * 32 * nx * ny * nz > PAIR_AUTO_BYTES:
                    continue
                g.pair_with(setup.coulomb)
    ew, enc, static = None, 0.0, 0.0
    if has_coulomb and reciprocal and setup.ewald.alpha != 0:
        ctx = setup.ewald_ctx()
        ew, enc, static = setup.ewald.handle, ctx.energy_net_charges, ctx.static_contribution
    return vdw_handles, coul, ew, setup.block_handle(), enc, static


def energy_points(setup: CrystalEnergySetup, positions, reciprocal: bool = True):
    """Batched energy_point: positions (n_pos, n_atoms, 3) -> (vdw[n_pos], coulomb[n_pos]).
    With reciprocal=False the Coulomb part is only the interpolated real-space grid term, i.e.
    framework_interactions of the Monte-Carlo loop (montecarlo.jl:490-504)."""
    key = (bool(reciprocal), setup.fused_pairs)
    c = setup._ep_cache
    if (c is None or c[0] != key or setup.coulomb._handle is not c[2] or len(c[1]) != len(setup.grids)
            or any(g._handle is not h or g._pair is not p for g, (h, p) in zip(setup.grids, c[1]))):
        # (re)marshal the setup: handles, fused pair records, Ewald constants — once per setup, not once per Monte-Carlo step
        vdw_handles, coul, ew, blk, enc, static = _energy_args(setup, reciprocal)
        call = _lib.EnergyPointsCall(vdw_handles, coul, ew, blk, setup.atomsidx, setup.charges, enc, static)
        c = setup._ep_cache = (key, tuple((g._handle, g._pair) for g in setup.grids), setup.coulomb._handle, call)
    return c[3](positions)


def energy_point(setup: CrystalEnergySetup, positions, ctx=None):
    """grids.jl:311-327 — (vdw, coulomb) in K; (1e100, 0) if an atom sits on a blocking sphere."""
    v, c = energy_points(setup, np.asarray(positions, dtype=np.float64)[None])
    return float(v[0]), float(c[0])


def energy_grid(setup: CrystalEnergySetup, step: float, num_rotate: int = 40, rotations=None):
    """grids.jl:346-424 — energies on a numA x numB x numC lattice along the CELL axes.

    Returns float64 array (n_rot, numA, numB, numC) with vdw+coulomb in K.  The reference draws
    its rotation set from a Lebedev artifact (lebedev.jl, out of scope): here the
    rotation matrices are an input (`rotations`, list of 3x3); monoatomic probes or
    num_rotate == 0 use the identity, exactly as grids.jl:385-390.
    """
    mol_pos = np.asarray(setup.molecule.position, dtype=np.float64).reshape(-1, 3)
    mat = np.asarray(setup.framework.mat, dtype=np.float64)
    axes = [mat[:, 0], mat[:, 1], mat[:, 2]]
    nums = [int(math.floor(math.sqrt(float(a @ a)) / step)) + 1 for a in axes]
    steps = [a / n for a, n in zip(axes, nums)]
    if num_rotate == 0 or len(mol_pos) == 1:
        rotpos = [mol_pos]
    elif rotations is None:
        # grids.jl:385-390 evaluates >= num_rotate Lebedev orientations here; silently using one orientation would
        # return a different array (shape and values) with no error
        raise ValueError("energy_grid of a polyatomic molecule needs `rotations` (the matrices of lebedev.jl's "
                         "get_rotation_matrices(molecule, num_rotate), whose artifact is not part of this package), "
                         "or num_rotate=0 for the single input orientation")
    else:
        rotpos = [mol_pos @ np.asarray(r, dtype=np.float64).T for r in rotations]
    # thisofs = (iA-1)*stepA + (iB-1)*stepB + (iC-1)*stepC, summed left to right (grids.jl:396): generated on the
    # device together with the evaluation (ceg_energy_lattice)
    vdw_handles, coul, ew, blk, enc, static = _energy_args(setup, True)
    return _lib.energy_lattice(vdw_handles, coul, ew, blk, nums, np.array(steps), np.array(rotpos), setup.atomsidx,
                               setup.charges, enc, static)
