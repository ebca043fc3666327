"""Ewald summation, host side: parameters, k-space list, k-factors, context bookkeeping.

Mirrors /root/reference/src/ewald.jl:
  EwaldKspace / EwaldFramework                ewald.jl:28-50
  initialize_ewald                            ewald.jl:195-281  (S_f(k) itself is computed on the
                                              device: ceg_ewald_create, csrc/ewald.cuh)
  EwaldContext ctor (static terms)            ewald.jl:475-541
  compute_ewald                               ewald.jl:555-581  (k-sum on the device:
                                              ceg_ewald_reciprocal)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .cell import flat, inv3, prepare_periodic_distance_computations
from .constants import COULOMBIC_CONVERSION_FACTOR, CUTOFF_COULOMB, nint


@dataclass
class EwaldKspace:
    ks: tuple
    num_kvecs: int
    kindices: list   # (jy, jz, ix_first, ix_last, rangeidx)  — 0-based rangeidx

    def rows_array(self) -> np.ndarray:
        return np.array(self.kindices, dtype=np.int32).reshape(-1, 5)

    def kvec_array(self) -> np.ndarray:
        out = np.empty((self.num_kvecs, 3), dtype=np.int32)
        for (jy, jz, i0, i1, base) in self.kindices:
            n = i1 - i0 + 1
            out[base:base + n, 0] = np.arange(i0, i1 + 1)
            out[base:base + n, 1] = jy
            out[base:base + n, 2] = jz
        return out


def ewald_parameters(precision: float = 1e-6):
    """ewald.jl:198-204 — (alpha, tol1); RASPA's convention, cutoff hard-coded to 12 Å."""
    eps = CUTOFF_COULOMB * min(0.5, abs(precision))
    tol = math.sqrt(abs(math.log(eps)))
    alpha = math.sqrt(abs(math.log(eps * tol))) / CUTOFF_COULOMB
    tol1 = math.sqrt(-math.log(eps * 4.0 * (tol * alpha) ** 2))
    return alpha, tol1


def build_kspace(mat_super: np.ndarray, alpha: float, tol1: float) -> EwaldKspace:
    """ewald.jl:206-236."""
    la, lb, lc = (math.sqrt(float(mat_super[:, c] @ mat_super[:, c])) for c in range(3))
    _a = alpha * tol1 / math.pi
    kx, ky, kz = nint(0.25 + _a * la), nint(0.25 + _a * lb), nint(0.25 + _a * lc)
    recip_cutoff2 = (1.05 * max(kx, ky, kz)) ** 2
    num_kvecs = 0
    kindices = []
    nxt = 0
    for j in range(-ky, ky + 1):
        for k in range(-kz, kz + 1):
            started = False
            first = 1 if (j == 0 and k == 0) else 0
            closed = False
            for i in range(0, kx + 1):
                r2 = i * i + j * j + k * k
                if r2 != 0 and r2 < recip_cutoff2:
                    num_kvecs += 1
                    started = True
                elif started:
                    kindices.append((j, k, first, i - 1, nxt))
                    nxt += (i - 1) - first + 1
                    started = False
                    closed = True
                    break
            if started and not closed:
                kindices.append((j, k, first, kx, nxt))
                nxt += kx - first + 1
    assert nxt == num_kvecs, (nxt, num_kvecs)
    return EwaldKspace((kx, ky, kz), num_kvecs, kindices)


def build_kfactors(kspace: EwaldKspace, mat_super: np.ndarray, invmat: np.ndarray, alpha: float) -> np.ndarray:
    """ewald.jl:238-259."""
    a, b, c = mat_super[:, 0], mat_super[:, 1], mat_super[:, 2]
    det = float(a @ np.cross(b, c))
    volume_factor = COULOMBIC_CONVERSION_FACTOR * (2 * math.pi) / det
    assert volume_factor > 0
    alpha_factor = -0.25 / (alpha * alpha)
    # column-major unpack of invmat (ewald.jl:239)
    il_ax, il_ay, il_az = invmat[0, 0], invmat[1, 0], invmat[2, 0]
    il_bx, il_by, il_bz = invmat[0, 1], invmat[1, 1], invmat[2, 1]
    il_cx, il_cy, il_cz = invmat[0, 2], invmat[1, 2], invmat[2, 2]
    kf = np.empty(kspace.num_kvecs, dtype=np.float64)
    twopi = 2 * math.pi
    for (j, k, i0, i1, base) in kspace.kindices:
        rk0x = j * il_ay + k * il_az
        rk0y = j * il_by + k * il_bz
        rk0z = j * il_cy + k * il_cz
        for I, i in enumerate(range(i0, i1 + 1)):
            rkx = twopi * (rk0x + i * il_ax)
            rky = twopi * (rk0y + i * il_bx)
            rkz = twopi * (rk0z + i * il_cx)
            rksqr = rkx * rkx + rky * rky + rkz * rkz
            kf[base + I] = volume_factor * (1 + (i != 0)) * math.exp(alpha_factor * rksqr) / rksqr
    return kf


@dataclass
class EwaldFramework:
    """ewald.jl:40-50.  ``handle`` is the device-resident copy (k-list, k-factors, S_f)."""
    kspace: EwaldKspace
    alpha: float
    mat: np.ndarray
    invmat: np.ndarray
    kfactors: np.ndarray
    UIon: float
    StoreRigidChargeFramework: np.ndarray    # complex128[num_kvecs]
    net_charges_framework: float
    precision: float
    handle: object = None

    @staticmethod
    def empty(mat) -> "EwaldFramework":
        m = np.array(mat, dtype=np.float64).reshape(3, 3)
        return EwaldFramework(EwaldKspace((0, 0, 0), 0, []), 0.0, m, inv3(m) if np.isfinite(m).all() else m * 0,
                              np.zeros(0), 0.0, np.zeros(0, dtype=np.complex128), 0.0, 0.0)

    def __del__(self):
        h, self.handle = self.handle, None
        if h is not None:
            try:
                _lib.ewald_free(h)
            except Exception:
                pass


def framework_sites(syst, supercell) -> tuple[np.ndarray, np.ndarray]:
    """ewald.jl:123-141, 264 — supercell-fractional coordinates of every image of every atom
    (site index = atom*Π + πa*ΠBC + πb*ΠC + πc) and the matching charges."""
    PA, PB, PC = supercell
    mat_super = syst.mat * np.array(supercell, dtype=np.float64)[None, :]
    invmat = inv3(mat_super)
    n = len(syst)
    frac0 = syst.position @ invmat.T
    pa, pb, pc = np.meshgrid(np.arange(PA), np.arange(PB), np.arange(PC), indexing="ij")
    img = np.stack([pa.ravel() / PA, pb.ravel() / PB, pc.ravel() / PC], axis=1)
    frac = (frac0[:, None, :] + img[None, :, :]).reshape(n * PA * PB * PC, 3)
    charges = np.repeat(np.asarray(syst.atomic_charge, dtype=np.float64), PA * PB * PC)
    return frac, charges


def initialize_ewald(syst, supercell=None, precision: float = 1e-6, device: int = 0) -> EwaldFramework:
    """ewald.jl:195-281."""
    from .cell import find_supercell
    if supercell is None:
        supercell = find_supercell(syst.mat, CUTOFF_COULOMB)
    supercell = tuple(int(x) for x in supercell)
    alpha, tol1 = ewald_parameters(precision)
    mat = np.asarray(syst.mat, dtype=np.float64) * np.array(supercell, dtype=np.float64)[None, :]
    kspace = build_kspace(mat, alpha, tol1)
    invmat = inv3(mat)
    kfactors = build_kfactors(kspace, mat, invmat, alpha)
    UIon = COULOMBIC_CONVERSION_FACTOR * alpha / math.sqrt(math.pi) - float(np.sum(kfactors))
    frac, charges = framework_sites(syst, supercell)
    handle, S = _lib.ewald_create(device, flat(invmat), np.array(kspace.ks, dtype=np.int32), kspace.kvec_array(),
                                  kfactors, frac, charges)
    net = float(np.sum(charges)) if len(charges) else 0.0
    return EwaldFramework(kspace, alpha, mat, invmat, kfactors, UIon, S, net, precision, handle)


@dataclass
class EwaldContext:
    """ewald.jl:321-337 restricted to what compute_ewald needs: the flat list of adsorbate
    charges/positions and the position-independent energy terms."""
    eframework: EwaldFramework
    charges: np.ndarray           # flat, all atoms of all molecules
    positions: np.ndarray         # (n_atoms, 3)
    static_contribution: float
    energy_net_charges: float


def ewald_context(eframework: EwaldFramework, systems) -> EwaldContext:
    """ewald.jl:475-541.  ``systems`` = tuple of kinds, each a list of RASPASystem molecules."""
    from scipy.special import erf as _erf  # SpecialFunctions.erf (ewald.jl:523)
    if eframework.alpha == 0:
        return EwaldContext(eframework, np.zeros(0), np.zeros((0, 3)), 0.0, 0.0)
    allcharges = [np.asarray(kind[0].atomic_charge, dtype=np.float64) for kind in systems]
    numspecies = [len(kind) for kind in systems]
    chargefactor = COULOMBIC_CONVERSION_FACTOR / math.sqrt(math.pi) * eframework.alpha
    energies = [float(np.sum(c * c)) * chargefactor for c in allcharges]
    energy_adsorbate_self = sum(e * n for n, e in zip(numspecies, energies))
    net_charges = [float(np.sum(c)) for c in allcharges]
    total_net = sum(c * n for n, c in zip(numspecies, net_charges))
    ortho, safemin = prepare_periodic_distance_computations(eframework.mat)
    from .coordinates import _periodic_distance2
    energy_excluded = 0.0
    for num, kind, charges in zip(numspecies, systems, allcharges):
        syst = kind[0]
        this_energy = 0.0
        n = len(syst)
        for A in range(n):
            for B in range(A + 1, n):
                buf = (syst.position[B] - syst.position[A])[None, :]
                r = math.sqrt(float(_periodic_distance2(buf, eframework.mat, eframework.invmat, ortho, safemin * safemin)[0]))
                this_energy += float(_erf(eframework.alpha * r)) * charges[A] * charges[B] / r
        energy_excluded += num * this_energy * COULOMBIC_CONVERSION_FACTOR
    static = eframework.UIon * total_net ** 2 - energy_adsorbate_self - energy_excluded
    energy_net_charges = eframework.UIon * eframework.net_charges_framework * total_net
    q = np.concatenate([np.tile(c, n) for c, n in zip(allcharges, numspecies)]) if systems else np.zeros(0)
    pos = np.concatenate([m.position for kind in systems for m in kind]) if systems else np.zeros((0, 3))
    return EwaldContext(eframework, q, pos.reshape(-1, 3), static, energy_net_charges)


def compute_ewald(ctx_or_framework, systems=None, positions=None) -> float | np.ndarray:
    """ewald.jl:555-581.  With ``positions`` (n_pos, n_atoms, 3) evaluates many placements of the
    context's charges in one launch and returns an array."""
    ctx = ctx_or_framework if isinstance(ctx_or_framework, EwaldContext) else ewald_context(ctx_or_framework, systems)
    ef = ctx.eframework
    if ef.alpha == 0:
        return 0.0 if positions is None else np.zeros(len(positions))
    single = positions is None
    pos = ctx.positions[None, :, :] if single else np.asarray(positions, dtype=np.float64).reshape(-1, len(ctx.charges), 3)
    out = _lib.ewald_reciprocal(ef.handle, pos, ctx.charges, ctx.energy_net_charges, ctx.static_contribution)
    return float(out[0]) if single else out


def adsorbate_sums(eframework: EwaldFramework, charges, positions) -> np.ndarray:
    """S_a(k) = sum_a q_a exp(i k.r_a) of a set of point charges (ewald_main_loop!, ewald.jl:148-185) —
    one column of IncrementalEwaldContext.sums.  Host helper for setting up trial evaluations."""
    frac = np.asarray(positions, dtype=np.float64).reshape(-1, 3) @ eframework.invmat.T
    kv = eframework.kspace.kvec_array().astype(np.float64)
    ph = frac @ kv.T                      # (n_atoms, n_k)
    ph -= np.floor(ph)
    return (np.asarray(charges, dtype=np.float64)[:, None] * np.exp(2j * np.pi * ph)).sum(axis=0)


def single_contribution_ewald(eframework: EwaldFramework, rest: np.ndarray, charges, positions) -> np.ndarray:
    """ewald.jl:704-738, batched: contribution of one molecule (charges, n_atoms) at each of the trial
    placements `positions` (n_pos, n_atoms, 3) to the reciprocal Ewald sum, `rest` being the structure
    factor of everything else (framework + other adsorbates).  The difference of two such values is the
    reciprocal-space energy change of a Monte-Carlo move (montecarlo.jl:563-606)."""
    if eframework.alpha == 0:
        return np.zeros(np.asarray(positions).reshape(-1, len(charges), 3).shape[0])
    return _lib.ewald_single_contribution(eframework.handle, np.asarray(rest, dtype=np.complex128), positions, charges)
