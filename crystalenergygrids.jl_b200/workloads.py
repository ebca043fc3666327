"""Named workloads of BASELINE.json / SURVEY.md §8(d): real fixtures (flattened, tests/golden/*.npz)
and seeded synthetic frameworks.  Used by bench.py, __graft_entry__.smoke() and tests; pure host code.

RNG: splitmix64 on (seed, index) -> uniform [0,1) doubles, so the same streams can be regenerated
anywhere (including on the device).
"""
from __future__ import annotations

import math
import os

import numpy as np

from .cell import CellMatrix, find_supercell, inv3, prepare_periodic_distance_computations
from .coordinates import GridCoordinatesSetup
from .forcefields import (ForceField, InteractionRule, Kind, LennardJones, Mixing, NoInteraction, make_rule,
                          pair_dtype)
from .probes import ProbeInputs, probe_inputs
from .raspa import RASPASystem

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

_MASK = (1 << 64) - 1


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """Uniform [0,1) doubles from splitmix64(seed * 0x9E3779B97F4A7C15 + idx)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + idx.astype(np.uint64) + np.uint64(1)) \
            * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def random_positions(mat: np.ndarray, n: int, seed: int, start: int = 0) -> np.ndarray:
    """C2: positions mat·(u1,u2,u3), u from splitmix64(seed, 3*i + d)."""
    idx = np.arange(start * 3, (start + n) * 3, dtype=np.uint64)
    u = splitmix64(seed, idx).reshape(n, 3)
    m = np.asarray(mat, dtype=np.float64)
    # (m0 u0 + m1 u1) + m2 u2 with every product and sum rounded: the device generator of
    # ceg_interpolate_sweep produces the same bits
    return (u[:, 0:1] * m[:, 0] + u[:, 1:2] * m[:, 1]) + u[:, 2:3] * m[:, 2]


# ---- fixtures ------------------------------------------------------------------------------------
def load_fixture(name: str):
    """-> (ProbeInputs, GridCoordinatesSetup, dict of extras) from tests/golden/<name>.npz."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    table = np.ascontiguousarray(z["table"]).view(pair_dtype).reshape(z["table"].shape[:-1])
    charges = z["charges"] if z["charges"].size else None
    inp = ProbeInputs(z["mat"], z["invmat"], bool(z["ortho"]), float(z["safemin2"]), z["positions"],
                      z["types"].astype(np.int32), charges, int(z["n_types"]), table, float(z["cutoff2"]),
                      bool(z["want_coulomb"]), float(z["alpha"]), tuple(int(x) for x in z["supercell"]))
    cset = GridCoordinatesSetup.from_cell(z["unit_mat"], float(z["spacing"]))
    extras = {k: z[k] for k in z.files}
    return inp, cset, extras


# ---- synthetic frameworks ---------------------------------------------------------------------------
def _jittered_lattice(mat: np.ndarray, counts, jitter: float, min_dist: float, seed: int, n_target: int):
    """Atoms on a jittered lattice inside an orthorhombic cell; greedy rejection of close pairs."""
    L = np.array([mat[0, 0], mat[1, 1], mat[2, 2]])
    nx, ny, nz = counts
    pts = []
    idx = 0
    cell = max(min_dist, 1e-9)
    ncell = np.maximum(1, np.floor(L / cell).astype(int))
    buckets = {}

    def ok(p):
        c = np.floor(p / L * ncell).astype(int) % ncell
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    key = ((c[0] + dx) % ncell[0], (c[1] + dy) % ncell[1], (c[2] + dz) % ncell[2])
                    for q in buckets.get(key, ()):
                        d = p - q
                        d -= L * np.round(d / L)
                        if d @ d < min_dist * min_dist:
                            return False
        return True

    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                if len(pts) >= n_target:
                    break
                u = splitmix64(seed, np.arange(idx * 3, idx * 3 + 3, dtype=np.uint64))
                idx += 1
                p = (np.array([i + 0.5, j + 0.5, k + 0.5]) / np.array([nx, ny, nz])) * L + (u - 0.5) * 2 * jitter
                p = p % L
                if ok(p):
                    pts.append(p)
                    c = tuple(np.floor(p / L * ncell).astype(int) % ncell)
                    buckets.setdefault(c, []).append(p)
    return np.array(pts)


def _lj_forcefield(species: dict, cutoff=12.0) -> ForceField:
    """Lorentz–Berthelot LJ force field, truncated at the cutoff."""
    names = list(species)
    inputs = [((s, s), make_rule(Kind.LennardJones, species[s][:2], False, cutoff, True)) for s in names]
    return ForceField.build(inputs, Mixing.LorentzBerthelot, cutoff, False, True, sdict={s: i for i, s in enumerate(names)},
                            name="synthetic-LJ")


def mfi_shaped():
    """C3: orthorhombic 20.022 x 19.899 x 13.383 Å, 96 "Si" + 192 "O" on a jittered lattice (seed 2)."""
    mat = np.diag([20.022, 19.899, 13.383])
    pts = _jittered_lattice(mat, (8, 8, 5), 0.25, 1.4, 2, 288)
    n = len(pts)
    symbols = ["Si" if (i % 3 == 0) else "O" for i in range(n)]
    n_si = sum(s == "Si" for s in symbols)
    n_o = n - n_si
    q_si = 2.05
    q_o = -q_si * n_si / n_o
    species = {"O": (53.0, 3.30, q_o), "Si": (22.0, 2.30, q_si), "C_ch4": (148.0, 3.73, 0.0),
               "C_co2": (27.0, 2.80, 0.70), "O_co2": (79.0, 3.05, -0.35)}
    ff = _lj_forcefield(species)
    charges = np.array([species[s][2] for s in symbols])
    fw = RASPASystem(mat, pts, symbols, np.ones(n), charges, False)
    return fw, ff, species


def mof_shaped(n_atoms: int = 5000, edge: float = 50.0):
    """C5: cubic 50 Å cell, ~5000 atoms of 4 types on a jittered lattice (seed 3), neutral."""
    mat = np.diag([edge, edge, edge])
    m = int(math.ceil(n_atoms ** (1 / 3)))
    pts = _jittered_lattice(mat, (m, m, m), 0.6, 1.2, 3, n_atoms)
    n = len(pts)
    kinds = ["C", "H", "O", "Zn"]
    symbols = [kinds[i % 4] for i in range(n)]
    q = {"C": 0.25, "H": 0.15, "O": -0.70, "Zn": 0.30}
    species = {"C": (47.9, 3.47, q["C"]), "H": (7.65, 2.85, q["H"]), "O": (48.2, 3.03, q["O"]), "Zn": (27.7, 4.04, q["Zn"]),
               "C_ch4": (148.0, 3.73, 0.0), "Na": (15.1, 2.66, 1.0)}
    charges = np.array([q[s] for s in symbols])
    charges -= charges.mean()
    ff = _lj_forcefield(species)
    fw = RASPASystem(mat, pts, symbols, np.ones(n), charges, False)
    return fw, ff, species


def rhombo60_shaped():
    """Edge case of the minimum-image search (utils.jl:226-246): rhombohedral 10.2 Å cell with 60° angles.
    Its (3,3,3) supercell has safemin = 10.82 Å < 12 Å cutoff, so pairs with safemin^2 < d^2 < cutoff^2 go
    through the reference's +-1 image search, including its stale-vector exit."""
    from .cell import mat_from_parameters
    mat = mat_from_parameters((10.2, 10.2, 10.2), (60.0, 60.0, 60.0))
    pts = _jittered_lattice(mat, (3, 3, 3), 0.3, 1.4, 5, 27)
    n = len(pts)
    symbols = ["Si" if (i % 3 == 0) else "O" for i in range(n)]
    n_si = sum(s == "Si" for s in symbols)
    q_si = 2.05
    q_o = -q_si * n_si / (n - n_si)
    species = {"O": (53.0, 3.30, q_o), "Si": (22.0, 2.30, q_si), "C_ch4": (148.0, 3.73, 0.0), "Na": (15.1, 2.66, 1.0)}
    ff = _lj_forcefield(species)
    charges = np.array([species[s][2] for s in symbols])
    fw = RASPASystem(mat, pts, symbols, np.ones(n), charges, False)
    return fw, ff, species


def synthetic_inputs(kind: str, spacing: float, probes, coulomb: bool):
    """-> (ProbeInputs, GridCoordinatesSetup, framework)."""
    from .ewald import ewald_parameters
    fw, ff, _ = {"mfi": mfi_shaped, "mof": mof_shaped, "rhombo60": rhombo60_shaped}[kind]()
    alpha = ewald_parameters(1e-6)[0] if coulomb else 0.0
    inp = probe_inputs(fw, ff, tuple(probes), coulomb, alpha)
    cset = GridCoordinatesSetup.from_cell(fw.mat, spacing)
    return inp, cset, fw


def respace(cset: GridCoordinatesSetup, spacing: float) -> GridCoordinatesSetup:
    return GridCoordinatesSetup.from_cell(cset.cell, spacing)
