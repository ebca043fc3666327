"""Example: energy grids for CO2 in a synthetic MFI-shaped framework, an energy map, and a batch of Monte-Carlo
trial placements evaluated on the GPU.  Needs a B200 and the built library
(python -c "import __graft_entry__ as g; g.build()").

    python examples/mfi_co2_trials.py [spacing Å, default 0.2]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

spacing = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2

# --- framework, force field, probe molecule -------------------------------------------------------------------
fw, ff, species = wl.mfi_shaped()                      # 288 atoms, 20.0 x 19.9 x 13.4 Å, Lorentz-Berthelot LJ + charges
geom = np.array([[0, 0, 1.149], [0, 0, 0.0], [0, 0, -1.149]])
q = np.array([species["O_co2"][2], species["C_co2"][2], species["O_co2"][2]])
co2 = ceg.RASPASystem(np.diag([np.inf] * 3), geom, ["O_co2", "C_co2", "O_co2"], np.ones(3), q, True)

# --- one fused pass: the two VdW grids of the CO2 sites + the Ewald real-space grid (grids.jl:137-185) ---------
ewald = ceg.initialize_ewald(fw)                       # k-space, S_f on the device (ewald.jl:195-281)
t0 = time.perf_counter()
cset, raw, _ = ceg.build_grids(fw, ff, spacing, atoms=("O_co2", "C_co2"), coulomb=True, alpha=ewald.alpha)
print(f"3 grids of {cset.nodes} nodes built in {1e3 * (time.perf_counter() - t0):.1f} ms")
to_K = lambda g: (g.astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32)          # what parse_grid does
sc = ceg.find_supercell(fw.mat, 12.0)
grids = [ceg.EnergyGrid(cset, sc, np.inf, True, to_K(raw[0])), ceg.EnergyGrid(cset, sc, np.inf, True, to_K(raw[1]))]
coulomb = ceg.EnergyGrid(cset, sc, 1e-6, True, to_K(raw[2]))
block = ceg.blockfile_from_grid(grids[0])              # BlockFile(g): voxels above 5e6 K (grids.jl:188-204)
setup = ceg.CrystalEnergySetup(fw, co2, coulomb, q, grids, [0, 1, 0], ewald, None, block)

# --- energy map on a 0.5 Å lattice, two orientations (energy_grid, grids.jl:346-424) ---------------------------
rot = [np.eye(3), np.array([[0.0, 0, 1], [0, 1, 0], [-1, 0, 0]])]
t0 = time.perf_counter()
emap = ceg.energy_grid(setup, 0.5, num_rotate=2, rotations=rot)
ok = emap < 1e99
print(f"energy map {emap.shape} in {1e3 * (time.perf_counter() - t0):.1f} ms: {ok.mean():.1%} accessible, "
      f"minimum {emap[ok].min():.1f} K at {np.unravel_index(np.where(ok, emap, np.inf).argmin(), emap.shape)}")

# --- a batch of trial moves of one molecule among 20 others (montecarlo.jl:563-606) ----------------------------
names = ["O_co2", "C_co2"]
table = ceg.pair_energy_table(ff, names)
rng = np.random.default_rng(0)
centers = wl.random_positions(fw.mat, 20, 1)
others = (centers[:, None, :] + geom[None]).reshape(-1, 3)
env = ceg.GuestEnvironment(fw.mat, 12.0, table, others, np.tile(np.array([0, 1, 0], dtype=np.int32), 20),
                           np.repeat(np.arange(20, dtype=np.int32), 3))
rest = ewald.StoreRigidChargeFramework + ceg.adsorbate_sums(ewald, np.tile(q, 19), others[3:])   # framework + the 19 others
trials = wl.random_positions(fw.mat, 200_000, 2)[:, None, :] + geom[None]
t0 = time.perf_counter()
f_vdw, f_coul, recip, guest = ceg.movement_energies(setup, env, rest, [0, 1, 0], trials, skip_molecule=0)
dt = time.perf_counter() - t0
total = f_vdw + f_coul + recip + guest
acc = total < 1e99
print(f"{len(trials)} trial placements in {1e3 * dt:.1f} ms ({len(trials) / dt / 1e6:.1f} M/s): {acc.mean():.1%} not blocked, "
      f"best {total[acc].min():.1f} K = framework {f_vdw[total.argmin()]:.1f} + {f_coul[total.argmin()]:.1f}, "
      f"reciprocal {recip[total.argmin()]:.1f}, guests {guest[total.argmin()]:.1f}")
