"""ceg-b200: B200-native (sm_100a) hot path of CrystalEnergyGrids.jl behind the reference's API.

Import as ``ceg_b200`` (see /ceg_b200.py at the repo root; the directory name carries a dot, which
Python cannot import directly).  The host layer mirrors the reference's Julia entry points
(grids.jl, probes.jl, ewald.jl, forcefields.jl, raspa.jl) and calls hand-written CUDA kernels
through the C ABI declared in include/ceg_cuda.h.  There is no CPU fallback.
"""
from . import _lib
from ._lib import CegError, LIB_PATH
from .cell import (CellMatrix, cell_parameters, find_supercell, flat, inv3, mat_from_parameters,
                   perpendicular_lengths, prepare_periodic_distance_computations)
from .constants import COULOMBIC_CONVERSION_FACTOR, GRID_TO_KELVIN
from .coordinates import (BlockFile, GridCoordinatesSetup, abc_to_xyz, blockfile_from_grid, inverse_offsetpoint,
                          offsetpoint, parse_blockfile, wrap_atom)
from .energy import (GuestEnvironment, compute_vdw, flatten_pair_energy, movement_energies, pair_energy_dtype,
                     pair_energy_table, single_contribution_vdw)
from .ewald import (EwaldContext, EwaldFramework, EwaldKspace, adsorbate_sums, build_kfactors, build_kspace, compute_ewald,
                    single_contribution_ewald,
                    ewald_context, ewald_parameters, framework_sites, initialize_ewald)
from .forcefields import (ForceField, InteractionRule, InteractionRuleSum, Kind, Mixing, UndefinedInteractionError,
                          get_atom_name, needsvdwgrid, pair_dtype, pair_table)
from .grids import (CrystalEnergySetup, EnergyGrid, ShardedGridBuild, build_grids, build_grids_device, coulomb_scaling, create_grid_coulomb,
                    create_grid_vdw, energy_grid, energy_point, energy_points, interpolate_grid, interpolate_pair,
                    interpolate_sweep, parse_grid, vdw_scaling, write_grid)
from .montecarlo import MonteCarloGrids, framework_interactions, setup_montecarlo
from .probes import (ProbeInputs, ProbeSystem, compute_derivatives_ewald, compute_derivatives_vdw, probe_inputs,
                     replicate_framework)
from .raspa import (RASPASystem, getdir_RASPA, grid_locations, hash_forcefield, load_framework_RASPA, load_molecule_RASPA,
                    parse_forcefield_RASPA, parse_pseudoatoms_RASPA, read_cif_P1, retrieve_or_create_grid,
                    setdir_RASPA, setup_RASPA)

__version__ = "0.1.0"
