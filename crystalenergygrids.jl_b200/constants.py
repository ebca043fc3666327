"""Physical constants of the hot path, in plain floats (Å, K, e).

Mirrors /root/reference/src/constants.jl:20-21.  The reference builds these from
Unitful/UnitfulAtomic; the doubles below are the values that reproduce the golden
energies of test/runtests.jl:40-56: k_B = 1.380649e-23, u = 1.66053906660e-27,
e = 1.602176634e-19 and eps0 = 1/(mu0 c^2) with mu0 = 4*pi*1e-7.

The device code carries none of these: every unit factor is passed through the C ABI.
"""

# u * Å^2 / ps^2 / k_B / K                              (constants.jl:20)
GRID_TO_KELVIN = 1.202723550011625
# e^2 / (4 pi eps0 Å k_B) in K*Å/e^2                    (constants.jl:21)
COULOMBIC_CONVERSION_FACTOR = 167100.9468073144

# Hard-coded cutoff used for supercell selection and the Ewald parameters
# (probes.jl:24, grids.jl:160, ewald.jl:198).
CUTOFF_COULOMB = 12.0

# interpolate_grid: VdW corner value above which the point is blocked, and the
# returned sentinel (grids.jl:245-248).
VDW_BLOCK_THRESHOLD = 5e6
BLOCKED_ENERGY = 1e100


def nint(x: float) -> int:
    """constants.jl:92 — round half away from zero."""
    import math
    return int(math.floor(x + 0.5 if x >= 0.0 else x - 0.5))
