"""Shared helpers of the parity tests."""
import numpy as np

import ceg_b200 as ceg

# north_star tolerance: 1e-9 relative or 1e-6 K absolute per grid point (FP64 pre-quantisation values)
RTOL = 1e-9
ATOL_K = 1e-6


def atol_for(is_coulomb: bool) -> float:
    # Coulomb channels are stored in e/Å; 1e-6 K corresponds to 1e-6/CCF there (|q_probe| <= 1)
    return ATOL_K / ceg.COULOMBIC_CONVERSION_FACTOR if is_coulomb else ATOL_K


def assert_close_f64(got, want, is_coulomb, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    # identical non-finite pattern (Inf from hard spheres / <1 Å, NaN from Inf-Inf)
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: NaN masks differ"
    assert np.array_equal(np.isposinf(got), np.isposinf(want)), f"{what}: +Inf masks differ"
    assert np.array_equal(np.isneginf(got), np.isneginf(want)), f"{what}: -Inf masks differ"
    fin = np.isfinite(want)
    err = np.abs(got[fin] - want[fin])
    tol = RTOL * np.abs(want[fin]) + atol_for(is_coulomb)
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} of {bad.size} values out of tolerance; worst abs {err.max():.3e}, "
                           f"worst rel {np.max(err / np.maximum(np.abs(want[fin]), 1e-300)):.3e}")


def assert_f32_grids_match(got, want, raw_want, lam, thr, what=""):
    """Float32 `.grid` arrays: bit-equal except where the FP64 value sits on a rounding boundary.
    Clamped points (value channel == 2*thr*lam) must be the identical set."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape
    clamp_val = np.float32(2 * thr * lam)
    assert np.array_equal(got[0] == clamp_val, want[0] == clamp_val), f"{what}: clamp masks differ"
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: NaN masks differ"
    neq = (got != want) & ~(np.isnan(got) & np.isnan(want))
    if neq.any():
        # every mismatch must be a one-ulp flip of a value within tolerance
        g = got[neq].astype(np.float64)
        w = want[neq].astype(np.float64)
        ulp = np.spacing(np.abs(want[neq])).astype(np.float64)
        assert (np.abs(g - w) <= np.maximum(1.0001 * ulp, ATOL_K)).all(), f"{what}: float32 mismatch larger than one ulp"
        assert neq.mean() < 1e-4, f"{what}: too many float32 rounding flips ({neq.mean():.2e})"


def slab_of(full, i0, i1):
    return full[:, i0:i1]
