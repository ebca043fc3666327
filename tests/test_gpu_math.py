"""Accuracy of the branch-free device math used by the construction kernels (csrc/fastmath.cuh),
measured on the device against mpmath/numpy."""
import mpmath as mp
import numpy as np
import pytest

import ceg_b200 as ceg

pytestmark = pytest.mark.gpu


def test_fastmath_accuracy(cuda_lib):
    from scipy.special import erfcx
    rng = np.random.default_rng(0)
    x = np.concatenate([0.25 + rng.random(200000) * 3.0, np.linspace(0.25, 3.25, 20001)])
    out = ceg._lib.selftest_math(x)
    # erfcx_fast on its fitted range [0.25, 3.25] = [alpha * 1 A, alpha * cutoff] with margin (degree 11: 5.1e-13)
    assert np.max(np.abs(out[:, 1] - erfcx(x)) / erfcx(x)) < 8e-13
    # exp over the range the kernels use it on: [-B r, ln A] with B r up to ~700
    y = np.concatenate([rng.random(200000) * 60.0, rng.random(20000) * 690.0, [0.0, 1e-300, 700.0]])
    ym = np.maximum(y, 1e-300)
    o2 = ceg._lib.selftest_math(ym)
    ref = np.array([float(mp.exp(-mp.mpf(float(v)))) for v in ym[::40]])
    assert np.max(np.abs(o2[::40, 0] - ref) / ref) < 1e-13                     # exp_fast (degree 9: 1.8e-14)
    assert np.max(np.abs(o2[::40, 4] - ref) / ref) < 1e-13                     # exp_acc (degree 10)
    assert np.max(np.abs(o2[:, 0] - np.exp(-ym)) / np.exp(-ym)) < 1e-13
    assert np.max(np.abs(o2[:, 4] - np.exp(-ym)) / np.exp(-ym)) < 1e-13
    pos = ym[ym < 600.0]
    o2p = ceg._lib.selftest_math(pos)
    assert np.max(np.abs(o2p[:, 5] - np.exp(pos)) / np.exp(pos)) < 1e-13       # exp_fast, positive arguments (ln A folded in)
    z = np.concatenate([rng.random(200000) * 200.0 + 1e-3, 10.0 ** rng.uniform(-6, 6, 50000)])
    o3 = ceg._lib.selftest_math(z)
    assert np.max(np.abs(o3[:, 2] * np.sqrt(z) - 1.0)) < 1e-14               # rsqrt_fast
    assert np.max(np.abs(o3[:, 3] * z - 1.0)) < 1e-14                        # rcp_fast
