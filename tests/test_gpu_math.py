"""Accuracy of the branch-free device math used by the construction kernels (csrc/fastmath.cuh),
measured on the device against mpmath/numpy."""
import numpy as np
import pytest

import ceg_b200 as ceg

pytestmark = pytest.mark.gpu


def test_fastmath_accuracy(cuda_lib):
    from scipy.special import erfcx
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(200000) * 3.5, np.linspace(1e-6, 3.5, 20001), [3.5, 1e-300 * 0 + 1e-12]])
    out = ceg._lib.selftest_math(x)
    assert np.max(np.abs(out[:, 1] - erfcx(x)) / erfcx(x)) < 5e-13          # erfcx_fast on [0, 3.5] (degree 12: 3.2e-13)
    # exp over the range the kernels use it on: [-B r, 0] with B r up to ~700
    y = np.concatenate([rng.random(200000) * 60.0, rng.random(20000) * 690.0, [0.0, 1e-300, 700.0]])
    o2 = ceg._lib.selftest_math(np.maximum(y, 1e-300))
    ref = np.exp(-np.maximum(y, 1e-300))
    assert np.max(np.abs(o2[:, 0] - ref) / ref) < 1e-13                       # exp_neg_fast
    z = np.concatenate([rng.random(200000) * 200.0 + 1e-3, 10.0 ** rng.uniform(-6, 6, 50000)])
    o3 = ceg._lib.selftest_math(z)
    assert np.max(np.abs(o3[:, 2] * np.sqrt(z) - 1.0)) < 1e-14               # rsqrt_fast
    assert np.max(np.abs(o3[:, 3] * z - 1.0)) < 1e-14                        # rcp_fast
