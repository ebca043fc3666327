// Branch-free FP64 transcendentals for the grid-build inner loop (sm_100a).
//
// The CUDA math library's exp/erfcx/rsqrt are accurate to ~1 ulp but cost branches, special-case
// handling and — because their polynomial coefficients are immediates — two UMOV/IMAD.MOV per
// coefficient once registers are short (ncu, round 1: 34 % of all issued instructions).  The
// versions below keep every coefficient in the constant bank (a direct DFMA operand), have no
// branches and target <= 1e-13 relative error, four orders of magnitude inside the 1e-9 parity
// tolerance of north_star.  tests/test_gpu_math.py measures their error on the device.
#pragma once
#include "math_coeffs.cuh"

namespace ceg {

// 1/sqrt(s), s > 0 normal: MUFU.RSQ64H seed (~2^-22) + one third-order step -> ~2^-60.
__device__ __forceinline__ double rsqrt_fast(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double e = fma(-(s * y), y, 1.0);         // 1 - s y^2
    const double t = fma(e, 0.375, 0.5);
    return fma(y, e * t, y);                        // y (1 + e/2 + 3 e^2/8)
}

// 1/d, d > 0 normal: MUFU.RCP64H seed + one third-order step.
__device__ __forceinline__ double rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    return fma(y, fma(e, e, e), y);                 // y (1 + e + e^2)
}

// exp(x) for x <= 0: 2^n * p(f), n = rint(x log2 e), f = x - n ln2.  n is clamped at -1021 on the
// integer pipe, so arguments below -707 return ~2^-1021 (the true value underflows) instead of wrapping.
__device__ __forceinline__ double exp_neg_fast(double x) {
    const double MAGIC = 6755399441055744.0;        // 1.5 * 2^52: low word of (t) holds n
    const double t = fma(x, 1.4426950408889634, MAGIC);
    const double n = t - MAGIC;
    double f = fma(n, -6.93147180369123816490e-01, x);
    f = fma(n, -1.90821492927058770002e-10, f);
    double p = c_exp[EXP_DEG];
#pragma unroll
    for (int k = EXP_DEG - 1; k >= 0; k--) p = fma(p, f, c_exp[k]);
    const int ni = max(__double2loint(t), -1021);
    return __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
}

// erfcx(x) = exp(x^2) erfc(x) for 0 <= x <= ERFCX_XMAX.
__device__ __forceinline__ double erfcx_fast(double x) {
    const double u = rcp_fast(fma(ERFCX_P, x, 1.0));
    const double z = fma(u, ERFCX_ZSCALE, ERFCX_ZSHIFT);
    double p = c_erfcx[ERFCX_DEG];
#pragma unroll
    for (int k = ERFCX_DEG - 1; k >= 0; k--) p = fma(p, z, c_erfcx[k]);
    return p;
}

}  // namespace ceg
