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

// 1/sqrt(s), s > 0 normal: MUFU.RSQ64H seed (~2^-19) + one third-order step -> ~1e-16.
__device__ __forceinline__ double rsqrt_fast(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double e = fma(-(s * y), y, 1.0);         // 1 - s y^2
    const double ye = y * e;                        // independent of t: dependency depth 4, not 5
    const double t = fma(e, 0.375, 0.5);
    return fma(ye, t, y);                           // y (1 + e/2 + 3 e^2/8)
}

// 1/d, d > 0 normal: MUFU.RCP64H seed (measured ~2^-19) + one third-order step -> ~1e-17.
// (A plain Newton step leaves 3e-12, measured — not enough for erfcx_fast.)
__device__ __forceinline__ double rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    return fma(y, fma(e, e, e), y);                 // y (1 + e + e^2)
}

// exp(x) for x <= 0: 2^k * T[j] * p(f) with N = rint(x * EXP_TAB / ln 2) = EXP_TAB k + j,
// f = x - N ln2/EXP_TAB (|f| <= 1.4e-3, cubic p), T[j] = 2^(j/EXP_TAB) read from a shared-memory copy
// of c_exp_tab.  8 FP64 instructions; the table look-up and the exponent arithmetic run on the LSU
// and integer pipes.  N is clamped so that arguments below ~-707 return ~2^-1021 instead of wrapping.
#ifndef CEG_EXP_TABLE
#define CEG_EXP_TABLE 0
#endif
#ifndef CEG_ESTRIN
#define CEG_ESTRIN 0   // measured slower: the build kernel is FP64-issue bound, not latency bound
#endif
__device__ __forceinline__ double exp_neg_fast(double x, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;        // 1.5 * 2^52: low word of (t) holds N
#if !CEG_EXP_TABLE
    // polynomial-only variant: 2^n * p9(f), |f| <= ln2/2 (12 FP64 instructions, no memory access).
    // Measured faster than the table variant in the build kernel, which is latency- not issue-bound.
    (void)tab;
    const double t0 = fma(x, 1.4426950408889634, MAGIC);
    const double n0 = t0 - MAGIC;
    // one-constant reduction: |n0| (ln 2 - fl(ln 2)) <= 2.4e-14 over the whole range, 2e-15 for x >= -60
    const double f0 = fma(n0, -0.6931471805599453, x);
#if CEG_ESTRIN
    // Estrin's scheme: 12 instructions instead of 9, dependency depth 4 instead of 9
    static_assert(EXPP_DEG == 9, "Estrin form written for degree 9");
    const double f2 = f0 * f0, f4 = f2 * f2, f8 = f4 * f4;
    const double p01 = fma(c_expp[1], f0, c_expp[0]), p23 = fma(c_expp[3], f0, c_expp[2]);
    const double p45 = fma(c_expp[5], f0, c_expp[4]), p67 = fma(c_expp[7], f0, c_expp[6]);
    const double p89 = fma(c_expp[9], f0, c_expp[8]);
    const double qa = fma(f2, p23, p01), qb = fma(f2, p67, p45);
    const double p0 = fma(f8, p89, fma(f4, qb, qa));
#else
    double p0 = c_expp[EXPP_DEG];
#pragma unroll
    for (int k = EXPP_DEG - 1; k >= 0; k--) p0 = fma(p0, f0, c_expp[k]);
#endif
    const int ni = max(__double2loint(t0), -1021);
    return __hiloint2double(__double2hiint(p0) + (ni << 20), __double2loint(p0));
#else
    const double t = fma(x, EXP_INV_STEP, MAGIC);
    const double n = t - MAGIC;
    double f = fma(n, -EXP_STEP_HI, x);
    f = fma(n, -EXP_STEP_LO, f);
    double p = c_exp[EXP_DEG];
#pragma unroll
    for (int k = EXP_DEG - 1; k >= 0; k--) p = fma(p, f, c_exp[k]);
    const int N = max(__double2loint(t), -1021 * EXP_TAB);
    const double r = tab[N & (EXP_TAB - 1)] * p;
    return __hiloint2double(__double2hiint(r) + ((N >> 8) << 20), __double2loint(r));
#endif
}
static_assert(EXP_TAB == 256, "exp_neg_fast shifts by log2(EXP_TAB) = 8");

// erfcx(x) = exp(x^2) erfc(x) for 0 <= x <= ERFCX_XMAX.
__device__ __forceinline__ double erfcx_fast(double x) {
    const double u = rcp_fast(fma(ERFCX_P, x, 1.0));
    const double z = fma(u, ERFCX_ZSCALE, ERFCX_ZSHIFT);
#if CEG_ESTRIN
    static_assert(ERFCX_DEG == 12, "Estrin form written for degree 12");
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double p01 = fma(c_erfcx[1], z, c_erfcx[0]), p23 = fma(c_erfcx[3], z, c_erfcx[2]);
    const double p45 = fma(c_erfcx[5], z, c_erfcx[4]), p67 = fma(c_erfcx[7], z, c_erfcx[6]);
    const double p89 = fma(c_erfcx[9], z, c_erfcx[8]), pab = fma(c_erfcx[11], z, c_erfcx[10]);
    const double q0 = fma(z2, p23, p01), q1 = fma(z2, p67, p45), q2 = fma(z2, pab, p89);
    const double r0 = fma(z4, q1, q0), r1 = fma(z4, c_erfcx[12], q2);
    return fma(z8, r1, r0);
#else
    double p = c_erfcx[ERFCX_DEG];
#pragma unroll
    for (int k = ERFCX_DEG - 1; k >= 0; k--) p = fma(p, z, c_erfcx[k]);
    return p;
#endif
}

}  // namespace ceg
