// Branch-free FP64 transcendentals for the grid-build inner loop (sm_100a).
//
// The CUDA math library's exp/erfcx/rsqrt are accurate to ~1 ulp but cost branches, special-case
// handling and — because their polynomial coefficients are immediates — two UMOV/IMAD.MOV per
// coefficient once registers are short (ncu, round 1: 34 % of all issued instructions).  The
// versions below keep every coefficient in the constant bank (a direct DFMA operand) and have no
// branches.  Error budget (north_star: 1e-9 relative or 1e-6 K per grid point, on SUMS of ~300 pair
// terms that partly cancel): every hot-path routine stays at or below ~1e-12 relative per term, three
// orders inside the tolerance; tests/test_gpu_math.py measures the errors on the device and
// tests/test_gpu_build.py::test_error_budget_* the summed channels against the oracle.
#pragma once
#include "math_coeffs.cuh"

namespace ceg {

// Literal doubles whose low word is not zero cost two moves each inside the loop; as constant-bank values they are
// direct FP64 operands.
__constant__ double c_misc[5] = {1.4426950408889634, -0.6931471805599453, ERFCX_P, ERFCX_ZSCALE, ERFCX_ZSHIFT};

// 1/sqrt(s), s > 0 normal: MUFU.RSQ64H seed (~2^-19) + one third-order step -> ~1e-16.
__device__ __forceinline__ double rsqrt_seed(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    return y;
}
__device__ __forceinline__ double rsqrt_refine(double s, double y) {
    const double e = fma(-(s * y), y, 1.0);         // 1 - s y^2
    const double ye = y * e;                        // independent of t: dependency depth 4, not 5
    const double t = fma(e, 0.375, 0.5);
    return fma(ye, t, y);                           // y (1 + e/2 + 3 e^2/8)
}
__device__ __forceinline__ double rsqrt_fast(double s) { return rsqrt_refine(s, rsqrt_seed(s)); }

// 1/d, d > 0 normal: MUFU.RCP64H seed (measured ~2^-19) + one third-order step -> ~1e-17.
// (A plain Newton step leaves 3e-12, measured — not enough for erfcx_fast.)
__device__ __forceinline__ double rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    return fma(y, fma(e, e, e), y);                 // y (1 + e + e^2)
}

// exp(x) = 2^n p(f), n = rint(x / ln 2), f = x - n ln 2 (|f| <= ln2/2), for -707 < x < 700; no memory access.
// One-constant reduction: |n| (ln 2 - fl(ln 2)) <= 2.4e-14 over the whole range, 2e-15 for |x| <= 60.
// Arguments below ~-707 return ~2^-1021 instead of wrapping.  DEG/coef: c_expp (degree 9, 1.8e-14: the one
// exponential of a Buckingham pair) or c_expa (degree 10, 5e-16: the per-tile Gaussian factors).
#ifndef CEG_POLY_SPLIT
#define CEG_POLY_SPLIT 1   // even/odd Horner: two half-length chains (+1 instruction), the inner loop is latency bound
#endif
template <int DEG, bool SPLIT = true>
__device__ __forceinline__ double horner(double z, const double* __restrict__ coef) {
#if CEG_POLY_SPLIT
  if (SPLIT) {
    const double z2 = z * z;
    double pe = coef[DEG & ~1], po = coef[(DEG - 1) | 1];
#pragma unroll
    for (int k = (DEG & ~1) - 2; k >= 0; k -= 2) pe = fma(pe, z2, coef[k]);
#pragma unroll
    for (int k = ((DEG - 1) | 1) - 2; k >= 1; k -= 2) po = fma(po, z2, coef[k]);
    return fma(po, z, pe);
  }
#endif
    double p = coef[DEG];
#pragma unroll
    for (int k = DEG - 1; k >= 0; k--) p = fma(p, z, coef[k]);
    return p;
}
template <int DEG, bool SPLIT = true>
__device__ __forceinline__ double exp_poly(double x, const double* __restrict__ coef) {
    const double MAGIC = 6755399441055744.0;        // 1.5 * 2^52: low word of (t) holds n
    const double t0 = fma(x, c_misc[0], MAGIC);
    const double n0 = t0 - MAGIC;
    const double f0 = fma(n0, c_misc[1], x);
    const double p0 = horner<DEG, SPLIT>(f0, coef);
    const int ni = max(__double2loint(t0), -1021);
    return __hiloint2double(__double2hiint(p0) + (ni << 20), __double2loint(p0));
}
template <bool SPLIT = true>
__device__ __forceinline__ double exp_fast(double x) { return exp_poly<EXPP_DEG, SPLIT>(x, c_expp); }
__device__ __forceinline__ double exp_acc(double x) { return exp_poly<EXPA_DEG>(x, c_expa); }

// erfcx(x) = exp(x^2) erfc(x) for ERFCX_XMIN <= x <= ERFCX_XMAX (5e-13), in two steps so that the argument map
// (one reciprocal) of the next pair can be computed ahead of its polynomial:
//   z = erfcx_arg(r, alpha * ERFCX_P)  [x = alpha r],   erfcx(x) = erfcx_poly(z).
__device__ __forceinline__ double erfcx_arg(double r, double alpha_p) {
    const double u = rcp_fast(fma(r, alpha_p, 1.0));
    return fma(u, c_misc[3], c_misc[4]);
}
template <bool SPLIT = true>
__device__ __forceinline__ double erfcx_poly(double z) { return horner<ERFCX_DEG, SPLIT>(z, c_erfcx); }
__device__ __forceinline__ double erfcx_fast(double x) { return erfcx_poly<true>(erfcx_arg(x, c_misc[2])); }

}  // namespace ceg
