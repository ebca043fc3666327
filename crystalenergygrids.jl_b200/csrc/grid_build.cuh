// Energy-grid construction kernels (FP64, sm_100a).
//
// Replaces the per-point loops of the reference:
//   compute_derivatives_vdw / compute_derivatives_ewald      probes.jl:71-117
//   derivativesGrid (LJ / Buckingham / HardSphere, sums)     interactions.jl:432-472, 599-610
//   derivatives_ewald                                        ewald.jl:299-312
//   _set_gridpoint! and the create_grid_* loops              grids.jl:118-185
//
// Two kernels:
//  * derivatives_literal_kernel — arbitrary positions, the reference's per-pair algorithm verbatim
//    (min-image search of utils.jl:210-246 for every supercell atom, division-based formulas).
//    Backs ceg_compute_derivatives; its device functions are also the "exact" path of the tiled
//    kernel.
//  * grid_build_tiled_kernel — the production kernel.  One CTA owns a TXxTYxTZ tile (4x8x8) of grid
//    nodes; it scans the pre-expanded list of periodic atom IMAGES (atom-major order, built on the
//    host for the slab's bounding region), keeps those within cutoff + tile radius of the tile
//    centre with an ORDERED compaction into shared memory, and every thread (one node each) walks
//    that list, broadcast from shared memory.  All VdW probes and the Coulomb grid share the one
//    distance computation.  Because the probe supercell has perpendicular widths >= 2*cutoff
//    (probes.jl:24), the in-cutoff image of an atom is unique and equals the reference's wrapped
//    image, so the in-cutoff set is identical; pairs whose d^2 lands within a guard band of a mask
//    threshold (cutoff^2, 1 Å^2, hard-sphere radius^2), that are closer than 0.1 Å, or that would
//    enter the reference's non-orthogonal image search are re-evaluated with the literal algorithm
//    and formulas, which makes every mask decision (and every Inf/NaN) bit-identical to the oracle's.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace ceg {

struct PairDev {
    int flags;
    int _pad;
    double hs_r2;
    double lj_eps, lj_sig2, lj_shift;
    double bA, bB, bC, b_shift;
    double bB2;        // B^2
    double lj_c6, lj_c12;   // 4 eps sigma^6, 4 eps sigma^12 (hot path)
};
// What the hot path reads of a PairDev (the shared-memory copy; the literal path reads the global table)
struct PairHot {
    int flags, _pad;
    double lj_c6, lj_c12, lj_shift;
    double bA, bB, bC, b_shift, bB2;
    __device__ PairHot& operator=(const PairDev& d) {
        flags = d.flags; _pad = 0;
        lj_c6 = d.lj_c6; lj_c12 = d.lj_c12; lj_shift = d.lj_shift;
        bA = d.bA; bB = d.bB; bC = d.bC; b_shift = d.b_shift; bB2 = d.bB2;
        return *this;
    }
};

#ifndef CEG_COMPACT_WARPS
#define CEG_COMPACT_WARPS 1
#endif
constexpr int MAX_GRIDS = 5;       // up to 4 fused VdW probes + Coulomb per pass
constexpr int LIST_CAP = 1024;     // shared-memory image list capacity per round
constexpr int TABLE_SMEM_TYPES = 24;
constexpr double GUARD_BAND = 1e-9;  // Å^2; fast-vs-literal d^2 differ by ~1e-13
constexpr double NEAR_R2 = 0.01;     // below (0.1 Å)^2 always take the literal formulas (r -> 0 semantics)

// ---- literal pair formulas (reference operation order; used off the hot path) -----------------------
__device__ __noinline__ void vdw_pair_ref(const PairDev& p, double r2, double& v, double& p1, double& p2, double& p3) {
    v = 0.0; p1 = 0.0; p2 = 0.0; p3 = 0.0;
    const int fl = p.flags;
    if (fl & CEG_PAIR_HS) v += (r2 < p.hs_r2) ? CUDART_INF : 0.0;
    if (fl & CEG_PAIR_LJ) {                       // interactions.jl:434-441
        const double eps = p.lj_eps;
        const double t = p.lj_sig2 / r2;
        const double x6 = t * t * t;
        const double r4 = r2 * r2;
        v += 4.0 * eps * x6 * (x6 - 1.0) - p.lj_shift;
        p1 += 24.0 * eps * (x6 * (1.0 - 2.0 * x6)) / r2;
        p2 += 96.0 * eps * (x6 * (7.0 * x6 - 2.0)) / r4;
        p3 += 384.0 * eps * (x6 * (5.0 - 28.0 * x6)) / (r4 * r4);
    }
    if (fl & CEG_PAIR_BUCK) {                     // interactions.jl:447-457
        const double A = p.bA, B = p.bB, C = p.bC;
        const double r4 = r2 * r2;
        const double r = sqrt(r2);
        const double r6 = r4 * r2;
        const double x6 = C / r6;
        const double xe = A * exp(-B * r);
        v += (xe - x6) - p.b_shift;
        p1 += -B * xe / r + 6.0 * x6 / r2;
        p2 += -48.0 * x6 / r4 + B * xe * (1.0 + B * r) / (r2 * r);
        p3 += -(3.0 * B * r + B * B * r2 + 3.0) * B * xe * r / r6 + 480.0 * C / (r6 * r6);
    }
}

__device__ __noinline__ void ewald_pair_ref(double alpha, double q, double r2, double& v, double& p1, double& p2,
                                            double& p3) {   // ewald.jl:299-312
    const double r = sqrt(r2);
    const double r3 = r2 * r, r5 = r3 * r2;
    const double a2 = r2 * (alpha * alpha);
    const double e2 = 2.0 * alpha * r * exp(-a2) / sqrt(CUDART_PI);
    const double c = erfc(alpha * r);
    v = q * c / r;
    p1 = -q * (e2 + c) / r3;
    p2 = q * (e2 * (3.0 + 2.0 * a2) + 3.0 * c) / r5;
    p3 = q * (-e2 * (15.0 + 10.0 * a2 + 4.0 * (a2 * a2)) - 15.0 * c) / (r5 * r2);
}

// ---- hot-path pair formulas ---------------------------------------------------------------------------
// D = (1/r) d/dr.  Reciprocal powers come from one rsqrt, exp/erfcx are the branch-free versions of
// fastmath.cuh; differences from the literal formulas are ~1e-13 relative.
struct PairGeom {
    double r2, ir, ir2, r;
};
__device__ __forceinline__ PairGeom make_geom(double r2) {
    PairGeom g;
    g.r2 = r2;
    g.ir = rsqrt_fast(r2);
    g.ir2 = g.ir * g.ir;
    g.r = r2 * g.ir;
    return g;
}

// Lennard-Jones terms (interactions.jl:434-441), no branches.  With a = 4 eps sigma^6 / r^6 and
// b = 4 eps sigma^12 / r^12 the reference's four outputs are
//   U = b - a - shift,  DU = (6/r^2)(a - 2b),  D^2U = (48/r^4)(3.5 b - a),  D^3U = (480/r^8)(a - 5.6 b);
// the powers of 1/r^2 are shared by every probe of the pass (7 multiplications per pair), each probe then
// costs 10 FP64 instructions, none with three register operands.
struct LJPowers {
    double u3, u6, k1, k2, k3;   // r^-6, r^-12, 6 r^-2, 48 r^-4, 480 r^-8
};
__device__ __forceinline__ LJPowers make_lj_powers(const PairGeom& g) {
    LJPowers w;
    const double u2 = g.ir2 * g.ir2;
    w.u3 = u2 * g.ir2;
    w.u6 = w.u3 * w.u3;
    w.k1 = 6.0 * g.ir2;
    w.k2 = 48.0 * u2;
    w.k3 = 480.0 * (u2 * u2);
    return w;
}
__device__ __forceinline__ void lj_terms(const PairHot& p, const LJPowers& w, double& v, double& p1, double& p2,
                                         double& p3) {
    const double a = p.lj_c6 * w.u3;
    const double b = p.lj_c12 * w.u6;
    v = (b - a) - p.lj_shift;
    p1 = w.k1 * fma(b, -2.0, a);
    p2 = w.k2 * fma(b, 3.5, -a);
    p3 = w.k3 * fma(b, -5.6, a);
}

// Buckingham terms (interactions.jl:447-457), no branches.  With X = A e^{-Br}, y = B X / r:
//   D X = -y,  D^2 X = (y + B^2 X)/r^2,  D^3 X = -(3 D^2 X + B^2 y)/r^2   (D = (1/r) d/dr),
// and the dispersion part -C/r^6 contributes 6, -48, 480 times C/r^8, C/r^10, C/r^12.
__device__ __forceinline__ void buck_terms(const PairHot& p, const PairGeom& g, const double* etab, double& v,
                                           double& p1, double& p2, double& p3) {
    const double B = p.bB;
    const double X = p.bA * exp_neg_fast(-(B * g.r), etab);
    const double BX = B * X;
    const double y = BX * g.ir;
    const double d2x = fma(B, BX, y) * g.ir2;
    const double d3x = fma(3.0, d2x, p.bB2 * y) * g.ir2;        // = -D^3 X
    const double ir4 = g.ir2 * g.ir2;
    const double x6 = p.bC * (ir4 * g.ir2);
    const double x8 = x6 * g.ir2, x10 = x8 * g.ir2, x12 = x10 * g.ir2;
    v = (X - x6) - p.b_shift;
    p1 = fma(6.0, x8, -y);
    p2 = fma(-48.0, x10, d2x);
    p3 = fma(480.0, x12, -d3x);
}

__device__ __forceinline__ void vdw_pair_fast(const PairHot& p, const PairGeom& g, const LJPowers& w, const double* etab,
                                              double& v, double& p1, double& p2, double& p3) {
    v = 0.0; p1 = 0.0; p2 = 0.0; p3 = 0.0;
    const int fl = p.flags;
    // hard sphere: the fast path only sees r2 >= fast_lo > every hs_r2, where the term is exactly 0
    if (fl & CEG_PAIR_LJ) lj_terms(p, w, v, p1, p2, p3);
    if (fl & CEG_PAIR_BUCK) {
        double bv, b1, b2, b3;
        buck_terms(p, g, etab, bv, b1, b2, b3);
        v += bv; p1 += b1; p2 += b2; p3 += b3;
    }
}

// ewald.jl:299-312 through the standard recurrence of the real-space Ewald radial functions.  With
// E = e^{-alpha^2 r^2}, K1 = 2 alpha/sqrt(pi), B0 = erfc(alpha r)/r and
//   B_{n} = ((2n-1) B_{n-1} + K1 (2 alpha^2)^{n-1} E) / r^2,
// the reference's (U, DU, D^2U, D^3U) are q (B0, -B1, B2, -B3): 11 FP64 operations after erfc instead
// of 19 for the expanded polynomials in alpha^2 r^2.  erfc(x) = E erfcx(x) shares the one exponential.
// POLY: alpha*cutoff lies inside the fitted range of erfcx_fast (the default Ewald precision), else
// the library erfcx.  kk = {K1, 2 alpha^2 K1, (2 alpha^2)^2 K1}.
template <bool POLY>
__device__ __forceinline__ void ewald_pair_fast(double alpha, double alpha2, const double* kk, double q,
                                                const PairGeom& g, const double* etab, double& v, double& p1,
                                                double& p2, double& p3) {
    const double E = exp_neg_fast(-(alpha2 * g.r2), etab);   // independent of the rsqrt chain
    const double ar = alpha * g.r;
    const double w = POLY ? erfcx_fast(ar) : erfcx(ar);
    const double qE = q * E;
    v = (qE * w) * g.ir;                                     // q B0
    const double b1 = fma(kk[0], qE, v) * g.ir2;             // q B1
    const double b2 = fma(3.0, b1, kk[1] * qE) * g.ir2;      // q B2
    const double b3 = fma(5.0, b2, kk[2] * qE) * g.ir2;      // q B3
    p1 = -b1;
    p2 = b2;
    p3 = -b3;
}

__device__ __forceinline__ void accumulate(double* acc, double v, double p1, double p2, double p3, double dx,
                                           double dy, double dz, double dxy, double dxz, double dyz) {
    acc[0] += v;
    acc[1] = fma(p1, dx, acc[1]);
    acc[2] = fma(p1, dy, acc[2]);
    acc[3] = fma(p1, dz, acc[3]);
    acc[4] = fma(p2, dxy, acc[4]);
    acc[5] = fma(p2, dxz, acc[5]);
    acc[6] = fma(p2, dyz, acc[6]);
    acc[7] = fma(p3, dxz * dy, acc[7]);
}

__device__ __forceinline__ double clamp_ref(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

// grids.jl:118-135 _set_gridpoint!
__device__ __forceinline__ void set_gridpoint(const double* acc, const double* delta, double lam, double thr,
                                              float* o) {
    double v = acc[0], a1 = acc[1], a2 = acc[2], a3 = acc[3], b1 = acc[4], b2 = acc[5], b3 = acc[6], c = acc[7];
    if (v > thr) {
        v = xmul(2.0, thr);
        a1 = clamp_ref(a1, -thr, thr); a2 = clamp_ref(a2, -thr, thr); a3 = clamp_ref(a3, -thr, thr);
        b1 = b2 = b3 = 0.0;
        c = 0.0;
    }
    o[0] = (float)xmul(v, lam);
    o[1] = (float)xmul(xmul(a1, delta[0]), lam);
    o[2] = (float)xmul(xmul(a2, delta[1]), lam);
    o[3] = (float)xmul(xmul(a3, delta[2]), lam);
    o[4] = (float)xmul(xmul(b1, xmul(delta[0], delta[1])), lam);
    o[5] = (float)xmul(xmul(b2, xmul(delta[0], delta[2])), lam);
    o[6] = (float)xmul(xmul(b3, xmul(delta[1], delta[2])), lam);
    o[7] = (float)xmul(xmul(c, xmul(xmul(delta[0], delta[1]), delta[2])), lam);
}

// ---- literal kernel (arbitrary positions) ---------------------------------------------------------
struct LiteralParams {
    CellDev cell;
    double cutoff2, alpha;
    int64_t n_atoms, n_pos;
    const double* xyz;      // n_atoms x 3
    const int* type;
    const double* q;
    const PairDev* table;   // n_probes x n_types
    int n_types, n_probes, want_coulomb;
    const double* pos;      // n_pos x 3
    double* out;            // [(g*n_pos + i)*8 + c]
};

__global__ void __launch_bounds__(128) derivatives_literal_kernel(const LiteralParams P) {
    __shared__ CellDev s_cell;
    if (threadIdx.x == 0) s_cell = P.cell;
    __syncthreads();
    const int n_grids = P.n_probes + P.want_coulomb;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P.n_pos; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = P.pos[3 * i], py = P.pos[3 * i + 1], pz = P.pos[3 * i + 2];
        for (int g = 0; g < n_grids; g++) {
            const bool coul = (g == P.n_probes);
            double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            double smallest = CUDART_INF;
            for (int64_t a = 0; a < P.n_atoms; a++) {
                double b[3] = {xsub(px, P.xyz[3 * a]), xsub(py, P.xyz[3 * a + 1]), xsub(pz, P.xyz[3 * a + 2])};
                const double d2 = periodic_distance2_ref(b, s_cell);
                if (d2 >= P.cutoff2) continue;
                double v, p1, p2, p3;
                if (coul) {
                    smallest = fmin(smallest, d2);
                    ewald_pair_ref(P.alpha, P.q[a], d2, v, p1, p2, p3);
                } else {
                    const PairDev pr = P.table[g * P.n_types + P.type[a]];
                    if (pr.flags == 0) continue;
                    vdw_pair_ref(pr, d2, v, p1, p2, p3);
                }
                accumulate(acc, v, p1, p2, p3, b[0], b[1], b[2], b[0] * b[1], b[0] * b[2], b[1] * b[2]);
            }
            if (coul && smallest < 1.0) acc[0] = CUDART_INF;
            double* o = P.out + ((int64_t)g * P.n_pos + i) * 8;
#pragma unroll
            for (int c = 0; c < 8; c++) o[c] = acc[c];
        }
    }
}

// ---- tiled production kernel ----------------------------------------------------------------------
struct ExtAtom {          // one periodic image of one supercell atom
    double x, y, z, q;
};
struct ExtMeta {
    int type;
    int orig;             // index of the supercell atom (for the literal re-evaluation)
};

struct BuildParams {
    CellDev cell;
    GridSpecDev gs;
    double cutoff2, alpha;
    double alpha2, kk[3];      // alpha^2; K1, 2 alpha^2 K1, (2 alpha^2)^2 K1 with K1 = 2 alpha / sqrt(pi): constant-bank operands
    double cut_hi;             // cutoff2 + GUARD_BAND: pairs below enter the pair code
    double fast_lo, fast_hi;   // d2 in [fast_lo, fast_hi] takes the fast path, anything else the literal one
    int hi_cut, hi_fast_lo;    // the same windows on the high 32 bits of d2 (see the kernel)
    unsigned hi_fast_span;
    double reach2;             // (cutoff + tile radius + slack)^2
    double lam_v, thr_v, lam_c, thr_c;
    int i0, i1;                // slab of x planes held by the output buffers
    int a0, a1;                // planes computed by this launch (i0 <= a0 < a1 <= i1)
    int tiles_x, tiles_y, tiles_z;
    int tile[3];               // nodes per tile along x, y, z (product = threads per CTA)
    int poly_erfcx;            // alpha*cutoff within the fitted range of erfcx_fast
    const ExtAtom* ext;
    const ExtMeta* ext_meta;
    int n_ext;
    const double* atoms_xyz;   // original supercell atoms
    const PairDev* table;
    int n_types;
    float* out[MAX_GRIDS];     // each [8][i1-i0][ny][nz]
    double* raw;               // optional [g][8][raw_slab/plane][ny][nz]; this launch's slab starts at raw_plane0
    int64_t raw_slab;
    int raw_plane0;
    unsigned long long* stats; // [0] pairs tested, [1] pairs in cutoff
};

// 128 registers per thread whatever the CTA size (512/NT CTAs, 16 warps resident per SM).  The 3 and 4 probe
// variants would like ~200 registers for their 32-40 accumulators, but at 8 resident warps per SM the FP64 pipe
// idles on dependency stalls: capped at 128 (a few hundred bytes of spills) they run 21 % faster (C3: 11.4 -> 9.1 ms).
// With fewer registers still (96 or 80, more CTAs) every variant gets slower.
#ifndef CEG_BUILD_MINB
#define CEG_BUILD_MINB(NT) (512 / (NT))
#endif
template <int NT, int NP, bool COUL, bool POLY>
__global__ void __launch_bounds__(NT, CEG_BUILD_MINB(NT))
grid_build_tiled_kernel(const BuildParams P) {
    constexpr int NW = NT / 32;
    const int TX = P.tile[0], TY = P.tile[1], TZ = P.tile[2];   // TX*TY*TZ == NT (4x8x8 main, flatter tail shapes)
    constexpr int NG = NP + (COUL ? 1 : 0);
    __shared__ ExtAtom s_atom[LIST_CAP];
    __shared__ ExtMeta s_meta[LIST_CAP];
    __shared__ PairHot s_table[NP > 0 ? NP * TABLE_SMEM_TYPES : 1];   // (probe x present species), <= 24 species
    __shared__ CellDev s_cell;
    __shared__ int s_wcount[NW];
    __shared__ double s_etab[EXP_TAB];   // 2^(j/256) for exp_neg_fast
    __shared__ unsigned char s_all_lj[TABLE_SMEM_TYPES];   // species whose rule is pure LJ for every probe

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_cell = P.cell;
    for (int i = tid; i < EXP_TAB; i += NT) s_etab[i] = c_exp_tab[i];
    if (NP > 0)
        for (int i = tid; i < NP * P.n_types; i += NT) s_table[i] = P.table[i];
    if (NP > 1 && tid < P.n_types) {
        bool all = true;
        for (int g = 0; g < NP; g++) all &= ((P.table[g * P.n_types + tid].flags & ~CEG_PAIR_HS) == CEG_PAIR_LJ);
        s_all_lj[tid] = all;
    }

    // tile -> node indices
    int t = blockIdx.x;
    const int tz = t % P.tiles_z; t /= P.tiles_z;
    const int ty = t % P.tiles_y; t /= P.tiles_y;
    const int tx = t;
    // thread -> node of the tile.  A warp evaluates a pair as soon as one of its lanes is inside the cutoff, so the more
    // compact a warp's 32 nodes, the fewer lanes idle at the cutoff sphere: 2 x 4 x 4 nodes per warp where the tile allows
    // it (CEG_COMPACT_WARPS), 1 x 4 x 8 (z fastest) otherwise.
    int lz = tid % TZ, ly = (tid / TZ) % TY, lx = tid / (TZ * TY);
#if CEG_COMPACT_WARPS
    if (TX >= 2 && (TY & 3) == 0 && (TZ & 3) == 0) {
        const int wz_n = TZ >> 2, wy_n = TY >> 2;
        const int wz = warp % wz_n, wy = (warp / wz_n) % wy_n, wx = warp / (wz_n * wy_n);
        lz = (wz << 2) + (lane & 3);
        ly = (wy << 2) + ((lane >> 2) & 3);
        lx = (wx << 1) + (lane >> 4);
    }
#endif
    const int i = P.a0 + tx * TX + lx, j = ty * TY + ly, k = tz * TZ + lz;
    const bool valid = (i < P.a1) && (j < P.gs.n[1]) && (k < P.gs.n[2]);
    // abc_to_xyz (coordinates.jl:72-76): multiply, divide, add — bit-identical to the oracle
    const double px = xadd(xdiv(xmul((double)i, P.gs.size[0]), (double)P.gs.dims[0]), P.gs.shift[0]);
    const double py = xadd(xdiv(xmul((double)j, P.gs.size[1]), (double)P.gs.dims[1]), P.gs.shift[1]);
    const double pz = xadd(xdiv(xmul((double)k, P.gs.size[2]), (double)P.gs.dims[2]), P.gs.shift[2]);
    // tile centre (nominal full tile)
    const double cx = ((double)(P.a0 + tx * TX) + 0.5 * (TX - 1)) * P.gs.delta[0] + P.gs.shift[0];
    const double cy = ((double)(ty * TY) + 0.5 * (TY - 1)) * P.gs.delta[1] + P.gs.shift[1];
    const double cz = ((double)(tz * TZ) + 0.5 * (TZ - 1)) * P.gs.delta[2] + P.gs.shift[2];

    double acc[NG > 0 ? NG : 1][8];
#pragma unroll
    for (int g = 0; g < NG; g++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[g][c] = 0.0;
    double smallest = CUDART_INF;
    unsigned int n_tested = 0, n_in = 0;
    int count = 0;
    __syncthreads();

    auto process = [&](int cnt) {
        if (!valid) return;
        n_tested += cnt;
        for (int e = 0; e < cnt; e++) {
            const ExtAtom a = s_atom[e];
            double dx = px - a.x, dy = py - a.y, dz = pz - a.z;
            double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
            // Window tests on the high word of d2 (>= 0, so it orders like an integer): they run on the
            // integer pipe and are conservative — anything not strictly inside the fast window, up to the
            // first high word above cutoff^2 + band, goes to the literal path, which decides exactly.
            const int d2hi = __double2hiint(d2);
            if (d2hi <= P.hi_cut) {
                const int ty_ = s_meta[e].type;
                if ((unsigned)(d2hi - P.hi_fast_lo) < P.hi_fast_span) {
                    // ---- hot path ----
                    n_in++;
                    const PairGeom geo = make_geom(d2);
                    const LJPowers ljw = make_lj_powers(geo);   // dead code in the blocks without LJ
                    const double dxy = dx * dy, dxz = dx * dz, dyz = dy * dz;
                    if (NP == 1 && COUL) {
                        // one straight-line block per interaction class, so that the independent exp / erfcx /
                        // algebra chains of the VdW and Ewald terms are scheduled against each other
                        const PairHot& pr = s_table[ty_];
                        const int fl = pr.flags & ~CEG_PAIR_HS;   // hard spheres contribute exactly 0 here
                        double v, p1, p2, p3, ev, e1, e2, e3;
                        if (fl == CEG_PAIR_BUCK) {
                            buck_terms(pr, geo, s_etab, v, p1, p2, p3);
                            ewald_pair_fast<POLY>(P.alpha, P.alpha2, P.kk, a.q, geo, s_etab, ev, e1, e2, e3);
                            accumulate(acc[0], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            accumulate(acc[1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        } else if (fl == CEG_PAIR_LJ) {
                            lj_terms(pr, ljw, v, p1, p2, p3);
                            ewald_pair_fast<POLY>(P.alpha, P.alpha2, P.kk, a.q, geo, s_etab, ev, e1, e2, e3);
                            accumulate(acc[0], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            accumulate(acc[1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        } else if (fl == 0) {
                            ewald_pair_fast<POLY>(P.alpha, P.alpha2, P.kk, a.q, geo, s_etab, ev, e1, e2, e3);
                            accumulate(acc[1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        } else {
                            vdw_pair_fast(pr, geo, ljw, s_etab, v, p1, p2, p3);
                            ewald_pair_fast<POLY>(P.alpha, P.alpha2, P.kk, a.q, geo, s_etab, ev, e1, e2, e3);
                            accumulate(acc[0], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            accumulate(acc[1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        }
                    } else if (NP > 1 && s_all_lj[ty_]) {
                        // every probe sees this species through a pure Lennard-Jones rule (the CH4 / CO2-site case):
                        // one straight-line block, so that the probes' short chains and the Ewald chains interleave
                        constexpr int NPA = NP > 0 ? NP : 1;
                        double v[NPA], p1[NPA], p2[NPA], p3[NPA];
#pragma unroll
                        for (int g = 0; g < NP; g++) lj_terms(s_table[g * P.n_types + ty_], ljw, v[g], p1[g], p2[g], p3[g]);
                        double ev = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
                        if (COUL) ewald_pair_fast<POLY>(P.alpha, P.alpha2, P.kk, a.q, geo, s_etab, ev, e1, e2, e3);
#pragma unroll
                        for (int g = 0; g < NP; g++) accumulate(acc[g], v[g], p1[g], p2[g], p3[g], dx, dy, dz, dxy, dxz, dyz);
                        if (COUL) accumulate(acc[NG - 1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                    } else {
#pragma unroll
                        for (int g = 0; g < NP; g++) {
                            const PairHot& pr = s_table[g * P.n_types + ty_];
                            if (pr.flags != 0) {
                                double v, p1, p2, p3;
                                vdw_pair_fast(pr, geo, ljw, s_etab, v, p1, p2, p3);
                                accumulate(acc[g], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            }
                        }
                        if (COUL) {   // d2 >= fast_lo > 1 Å^2 here: `smallest` (probes.jl:108) is tracked on the literal path
                            double v, p1, p2, p3;
                            ewald_pair_fast<POLY>(P.alpha, P.alpha2, P.kk, a.q, geo, s_etab, v, p1, p2, p3);
                            accumulate(acc[NG - 1], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                        }
                    }
                } else {
                    // ---- literal re-evaluation from the original atom (mask thresholds, r -> 0, image search)
                    const int o = s_meta[e].orig;
                    double b[3] = {xsub(px, P.atoms_xyz[3 * o]), xsub(py, P.atoms_xyz[3 * o + 1]),
                                   xsub(pz, P.atoms_xyz[3 * o + 2])};
                    const double d2x = periodic_distance2_ref(b, s_cell);
                    // this list entry must be the image the reference measured d2 on
                    if (fabs(d2x - d2) > 1e-6) continue;
                    if (d2x >= P.cutoff2) continue;
                    n_in++;
                    const double bxy = b[0] * b[1], bxz = b[0] * b[2], byz = b[1] * b[2];
#pragma unroll
                    for (int g = 0; g < NP; g++) {
                        const PairDev& pr = P.table[g * P.n_types + ty_];
                        if (pr.flags != 0) {
                            double v, p1, p2, p3;
                            vdw_pair_ref(pr, d2x, v, p1, p2, p3);
                            accumulate(acc[g], v, p1, p2, p3, b[0], b[1], b[2], bxy, bxz, byz);
                        }
                    }
                    if (COUL) {
                        smallest = fmin(smallest, d2x);
                        double v, p1, p2, p3;
                        ewald_pair_ref(P.alpha, a.q, d2x, v, p1, p2, p3);
                        accumulate(acc[NG - 1], v, p1, p2, p3, b[0], b[1], b[2], bxy, bxz, byz);
                    }
                }
            }
        }
    };

    for (int base = 0; base < P.n_ext; base += NT) {
        const int e = base + tid;
        bool keep = false;
        ExtAtom a;
        ExtMeta m;
        if (e < P.n_ext) {
            a = P.ext[e];
            const double ddx = cx - a.x, ddy = cy - a.y, ddz = cz - a.z;
            keep = (ddx * ddx + ddy * ddy + ddz * ddz) < P.reach2;
            if (keep) m = P.ext_meta[e];
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        int off = count, total = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const int cw = s_wcount[w];
            if (w < warp) off += cw;
            total += cw;
        }
        if (keep) {
            const int slot = off + __popc(bal & ((1u << lane) - 1u));
            s_atom[slot] = a;
            s_meta[slot] = m;
        }
        count += total;
        __syncthreads();
        if (count + NT > LIST_CAP) {   // flush: the next chunk might not fit
            process(count);
            count = 0;
            __syncthreads();
        }
    }
    process(count);

    // stats: one atomic per warp
    if (P.stats) {
        unsigned long long t64 = n_tested, i64 = n_in;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t64 += __shfl_down_sync(0xffffffffu, t64, o);
            i64 += __shfl_down_sync(0xffffffffu, i64, o);
        }
        if (lane == 0) {
            atomicAdd(&P.stats[0], t64);
            atomicAdd(&P.stats[1], i64);
        }
    }
    if (!valid) return;
    if (COUL && smallest < 1.0) acc[NG - 1][0] = CUDART_INF;   // probes.jl:116

    const int64_t ny = P.gs.n[1], nz = P.gs.n[2];
    const int64_t plane = ny * nz;
    const int64_t slab = (int64_t)(P.i1 - P.i0) * plane;
    const int64_t node = (int64_t)(i - P.i0) * plane + (int64_t)j * nz + k;
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const bool coul = COUL && (g == NG - 1);
        if (P.raw) {
#pragma unroll
            for (int c = 0; c < 8; c++)
                P.raw[((int64_t)g * 8 + c) * P.raw_slab + (int64_t)P.raw_plane0 * plane + node] = acc[g][c];
        }
        if (P.out[g]) {
            float o[8];
            set_gridpoint(acc[g], P.gs.delta, coul ? P.lam_c : P.lam_v, coul ? P.thr_c : P.thr_v, o);
#pragma unroll
            for (int c = 0; c < 8; c++) P.out[g][(int64_t)c * slab + node] = o[c];
        }
    }
}

}  // namespace ceg
