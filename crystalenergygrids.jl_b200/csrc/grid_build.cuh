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
//    nodes; it scans the part of the pre-expanded list of periodic atom IMAGES (sorted along x on the
//    host) that can reach its x planes, keeps the images within cutoff + tile radius of the tile
//    centre with an ordered compaction into shared memory, and every thread (one node each) walks
//    that list, broadcast from shared memory.  All VdW probes and the Coulomb grid share the one
//    distance computation.  For the Coulomb grid the Gaussian e^{-alpha^2 r^2} of the Ewald term is
//    not evaluated per pair: it factorises over the axes, so the CTA tabulates e^{-alpha^2 (x_i - a_x)^2}
//    for the TX + TY + TZ node coordinates of its tile per image (20 exponentials per image instead of
//    256) and a pair takes the product of three table entries.  The walk is software-pipelined: the
//    distance, cutoff word and reciprocal-square-root seed of entry e+1 are computed inside the
//    arithmetic block of entry e.  Because the probe supercell has perpendicular widths >= 2*cutoff
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
    double b_lnA;      // ln A when CEG_PAIRHOT_LNA is set in flags
};
constexpr int CEG_PAIRHOT_LNA = 256;   // internal flag: A > 0, the hot path folds A into the exponent
// What the hot path reads of a PairDev (the shared-memory copy; the literal path reads the global table)
struct alignas(16) PairHot {     // 80 bytes, 16-byte aligned pairs for ld.shared.v2.f64
    int flags, _pad;
    double b_lnA;
    double bB, bB2;
    double bC, b_shift;
    double lj_c6, lj_c12;
    double lj_shift, bA;
};
static_assert(sizeof(PairHot) == 80, "PairHot layout");
__device__ __forceinline__ void to_hot(PairHot& h, const PairDev& d) {
    h.flags = d.flags; h._pad = 0;
    h.lj_c6 = d.lj_c6; h.lj_c12 = d.lj_c12; h.lj_shift = d.lj_shift;
    h.bA = d.bA; h.bB = d.bB; h.bC = d.bC; h.b_shift = d.b_shift; h.bB2 = d.bB2; h.b_lnA = d.b_lnA;
}

// explicit shared-memory loads at 32-bit shared-window addresses
__device__ __forceinline__ void lds_v2(uint32_t addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double a;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
    return a;
}
__device__ __forceinline__ void lds_v2i(uint32_t addr, int& a, uint32_t& b) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}
__device__ __forceinline__ int lds_i32(uint32_t addr) {
    int a;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a) : "r"(addr));
    return a;
}
struct HotLJ { double c6, c12, shift; };
struct HotBuck { double A, lnA, B, B2, C, shift; };
__device__ __forceinline__ HotLJ load_lj(uint32_t pr) {
    HotLJ h;
    lds_v2(pr + 48, h.c6, h.c12);
    h.shift = lds_f64(pr + 64);
    return h;
}
__device__ __forceinline__ HotBuck load_buck(uint32_t pr) {   // folded form: A is not read
    HotBuck h;
    h.A = 0.0;
    h.lnA = lds_f64(pr + 8);
    lds_v2(pr + 16, h.B, h.B2);
    lds_v2(pr + 32, h.C, h.shift);
    return h;
}
__device__ __forceinline__ HotBuck load_buck_general(uint32_t pr) {
    HotBuck h = load_buck(pr);
    h.A = lds_f64(pr + 72);
    return h;
}

#ifndef CEG_COMPACT_WARPS
#define CEG_COMPACT_WARPS 1
#endif
constexpr int MAX_GRIDS = 5;       // up to 4 fused VdW probes + Coulomb per pass
constexpr int MAX_BUILD_DEST = 8;  // replicas one build launch can write its interpolation records to (one per GPU of a node)
constexpr int LIST_CAP_MAX = 1024;   // shared-memory image list capacity per round (the host picks the actual one)
constexpr size_t TABLE_SMEM_MAX = 64 << 10;   // bytes of shared memory the (probe x species) pair table may take

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
// fastmath.cuh; differences from the literal formulas are ~1e-12 relative.
struct PairGeom {
    double ir, ir2, r;
};
__device__ __forceinline__ PairGeom make_geom(double ir, double r) {
    PairGeom g;
    g.ir = ir;
    g.ir2 = ir * ir;
    g.r = r;
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
__device__ __forceinline__ void lj_terms(const HotLJ& p, const LJPowers& w, double& v, double& p1, double& p2,
                                         double& p3) {
    const double a = p.c6 * w.u3;
    const double b = p.c12 * w.u6;
    v = (b - a) - p.shift;
    p1 = w.k1 * fma(b, -2.0, a);
    p2 = w.k2 * fma(b, 3.5, -a);
    p3 = w.k3 * fma(b, -5.6, a);
}

// Buckingham terms (interactions.jl:447-457), no branches.  With X = A e^{-Br}, y = B X / r:
//   D X = -y,  D^2 X = (y + B^2 X)/r^2,  D^3 X = -(3 D^2 X + B^2 y)/r^2   (D = (1/r) d/dr),
// and the dispersion part -C/r^6 contributes 6, -48, 480 times C/r^8, C/r^10, C/r^12.
// LNA: A > 0 is folded into the exponent, X = e^{ln A - B r} (one FMA instead of a product and a multiplication).
template <bool LNA, bool SPLIT = true>
__device__ __forceinline__ void buck_terms(const HotBuck& p, const PairGeom& g, double& v, double& p1, double& p2,
                                           double& p3) {
    const double B = p.B;
    const double X = LNA ? exp_fast<SPLIT>(fma(-B, g.r, p.lnA)) : p.A * exp_fast<SPLIT>(-(B * g.r));
    const double BX = B * X;
    const double y = BX * g.ir;
    const double d2x = fma(B, BX, y) * g.ir2;
    const double d3x = fma(3.0, d2x, p.B2 * y) * g.ir2;         // = -D^3 X
    const double ir4 = g.ir2 * g.ir2;
    const double x6 = p.C * (ir4 * g.ir2);
    const double x8 = x6 * g.ir2, x10 = x8 * g.ir2, x12 = x10 * g.ir2;
    v = (X - x6) - p.shift;
    p1 = fma(6.0, x8, -y);
    p2 = fma(-48.0, x10, d2x);
    p3 = fma(480.0, x12, -d3x);
}

// general entry (any combination of terms), read from the shared-memory table at address `pr`
template <bool SPLIT = true>
__device__ __forceinline__ void vdw_pair_fast(uint32_t pr, const PairGeom& g, const LJPowers& w, double& v, double& p1,
                                              double& p2, double& p3) {
    v = 0.0; p1 = 0.0; p2 = 0.0; p3 = 0.0;
    const int fl = lds_i32(pr);
    // hard sphere: the fast path only sees r2 >= fast_lo > every hs_r2, where the term is exactly 0
    if (fl & CEG_PAIR_LJ) lj_terms(load_lj(pr), w, v, p1, p2, p3);
    if (fl & CEG_PAIR_BUCK) {
        double bv, b1, b2, b3;
        buck_terms<false, SPLIT>(load_buck_general(pr), g, bv, b1, b2, b3);
        v += bv; p1 += b1; p2 += b2; p3 += b3;
    }
}

// ewald.jl:299-312 through the standard recurrence of the real-space Ewald radial functions.  With
// E = e^{-alpha^2 r^2}, K1 = 2 alpha/sqrt(pi), B0 = erfc(alpha r)/r and
//   B_{n} = ((2n-1) B_{n-1} + K1 (2 alpha^2)^{n-1} E) / r^2,
// the reference's (U, DU, D^2U, D^3U) are q (B0, -B1, B2, -B3): 11 FP64 operations after erfc instead
// of 19 for the expanded polynomials in alpha^2 r^2.  erfc(x) = E erfcx(x); E arrives as the product of the
// tile's three per-axis Gaussian factors (see the kernel).
// POLY: [alpha * r_min, alpha * cutoff] lies inside the fitted range of erfcx_fast (the default Ewald
// precision), else the library erfcx.  kk = {K1, 2 alpha^2 K1, (2 alpha^2)^2 K1}.
// z: erfcx_arg(r) of the pair, computed one list entry ahead (POLY only).
template <bool POLY, bool SPLIT = true>
__device__ __forceinline__ void ewald_pair_fast(double alpha, const double* kk, double q, const PairGeom& g, double E, double z,
                                                double& v, double& p1, double& p2, double& p3) {
    const double w = POLY ? erfcx_poly<SPLIT>(z) : erfcx(alpha * g.r);
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
    int type;             // dense species index
    int orig;             // index of the supercell atom (for the literal re-evaluation)
};
struct ListMeta {         // what the hot loop reads per list entry (shared memory)
    int type;             // species index | interaction class << 16
    uint32_t pr;          // shared-window address of the species' pair-table entry (probe 0)
};

struct BuildParams {
    CellDev cell;
    GridSpecDev gs;
    double cutoff2, alpha;
    double alpha2, kk[3];      // alpha^2; K1, 2 alpha^2 K1, (2 alpha^2)^2 K1 with K1 = 2 alpha / sqrt(pi): constant-bank operands
    double alpha_p;            // alpha * ERFCX_P (argument map of erfcx_fast)
    double gauss_step[3];      // exp(-2 alpha^2 delta_d^2) per axis (Gaussian factor recurrence)
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
    int poly_erfcx;            // [alpha r_min, alpha cutoff] within the fitted range of erfcx_fast
    int list_cap;              // entries of the shared-memory image list (dynamic shared memory, see build_smem_bytes)
    const ExtAtom* ext;        // images sorted by x
    const ExtMeta* ext_meta;
    int n_ext;
    const int* ext_lo;         // per x plane p (index p - ext_plane0): first image with x >= x_p - cutoff ...
    const int* ext_hi;         // ... and one past the last image with x <= x_p + cutoff
    int ext_plane0;
    const double* atoms_xyz;   // original supercell atoms
    const PairDev* table;
    int n_types;
    float* out[MAX_GRIDS];     // each [8][i1-i0][ny][nz]
    // Interpolation records written by the kernel itself: node (i, j, k) of grid g goes, converted like parse_grid would
    // (rec_scale), to its place in the record layout of n_dest destinations — this device's grid handle and, over NVLink
    // peer mappings, the replicas on the other GPUs of a sharded build (the all-gather of the slabs and the re-layout
    // pass fused into the epilogue; the reference hands the grids over through `.grid` files, raspa.jl:430-438).
    int n_dest;                // 0: no record output
    int rec_block;             // destination layout: 2x2x2 bricks (1) or node-major (0)
    Rec8* rec[MAX_BUILD_DEST][MAX_GRIDS];
    double rec_scale[MAX_GRIDS];
    double* raw;               // optional [g][8][raw_slab/plane][ny][nz]; this launch's slab starts at raw_plane0
    int64_t raw_slab;
    int raw_plane0;
    unsigned long long* stats; // [0] pairs tested, [1] pairs in cutoff
};

// dynamic shared memory of one CTA: image list (atoms, meta), for the Coulomb grid the per-axis Gaussian factors, then
// the (probe x species) pair table and the species classes
inline size_t build_table_bytes(int n_probes, int n_types) {
    return (size_t)n_probes * n_types * sizeof(PairHot) + (((size_t)n_types + 15) & ~(size_t)15);
}
inline size_t build_smem_bytes(int cap, int n_axes, bool coul, int n_probes, int n_types) {
    return (size_t)(cap + 1) * (sizeof(ExtAtom) + sizeof(ListMeta)) + (size_t)cap * sizeof(int) +
           (coul ? (size_t)cap * sizeof(double) * n_axes : 0) + build_table_bytes(n_probes, n_types);
}

// interaction class of a species as seen by the pass (kept in bits 16.. of the list's type word)
enum { CLS_NONE = 0, CLS_LJ = 1, CLS_BUCK = 2, CLS_GENERAL = 3 };

// 128 registers per thread whatever the CTA size (512/NT CTAs, 16 warps resident per SM).  The 3 and 4 probe
// variants would like ~200 registers for their 32-40 accumulators, but at 8 resident warps per SM the FP64 pipe
// idles on dependency stalls: capped at 128 (a few hundred bytes of spills) they run 21 % faster (C3: 11.4 -> 9.1 ms).
// With fewer registers still (96 or 80, more CTAs) every variant gets slower.
#ifndef CEG_BUILD_MINB
#define CEG_BUILD_MINB(NT) (512 / (NT))
#endif
// Software pipelining of the list walk: 0 = none (the state of an entry is computed when it is reached), 1 = the loads,
// distance and rsqrt seed of entry e+1 are issued inside the arithmetic block of entry e, 2 = also its 1/r, r and erfcx
// argument map.  Measured on B200 (C4, round 2): 61.2 / see DESIGN.md / 68.0 ms — with the short dependency chains of
// this version the carried state costs more (registers, spills) than the hidden latency is worth.
// Measured and rejected in round 2 (C4, B200): walking the list two entries at a time with both entries' arithmetic in one
// straight-line block (warp-uniform votes, masked contributions) — two independent chains for the in-order scheduler —
// 65.3 ms against 61.3 ms: the selects and the wasted lanes cost more than the second chain hides.
#ifndef CEG_PIPE_DEPTH
#define CEG_PIPE_DEPTH 0
#endif
// Passes with at most this many fused VdW probes take the Ewald Gaussian from the per-tile factor table; beyond, one
// exponential per pair (C3, 3 LJ probes + Coulomb: 9.0 ms with the table — its rows halve the list capacity and its three
// loads cost registers those passes do not have — 6.75 ms without; round 1: 6.9 ms).
#ifndef CEG_GTABLE_MAXNP
#define CEG_GTABLE_MAXNP 2
#endif
__host__ __device__ constexpr bool build_uses_gtable(int n_probes, bool coul) { return coul && n_probes <= CEG_GTABLE_MAXNP; }
#ifndef CEG_SPLIT_MAXNP
#define CEG_SPLIT_MAXNP 4
#endif
#ifndef CEG_PIPELINE_MAXNP
#define CEG_PIPELINE_MAXNP 1   // passes with more fused VdW probes never pipeline (registers are too short)
#endif
template <int NT, int NP, bool COUL, bool POLY>
__global__ void __launch_bounds__(NT, CEG_BUILD_MINB(NT))
grid_build_tiled_kernel(const BuildParams P) {
    constexpr int NW = NT / 32;
    const int TX = P.tile[0], TY = P.tile[1], TZ = P.tile[2];   // TX*TY*TZ == NT (4x8x8 main, flatter tail shapes)
    const int NAX = TX + TY + TZ;
    constexpr int NG = NP + (COUL ? 1 : 0);
    constexpr int NPA = NP > 0 ? NP : 1;
    constexpr int PIPE = NP <= CEG_PIPELINE_MAXNP ? CEG_PIPE_DEPTH : 0;
    constexpr bool SPLIT = NP <= CEG_SPLIT_MAXNP;   // even/odd Horner (two chains per polynomial)
    constexpr bool GT = build_uses_gtable(NP, COUL);
    extern __shared__ __align__(16) unsigned char s_build[];
    const int cap = P.list_cap;
    // layout: pair table | atoms (cap + 1) | Gaussian factors (cap x NAX) | meta (cap + 1) | atom indices (cap) | species classes
    PairHot* const s_table = reinterpret_cast<PairHot*>(s_build);                               // (probe x present species)
    ExtAtom* const s_atom = reinterpret_cast<ExtAtom*>(s_table + NP * P.n_types);
    double* const s_g = reinterpret_cast<double*>(s_atom + (cap + 1));
    ListMeta* const s_meta = reinterpret_cast<ListMeta*>(s_g + (GT ? (size_t)cap * NAX : 0));
    int* const s_orig = reinterpret_cast<int*>(s_meta + (cap + 1));
    unsigned char* const s_cls = reinterpret_cast<unsigned char*>(s_orig + cap);                // interaction class per species
    __shared__ CellDev s_cell;
    __shared__ int s_wcount[NW];
    __shared__ double s_chain_node[8];                      // Gaussian-table chains: anchor coordinate, axis, row offset, range
    __shared__ int s_chain_axis[8], s_chain_off[8], s_chain_lo[8], s_chain_hi[8], s_nchain;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_cell = P.cell;
    if (NP > 0)
        for (int i = tid; i < NP * P.n_types; i += NT) to_hot(s_table[i], P.table[i]);
    for (int sp = tid; sp < P.n_types; sp += NT) {
        int cls = CLS_NONE;
        if (NP == 1) {
            const int fl = P.table[sp].flags;
            const int k = fl & ~(CEG_PAIR_HS | CEG_PAIRHOT_LNA);   // hard spheres contribute exactly 0 on the fast path
            cls = k == 0 ? CLS_NONE : k == CEG_PAIR_LJ ? CLS_LJ : (k == CEG_PAIR_BUCK && (fl & CEG_PAIRHOT_LNA)) ? CLS_BUCK : CLS_GENERAL;
        } else if (NP > 1) {
            bool all = true;   // every probe sees this species through a pure Lennard-Jones rule
            for (int g = 0; g < NP; g++)
                all &= ((P.table[g * P.n_types + sp].flags & ~(CEG_PAIR_HS | CEG_PAIRHOT_LNA)) == CEG_PAIR_LJ);
            cls = all ? CLS_LJ : CLS_GENERAL;
        }
        s_cls[sp] = (unsigned char)cls;
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
    if (GT && tid == 0) {   // chains of the Gaussian factor table: per axis, one per block of 8 global node indices
        int nc = 0;
        for (int ax = 0; ax < 3; ax++) {
            const int s0 = ax == 0 ? P.a0 + tx * TX : (ax == 1 ? ty * TY : tz * TZ);   // first node of the tile
            const int n = ax == 0 ? TX : (ax == 1 ? TY : TZ);
            const int off = ax == 0 ? 0 : (ax == 1 ? TX : TX + TY);
            for (int b = s0 & ~7; b < s0 + n && nc < 8; b += 8, nc++) {
                s_chain_axis[nc] = ax;
                s_chain_node[nc] = xadd(xdiv(xmul((double)b, P.gs.size[ax]), (double)P.gs.dims[ax]), P.gs.shift[ax]);
                s_chain_off[nc] = off + (b - s0);
                s_chain_lo[nc] = max(0, s0 - b);
                s_chain_hi[nc] = min(8, s0 + n - b);
            }
        }
        s_nchain = nc;
    }
    // tile centre (nominal full tile)
    const double cx = ((double)(P.a0 + tx * TX) + 0.5 * (TX - 1)) * P.gs.delta[0] + P.gs.shift[0];
    const double cy = ((double)(ty * TY) + 0.5 * (TY - 1)) * P.gs.delta[1] + P.gs.shift[1];
    const double cz = ((double)(tz * TZ) + 0.5 * (TZ - 1)) * P.gs.delta[2] + P.gs.shift[2];
    // offsets of this node's three Gaussian factors inside an image's row of the table
    const int gox = lx, goy = TX + ly, goz = TX + TY + lz;

    double acc[NG > 0 ? NG : 1][8];
#pragma unroll
    for (int g = 0; g < NG; g++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[g][c] = 0.0;
    bool close_hit = false;          // an atom closer than 1 Å (probes.jl:108-116; only the literal path sees such pairs)
    unsigned int n_in = 0;
    int n_listed = 0;                // list entries walked by this CTA (uniform)
    int count = 0;
    __syncthreads();

    // 32-bit shared-window addresses of the list (the hot loop reads it with explicit ld.shared at running addresses:
    // no per-iteration index arithmetic, no generic-to-shared conversion)
    const uint32_t sa_table = (uint32_t)__cvta_generic_to_shared(s_build);
    const uint32_t sa_atom = sa_table + (uint32_t)(NP * P.n_types) * (uint32_t)sizeof(PairHot);
    const uint32_t sa_g = sa_atom + (uint32_t)(cap + 1) * (uint32_t)sizeof(ExtAtom);
    const uint32_t sa_meta = sa_g + (GT ? (uint32_t)cap * (uint32_t)NAX * 8u : 0u);
    const uint32_t g_stride = (uint32_t)NAX * 8u;
    const uint32_t row_stride = (uint32_t)P.n_types * (uint32_t)sizeof(PairHot);   // probe g's row of the pair table

    auto process = [&](int cnt) {
        if (GT) {
            // Gaussian factor table g = exp(-alpha^2 (x_node - a)^2) for the tile's node coordinates along each axis, by the
            // recurrence g_{k+1} = g_k rho_k, rho_{k+1} = rho_k c with rho_0 = exp(-alpha^2 D (2 t0 + D)), c = exp(-2 alpha^2 D^2)
            // (two exponentials per chain instead of one per node).  Chains are anchored at the nodes whose GLOBAL index is a
            // multiple of 8, whatever the tile origin or shape, so a node's factor — and with it every grid value — does not
            // depend on how the grid is split into slabs and tiles.  One thread per (image, chain).
            const double na2 = -P.alpha2;
            const int nchain = s_nchain;
            for (int idx = tid; idx < cnt * nchain; idx += NT) {      // every thread busy: (image, chain) pairs in dense order
                const int e = idx / nchain, j = idx - e * nchain;
                const int ax = s_chain_axis[j], lo = s_chain_lo[j], hi = s_chain_hi[j];
                const double node0 = s_chain_node[j], dlt = P.gs.delta[ax], cc = P.gauss_step[ax];
                const double t0 = node0 - reinterpret_cast<const double*>(s_atom + e)[ax];
                double g = exp_acc((na2 * t0) * t0);
                double rho = exp_acc((na2 * dlt) * fma(2.0, t0, dlt));
                double* dst = s_g + e * NAX + s_chain_off[j];     // position of the anchor in the image's row (may precede it)
                for (int k = 0; k < hi; k++) {
                    if (k >= lo) dst[k] = g;
                    g *= rho;
                    rho *= cc;
                }
            }
            __syncthreads();
        }
        n_listed += cnt;
        if (!valid || cnt == 0) return;
        // State of the entry being evaluated: displacement, high word of d^2, 1/r, r, the erfcx argument map z and the
        // entry's charge / class / pair-table address.  That of the next entry is produced inside the arithmetic of the
        // current one (CEG_NEXT): its loads, distance, reciprocal square root and the reciprocal of the erfcx map then sit
        // off the critical path (ncu, round 2: the walk is bound by dependent-issue latency, not by the FP64 pipe).
        // Entry `cnt` (one past the end, allocated) is read but never used.
        double dx, dy, dz, ir, r, z = 0.0, q;
        int d2hi, mt;
        uint32_t pr;
        // Running addresses (entry to load next; this node's Gaussian factors of entry e) where registers allow; the
        // passes with several fused probes recompute them from e instead (5 loop-carried registers less: ptxas had
        // put them into local memory, 17 LDL + 6 STL per list entry).
        constexpr bool RUNPTR = NP <= 1;
        uint32_t pa = sa_atom, pm = sa_meta;
        uint32_t pgx = sa_g + gox * 8u, pgy = sa_g + goy * 8u, pgz = sa_g + goz * 8u;

        // stage A: loads, displacement, d^2 and the rsqrt seed; stage B: 1/r, r, erfcx argument map
        auto stage_a = [&](int ei, double& dx_, double& dy_, double& dz_, double& d2_, int& d2hi_, double& seed_, double& q_,
                           int& mt_, uint32_t& pr_) {
            double bx_, by_, bz_;
            if (RUNPTR) {
                lds_v2(pa, bx_, by_);
                lds_v2(pa + 16, bz_, q_);
                lds_v2i(pm, mt_, pr_);
                pa += (uint32_t)sizeof(ExtAtom);
                pm += (uint32_t)sizeof(ListMeta);
            } else {
                const uint32_t pa_ = sa_atom + (uint32_t)ei * (uint32_t)sizeof(ExtAtom);
                lds_v2(pa_, bx_, by_);
                lds_v2(pa_ + 16, bz_, q_);
                lds_v2i(sa_meta + (uint32_t)ei * (uint32_t)sizeof(ListMeta), mt_, pr_);
            }
            dx_ = px - bx_; dy_ = py - by_; dz_ = pz - bz_;
            d2_ = fma(dx_, dx_, fma(dy_, dy_, dz_ * dz_));
            d2hi_ = __double2hiint(d2_);
            seed_ = rsqrt_seed(d2_);
        };
        auto stage_b = [&](double d2_, double seed_, double& ir_, double& r_, double& z_) {
            ir_ = rsqrt_refine(d2_, seed_);
            r_ = d2_ * ir_;
            if (COUL && POLY) z_ = erfcx_arg(r_, P.alpha_p);
        };
        double d2 = 0.0, seed = 0.0;
        if (PIPE >= 1) stage_a(0, dx, dy, dz, d2, d2hi, seed, q, mt, pr);
        if (PIPE == 2) stage_b(d2, seed, ir, r, z);
        for (int e = 0; e < cnt; e++) {
            double dxn, dyn, dzn, d2n, seedn, irn, rn, zn = 0.0, qn;
            int d2hin, mtn;
            uint32_t prn;
            if (PIPE == 0) stage_a(e, dx, dy, dz, d2, d2hi, seed, q, mt, pr);
#define CEG_NEXT()                                                         \
    if (PIPE >= 1) {                                                       \
        stage_a(e + 1, dxn, dyn, dzn, d2n, d2hin, seedn, qn, mtn, prn);    \
        if (PIPE == 2) stage_b(d2n, seedn, irn, rn, zn);                   \
    }
            // Window tests on the high word of d2 (>= 0, so it orders like an integer): they run on the
            // integer pipe and are conservative — anything not strictly inside the fast window, up to the
            // first high word above cutoff^2 + band, goes to the literal path, which decides exactly.
            if ((unsigned)(d2hi - P.hi_fast_lo) < P.hi_fast_span) {
                {
                    // ---- hot path ----
                    n_in++;
                    const int cls = mt >> 16;
                    if (PIPE < 2) stage_b(d2, seed, ir, r, z);
                    const PairGeom geo = make_geom(ir, r);
                    double E = 1.0;
                    if (COUL && !GT) {
                        if (PIPE == 2) d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                        E = exp_poly<EXPA_DEG, SPLIT>(-(P.alpha2 * d2), c_expa);   // partly cancelling Coulomb sums: the accurate degree
                    } else if (COUL) {
                        if (RUNPTR) {
                            E = (lds_f64(pgx) * lds_f64(pgy)) * lds_f64(pgz);
                        } else {
                            const uint32_t pg = sa_g + (uint32_t)e * g_stride;
                            E = (lds_f64(pg + gox * 8u) * lds_f64(pg + goy * 8u)) * lds_f64(pg + goz * 8u);
                        }
                    }
                    const double dxy = dx * dy, dxz = dx * dz, dyz = dy * dz;
                    if (NP == 1 && COUL) {
                        // one straight-line block per interaction class, so that the independent exp / erfcx /
                        // algebra chains of the VdW and Ewald terms are scheduled against each other
                        double v, p1, p2, p3, ev, e1, e2, e3;
                        if (cls == CLS_BUCK) {
                            buck_terms<true, SPLIT>(load_buck(pr), geo, v, p1, p2, p3);
                            CEG_NEXT();
                            ewald_pair_fast<POLY, SPLIT>(P.alpha, P.kk, q, geo, E, z, ev, e1, e2, e3);
                            accumulate(acc[0], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            accumulate(acc[NG - 1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        } else if (cls == CLS_NONE) {
                            ewald_pair_fast<POLY, SPLIT>(P.alpha, P.kk, q, geo, E, z, ev, e1, e2, e3);
                            CEG_NEXT();
                            accumulate(acc[NG - 1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        } else if (cls == CLS_LJ) {
                            const LJPowers ljw = make_lj_powers(geo);
                            lj_terms(load_lj(pr), ljw, v, p1, p2, p3);
                            CEG_NEXT();
                            ewald_pair_fast<POLY, SPLIT>(P.alpha, P.kk, q, geo, E, z, ev, e1, e2, e3);
                            accumulate(acc[0], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            accumulate(acc[NG - 1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        } else {
                            const LJPowers ljw = make_lj_powers(geo);
                            vdw_pair_fast<SPLIT>(pr, geo, ljw, v, p1, p2, p3);
                            CEG_NEXT();
                            ewald_pair_fast<POLY, SPLIT>(P.alpha, P.kk, q, geo, E, z, ev, e1, e2, e3);
                            accumulate(acc[0], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            accumulate(acc[NG - 1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        }
                    } else if (NP > 1 && cls == CLS_LJ) {
                        // every probe sees this species through a pure Lennard-Jones rule (the CH4 / CO2-site case):
                        // one straight-line block, so that the probes' short chains and the Ewald chains interleave
                        const LJPowers ljw = make_lj_powers(geo);
                        if (NP <= 2) {
                            double v[NPA], p1[NPA], p2[NPA], p3[NPA];
#pragma unroll
                            for (int g = 0; g < NP; g++)
                                lj_terms(load_lj(pr + (uint32_t)g * row_stride), ljw, v[g], p1[g], p2[g], p3[g]);
                            CEG_NEXT();
                            double ev = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
                            if (COUL) ewald_pair_fast<POLY, SPLIT>(P.alpha, P.kk, q, geo, E, z, ev, e1, e2, e3);
#pragma unroll
                            for (int g = 0; g < NP; g++) accumulate(acc[g], v[g], p1[g], p2[g], p3[g], dx, dy, dz, dxy, dxz, dyz);
                            if (COUL) accumulate(acc[NG - 1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                        } else {
                            // three or four probes: 24-40 accumulators leave no registers to hold every probe's terms at once
                            if (COUL) {
                                double ev, e1, e2, e3;
                                ewald_pair_fast<POLY, SPLIT>(P.alpha, P.kk, q, geo, E, z, ev, e1, e2, e3);
                                accumulate(acc[NG - 1], ev, e1, e2, e3, dx, dy, dz, dxy, dxz, dyz);
                            }
                            CEG_NEXT();
#pragma unroll
                            for (int g = 0; g < NP; g++) {
                                double v, p1, p2, p3;
                                lj_terms(load_lj(pr + (uint32_t)g * row_stride), ljw, v, p1, p2, p3);
                                accumulate(acc[g], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            }
                        }
                    } else {
                        const LJPowers ljw = make_lj_powers(geo);   // dead code in the blocks without LJ
                        if (COUL) {   // d2 >= fast_lo > 1 Å^2 here: `smallest` (probes.jl:108) is tracked on the literal path
                            double v, p1, p2, p3;
                            ewald_pair_fast<POLY, SPLIT>(P.alpha, P.kk, q, geo, E, z, v, p1, p2, p3);
                            accumulate(acc[NG - 1], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                        }
                        CEG_NEXT();
#pragma unroll
                        for (int g = 0; g < NP; g++) {
                            const uint32_t prg = pr + (uint32_t)g * row_stride;
                            if (lds_i32(prg) != 0) {
                                double v, p1, p2, p3;
                                vdw_pair_fast<SPLIT>(prg, geo, ljw, v, p1, p2, p3);
                                accumulate(acc[g], v, p1, p2, p3, dx, dy, dz, dxy, dxz, dyz);
                            }
                        }
                    }
                }
            } else if (d2hi <= P.hi_cut) {
                {
                    CEG_NEXT();
                    // ---- literal re-evaluation from the original atom (mask thresholds, r -> 0, image search)
                    const int ty_ = mt & 0xffff;
                    if (PIPE == 2) d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                    const int o = s_orig[e];
                    double b[3] = {xsub(px, P.atoms_xyz[3 * o]), xsub(py, P.atoms_xyz[3 * o + 1]),
                                   xsub(pz, P.atoms_xyz[3 * o + 2])};
                    const double d2x = periodic_distance2_ref(b, s_cell);
                    // this list entry must be the image the reference measured d2 on
                    if (fabs(d2x - d2) <= 1e-6 && d2x < P.cutoff2) {
                        n_in++;
                        const double bxy = b[0] * b[1], bxz = b[0] * b[2], byz = b[1] * b[2];
#pragma unroll
                        for (int g = 0; g < NP; g++) {
                            const PairDev& pr = P.table[g * P.n_types + ty_];
                            if ((pr.flags & 0xff) != 0) {
                                double v, p1, p2, p3;
                                vdw_pair_ref(pr, d2x, v, p1, p2, p3);
                                accumulate(acc[g], v, p1, p2, p3, b[0], b[1], b[2], bxy, bxz, byz);
                            }
                        }
                        if (COUL) {
                            close_hit |= d2x < 1.0;
                            double v, p1, p2, p3;
                            ewald_pair_ref(P.alpha, q, d2x, v, p1, p2, p3);
                            accumulate(acc[NG - 1], v, p1, p2, p3, b[0], b[1], b[2], bxy, bxz, byz);
                        }
                    }
                }
            } else {
                CEG_NEXT();
            }
#undef CEG_NEXT
            if (PIPE >= 1) { dx = dxn; dy = dyn; dz = dzn; d2 = d2n; seed = seedn; d2hi = d2hin; q = qn; mt = mtn; pr = prn; }
            if (PIPE == 2) { ir = irn; r = rn; z = zn; }
            if (RUNPTR) { pgx += g_stride; pgy += g_stride; pgz += g_stride; }
        }
    };

    // scan the images that can reach this tile's x planes (the list is sorted by x; ext_lo / ext_hi bracket, per
    // plane, the images within the cutoff along x), keep those within reach of the tile centre, compact in order
    const int pl0 = P.a0 + tx * TX - P.ext_plane0;
    const int pl1 = min(P.a0 + tx * TX + TX, P.a1) - 1 - P.ext_plane0;
    const int scan0 = P.ext_lo ? P.ext_lo[pl0] : 0;
    const int scan1 = P.ext_hi ? P.ext_hi[pl1] : P.n_ext;
    // the loads of chunk c + 1 are issued before chunk c is compacted (their latency is most of a chunk's time)
    ExtAtom a_n;
    ExtMeta m_n;
    bool have_n = scan0 + tid < scan1;
    if (have_n) { a_n = P.ext[scan0 + tid]; m_n = P.ext_meta[scan0 + tid]; }
    for (int base = scan0; base < scan1; base += NT) {
        const ExtAtom a = a_n;
        const ExtMeta m = m_n;
        const bool have = have_n;
        have_n = base + NT + tid < scan1;
        if (have_n) { a_n = P.ext[base + NT + tid]; m_n = P.ext_meta[base + NT + tid]; }
        bool keep = false;
        if (have) {
            const double ddx = cx - a.x, ddy = cy - a.y, ddz = cz - a.z;
            keep = (ddx * ddx + ddy * ddy + ddz * ddz) < P.reach2;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const int cw = s_wcount[w];
            if (w < warp) off += cw;
            total += cw;
        }
        if (count + total > cap) {   // this chunk does not fit any more: evaluate what is there first
            process(count);
            count = 0;
            __syncthreads();
        }
        if (keep) {
            const int slot = count + off + __popc(bal & ((1u << lane) - 1u));
            s_atom[slot] = a;
            s_meta[slot] = ListMeta{m.type | ((int)s_cls[m.type] << 16), sa_table + (uint32_t)m.type * (uint32_t)sizeof(PairHot)};
            s_orig[slot] = m.orig;
        }
        count += total;
        __syncthreads();
    }
    process(count);

    // stats: one atomic per warp
    if (P.stats) {
        unsigned long long t64 = valid ? (unsigned long long)n_listed : 0ull, i64 = n_in;
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
    if (COUL && close_hit) acc[NG - 1][0] = CUDART_INF;   // probes.jl:116

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
        if (P.out[g] || P.n_dest) {
            float o[8];
            set_gridpoint(acc[g], P.gs.delta, coul ? P.lam_c : P.lam_v, coul ? P.thr_c : P.thr_v, o);
            if (P.out[g]) {
#pragma unroll
                for (int c = 0; c < 8; c++) P.out[g][(int64_t)c * slab + node] = o[c];
            }
            if (P.n_dest) {
                const double sc = P.rec_scale[g];
                Rec8 r8;
                r8.lo = make_float4(record_value(o[0], sc), record_value(o[1], sc), record_value(o[2], sc), record_value(o[3], sc));
                r8.hi = make_float4(record_value(o[4], sc), record_value(o[5], sc), record_value(o[6], sc), record_value(o[7], sc));
                const int64_t d = record_index(P.rec_block != 0, i, j, k, (int)ny, (int)nz);
#pragma unroll 1
                for (int t = 0; t < P.n_dest; t++) P.rec[t][g][d] = r8;      // local, or a peer GPU's memory
            }
        }
    }
}

}  // namespace ceg
