// Ewald reciprocal-space kernels (FP64, sm_100a).
//
// Replaces setup_Eik + ewald_main_loop! (ewald.jl:109-185), move_one_system! (ewald.jl:350-364)
// and compute_ewald(ctx) (ewald.jl:555-577).
//
//  * structure_factor_kernel: S_f(k) = sum_sites q exp(2 pi i (ix fx + jy fy + jz fz)).  One thread
//    per k-vector and per block of sites; sites are staged through shared memory, the phase factor
//    comes from sincospi of the exactly-reduced phase; block partials are written out and reduced
//    in a fixed order by structure_factor_reduce_kernel (deterministic, no atomics).
//  * reciprocal_kernel: one warp per placement of the probe molecule.  The per-atom power tables
//    Eikx[0..kx], Eiky[-ky..ky], Eikz[-kz..kz] (the reference builds them by recurrence from one
//    cispi per axis) are evaluated directly with sincospi by the lanes into shared memory, the lanes
//    then stride over the k-list accumulating kf*Re(S_f conj S_a) and kf*|S_a|^2, and a warp-shuffle
//    reduction yields the energy of the placement.
#pragma once
#include "common.cuh"

namespace ceg {

struct KVecDev {       // one k-vector
    double kf;         // kfactor (ewald.jl:247-259)
    double sf_re, sf_im;
    int ix, jy, jz, _pad;
};

// (m * f) mod 1 with the rounding error of the product carried along (m integer, f in [0,1))
__device__ __forceinline__ double red_prod(double m, double f) {
    const double hi = m * f;
    const double lo = fma(m, f, -hi);
    double r = (hi - floor(hi)) + lo;
    return r - floor(r);
}

constexpr int SF_SITES_PER_BLOCK = 256;
constexpr int SF_THREADS = 128;

// partial[blockIdx.y][k] over a block of sites
__global__ void __launch_bounds__(SF_THREADS) structure_factor_kernel(int64_t n_k, const int* __restrict__ kvec,
                                                                      int64_t n_sites, const double* __restrict__ frac,
                                                                      const double* __restrict__ charge,
                                                                      double2* __restrict__ partial) {
    __shared__ double s_f[SF_SITES_PER_BLOCK][3];
    __shared__ double s_q[SF_SITES_PER_BLOCK];
    const int64_t s0 = (int64_t)blockIdx.y * SF_SITES_PER_BLOCK;
    const int ns = (int)min((int64_t)SF_SITES_PER_BLOCK, n_sites - s0);
    for (int s = threadIdx.x; s < ns; s += blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double f = frac[3 * (s0 + s) + d];
            s_f[s][d] = f - floor(f);          // exp(2 pi i m f) is periodic in f for integer m
        }
        s_q[s] = charge[s0 + s];
    }
    __syncthreads();
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_k) return;
    const double ix = kvec[3 * k], jy = kvec[3 * k + 1], jz = kvec[3 * k + 2];
    double re = 0.0, im = 0.0;
    for (int s = 0; s < ns; s++) {
        // each product is exact-ish (|m| <= ~60, f in [0,1)); reduce the phase before sincospi
        double ph = red_prod(ix, s_f[s][0]) + red_prod(jy, s_f[s][1]) + red_prod(jz, s_f[s][2]);
        ph -= floor(ph);
        double sn, cs;
        sincospi(2.0 * ph, &sn, &cs);
        re += s_q[s] * cs;
        im += s_q[s] * sn;
    }
    partial[(int64_t)blockIdx.y * n_k + k] = make_double2(re, im);
}

__global__ void __launch_bounds__(128) structure_factor_reduce_kernel(int64_t n_k, int n_blocks,
                                                                      const double2* __restrict__ partial,
                                                                      const double* __restrict__ kfactor,
                                                                      const int* __restrict__ kvec,
                                                                      KVecDev* __restrict__ kv, double* __restrict__ S_out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_k) return;
    double re = 0.0, im = 0.0;
    for (int b = 0; b < n_blocks; b++) {
        const double2 p = partial[(int64_t)b * n_k + k];
        re += p.x;
        im += p.y;
    }
    KVecDev o;
    o.kf = kfactor[k];
    o.sf_re = re;
    o.sf_im = im;
    o.ix = kvec[3 * k]; o.jy = kvec[3 * k + 1]; o.jz = kvec[3 * k + 2]; o._pad = 0;
    kv[k] = o;
    S_out[2 * k] = re;
    S_out[2 * k + 1] = im;
}

constexpr int RECIP_QV = 32;
struct RecipParams {
    double invmat[9];      // supercell inverse, column-major
    int kx, ky, kz;
    int64_t n_k;
    const KVecDev* kv;
    int64_t n_pos;
    int n_atoms;
    const double* xyz;     // n_pos x n_atoms x 3
    const double* q;       // n_atoms (device); null: the charges travel by value in qv
    double qv[RECIP_QV];   // (no handle-owned buffer that a second call on another stream could overwrite)
    double energy_net_charges, static_contribution;
    double* out;
    const double2* rest;   // optional: replaces S_f (single_contribution_ewald's `rest`, ewald.jl:722-729)
    const uint8_t* skip;   // optional: placements flagged here are left untouched
    int accumulate;        // 1: out[i] += E (fused energy_point), 0: out[i] = E
};

constexpr int RECIP_WARPS = 4;

// dynamic shared memory: per warp n_atoms * (kx+1 + 2ky+1 + 2kz+1) double2 + n_atoms doubles
__global__ void __launch_bounds__(RECIP_WARPS * 32) reciprocal_kernel(const RecipParams P) {
    extern __shared__ double2 s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nx = P.kx + 1, ny = 2 * P.ky + 1, nz = 2 * P.kz + 1;
    const int per_atom = nx + ny + nz;
    double2* tab = s_dyn + (size_t)warp * per_atom * P.n_atoms;
    const int64_t wid = (int64_t)blockIdx.x * RECIP_WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * RECIP_WARPS;
    for (int64_t i = wid; i < P.n_pos; i += nwarps) {
        if (P.skip && P.skip[i]) continue;
        const double* mol = P.xyz + (int64_t)3 * P.n_atoms * i;
        __syncwarp();
        // power tables, one entry per lane-iteration: entry t of atom a
        for (int t = lane; t < per_atom * P.n_atoms; t += 32) {
            const int a = t / per_atom, u = t - a * per_atom;
            const double x = mol[3 * a], y = mol[3 * a + 1], z = mol[3 * a + 2];
            int axis, m;
            if (u < nx) { axis = 0; m = u; }
            else if (u < nx + ny) { axis = 1; m = u - nx - P.ky; }
            else { axis = 2; m = u - nx - ny - P.kz; }
            // fractional coordinate along this axis: row `axis` of invmat * pos (ewald.jl:358)
            double f = P.invmat[axis] * x + P.invmat[3 + axis] * y + P.invmat[6 + axis] * z;
            f -= floor(f);
            double ph = red_prod((double)m, f);
            double sn, cs;
            sincospi(2.0 * ph, &sn, &cs);
            tab[t] = make_double2(cs, sn);
        }
        __syncwarp();
        double fa = 0.0, aa = 0.0;
        for (int64_t k = lane; k < P.n_k; k += 32) {
            const KVecDev kv = P.kv[k];
            double sre = 0.0, sim = 0.0;
            for (int a = 0; a < P.n_atoms; a++) {
                const double2* ta = tab + a * per_atom;
                const double2 ex = ta[kv.ix], ey = ta[nx + P.ky + kv.jy], ez = ta[nx + ny + P.kz + kv.jz];
                const double yzr = ey.x * ez.x - ey.y * ez.y, yzi = ey.x * ez.y + ey.y * ez.x;
                const double qa = P.q ? P.q[a] : P.qv[a];
                sre += qa * (ex.x * yzr - ex.y * yzi);
                sim += qa * (ex.x * yzi + ex.y * yzr);
            }
            const double2 sf = P.rest ? P.rest[k] : make_double2(kv.sf_re, kv.sf_im);
            fa += kv.kf * (sf.x * sre + sf.y * sim);
            aa += kv.kf * (sre * sre + sim * sim);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fa += __shfl_xor_sync(0xffffffffu, fa, o);
            aa += __shfl_xor_sync(0xffffffffu, aa, o);
        }
        if (lane == 0) {
            // ewald.jl:572-576
            const double host = 2.0 * (fa + P.energy_net_charges);
            const double ads = aa + P.static_contribution;
            const double E = host + ads;
            if (P.accumulate) P.out[i] += E; else P.out[i] = E;
        }
    }
}

// ---- thread-per-placement variant ------------------------------------------------------------------
// The k-list of ewald.jl:213-236 is a set of rows: for each (jy, jz) a contiguous range of ix.  One thread
// owns one placement and walks the full (2ky+1) x (2kz+1) rectangle of rows in a fixed order, so that every
// phase factor comes from a recurrence (the reference builds its tables the same way, ewald.jl:69-86):
//   ey <- ey * e^{2 pi i fy} per jy,  w = q ey ez <- w * e^{2 pi i fz} per row,  t = w Ex[ix] <- t * e^{2 pi i fx} per k.
// No tables, no shuffles; the k data (kfactor, kfactor * S_f) is read at warp-uniform addresses.  Per k-vector
// and atom 4 FP64 instructions (+5 per k-vector) against 16 for the warp-per-placement kernel above, which
// stays as the general fallback (molecules of more than 4 atoms, k-lists that are not made of rows).
struct RowDev {
    int ix0, count;        // first ix and number of k-vectors of the row (0: row absent from the list)
    int64_t offset;        // index of the row's first k-vector in the list
};

struct RowsParams {
    double invmat[9];
    int ky, kz;
    const RowDev* rows;    // (2ky+1) x (2kz+1), jy-major
    const double* kf;      // kfactor per k-vector (list order)
    const double2* ksf;    // kfactor * S_f (or kfactor * rest)
    double sum_kf;         // sum of the kfactors in list order (single-atom molecules: |S_a|^2 = q^2 for every k)
    int64_t n_pos;
    const double* xyz;
    const double* q;       // device; null: qv
    double qv[RECIP_QV];
    int n_atoms;           // grouped kernel: atoms of the molecule (the plain kernel's count is its template argument)
    double energy_net_charges, static_contribution;
    double* out;
    const uint8_t* skip;
    int accumulate;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cis2pi(double ph) {   // exp(2 pi i ph), ph in [0, 1)
    double sn, cs;
    sincospi(2.0 * ph, &sn, &cs);
    return make_double2(cs, sn);
}

__global__ void scale_structure_factor_kernel(int64_t n_k, const double* __restrict__ kf, const double2* __restrict__ sf,
                                              double2* __restrict__ ksf) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_k) ksf[k] = make_double2(kf[k] * sf[k].x, kf[k] * sf[k].y);
}

constexpr int ROWS_THREADS = 128;
constexpr int ROWS_MAX_ATOMS = 4;

template <int NA>
__global__ void __launch_bounds__(ROWS_THREADS) reciprocal_rows_kernel(const RowsParams P) {
    const int64_t i = blockIdx.x * (int64_t)ROWS_THREADS + threadIdx.x;
    if (i >= P.n_pos) return;
    if (P.skip && P.skip[i]) return;
    const int ny = 2 * P.ky + 1, nz = 2 * P.kz + 1;
    double2 e1x[NA], e1y[NA], e1z[NA], ey[NA], ez0[NA], w[NA], t[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) {
        const double* r = P.xyz + ((int64_t)NA * i + a) * 3;
        const double x = r[0], y = r[1], z = r[2];
        double f[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {   // row d of invmat * pos (ewald.jl:358), reduced to [0, 1)
            const double v = P.invmat[d] * x + P.invmat[3 + d] * y + P.invmat[6 + d] * z;
            f[d] = v - floor(v);
        }
        e1x[a] = cis2pi(f[0]);
        e1y[a] = cis2pi(f[1]);
        e1z[a] = cis2pi(f[2]);
        ey[a] = cis2pi(red_prod(-(double)P.ky, f[1]));
        const double2 ez = cis2pi(red_prod(-(double)P.kz, f[2]));
        const double qa = P.q ? P.q[a] : P.qv[a];
        ez0[a] = make_double2(qa * ez.x, qa * ez.y);
    }
    double fa = 0.0, aa = 0.0;
    const RowDev* row = P.rows;
    for (int jy = 0; jy < ny; jy++) {
#pragma unroll
        for (int a = 0; a < NA; a++) w[a] = cmul(ey[a], ez0[a]);
        for (int jz = 0; jz < nz; jz++, row++) {
            const RowDev rd = *row;
            if (rd.count > 0) {
#pragma unroll
                for (int a = 0; a < NA; a++) t[a] = w[a];
                for (int s = 0; s < rd.ix0; s++) {
#pragma unroll
                    for (int a = 0; a < NA; a++) t[a] = cmul(t[a], e1x[a]);
                }
                const double* kf = P.kf + rd.offset;
                const double2* ksf = P.ksf + rd.offset;
#pragma unroll 4
                for (int c = 0; c < rd.count; c++) {
                    double sre = t[0].x, sim = t[0].y;
#pragma unroll
                    for (int a = 1; a < NA; a++) { sre += t[a].x; sim += t[a].y; }
                    const double2 ks = ksf[c];
                    fa = fma(ks.x, sre, fma(ks.y, sim, fa));
                    if (NA > 1) aa = fma(kf[c], fma(sim, sim, sre * sre), aa);
                    // unconditional (one wasted product per row): a select per k costs more issue slots
#pragma unroll
                    for (int a = 0; a < NA; a++) t[a] = cmul(t[a], e1x[a]);
                }
            }
            if (jz + 1 < nz) {
#pragma unroll
                for (int a = 0; a < NA; a++) w[a] = cmul(w[a], e1z[a]);
            }
        }
        if (jy + 1 < ny) {
#pragma unroll
            for (int a = 0; a < NA; a++) ey[a] = cmul(ey[a], e1y[a]);
        }
    }
    // one atom: S_a(k) = q e^{i phase}, so sum_k kf |S_a|^2 = q^2 sum_k kf (equal to the reference's abs2 sum to
    // rounding, 1e-16 relative)
    if (NA == 1) { const double q0 = P.q ? P.q[0] : P.qv[0]; aa = (q0 * q0) * P.sum_kf; }
    // ewald.jl:572-576
    const double E = 2.0 * (fa + P.energy_net_charges) + (aa + P.static_contribution);
    if (P.accumulate) P.out[i] += E; else P.out[i] = E;
}

// Molecules of 5 to 32 atoms: the placement is shared by G = 2, 4 or 8 adjacent lanes, lane g carrying atoms 4g .. 4g + 3
// through the same row walk as above (absent atoms: charge 0); per k-vector the lanes' partial S_a are summed by
// xor-shuffles, so every lane of the group holds the molecule's S_a(k) and accumulates the same energy.  (The
// warp-per-placement kernel spends 4 times as many FP64 instructions per (atom, k).)
template <int G>
__global__ void __launch_bounds__(ROWS_THREADS) reciprocal_rows_group_kernel(const RowsParams P) {
    constexpr int NA = 4;
    const int64_t tg = blockIdx.x * (int64_t)ROWS_THREADS + threadIdx.x;
    const int64_t i_raw = tg / G;
    const int g = (int)(tg - i_raw * G);
    const bool live = i_raw < P.n_pos;                       // every lane walks the loops: the shuffles need full warps
    const int64_t i = live ? i_raw : P.n_pos - 1;
    const int ny = 2 * P.ky + 1, nz = 2 * P.kz + 1;
    double2 e1x[NA], e1y[NA], e1z[NA], ey[NA], ez0[NA], w[NA], t[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) {
        const int ag = 4 * g + a;
        const bool have = ag < P.n_atoms;
        const double* r = P.xyz + ((int64_t)P.n_atoms * i + (have ? ag : 0)) * 3;
        const double x = r[0], y = r[1], z = r[2];
        double f[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double v = P.invmat[d] * x + P.invmat[3 + d] * y + P.invmat[6 + d] * z;
            f[d] = v - floor(v);
        }
        e1x[a] = cis2pi(f[0]);
        e1y[a] = cis2pi(f[1]);
        e1z[a] = cis2pi(f[2]);
        ey[a] = cis2pi(red_prod(-(double)P.ky, f[1]));
        const double2 ez = cis2pi(red_prod(-(double)P.kz, f[2]));
        const double qa = have ? (P.q ? P.q[ag] : P.qv[ag]) : 0.0;
        ez0[a] = make_double2(qa * ez.x, qa * ez.y);
    }
    double fa = 0.0, aa = 0.0;
    const RowDev* row = P.rows;
    for (int jy = 0; jy < ny; jy++) {
#pragma unroll
        for (int a = 0; a < NA; a++) w[a] = cmul(ey[a], ez0[a]);
        for (int jz = 0; jz < nz; jz++, row++) {
            const RowDev rd = *row;
            if (rd.count > 0) {
#pragma unroll
                for (int a = 0; a < NA; a++) t[a] = w[a];
                for (int s = 0; s < rd.ix0; s++) {
#pragma unroll
                    for (int a = 0; a < NA; a++) t[a] = cmul(t[a], e1x[a]);
                }
                const double* kf = P.kf + rd.offset;
                const double2* ksf = P.ksf + rd.offset;
                for (int c = 0; c < rd.count; c++) {
                    double sre = (t[0].x + t[1].x) + (t[2].x + t[3].x), sim = (t[0].y + t[1].y) + (t[2].y + t[3].y);
#pragma unroll
                    for (int o = 1; o < G; o <<= 1) {
                        sre += __shfl_xor_sync(0xffffffffu, sre, o);
                        sim += __shfl_xor_sync(0xffffffffu, sim, o);
                    }
                    const double2 ks = ksf[c];
                    fa = fma(ks.x, sre, fma(ks.y, sim, fa));
                    aa = fma(kf[c], fma(sim, sim, sre * sre), aa);
#pragma unroll
                    for (int a = 0; a < NA; a++) t[a] = cmul(t[a], e1x[a]);
                }
            }
            if (jz + 1 < nz) {
#pragma unroll
                for (int a = 0; a < NA; a++) w[a] = cmul(w[a], e1z[a]);
            }
        }
        if (jy + 1 < ny) {
#pragma unroll
            for (int a = 0; a < NA; a++) ey[a] = cmul(ey[a], e1y[a]);
        }
    }
    if (!live || g != 0 || (P.skip && P.skip[i])) return;
    const double E = 2.0 * (fa + P.energy_net_charges) + (aa + P.static_contribution);      // ewald.jl:572-576
    if (P.accumulate) P.out[i] += E; else P.out[i] = E;
}

}  // namespace ceg
