// Guest–guest pair energies for batched Monte-Carlo trial evaluation (sm_100a).
//
// Replaces single_contribution_vdw_noneighbour / compute_vdw_noneighbour (energy.jl:339-427): the energy of
// one molecule placed at each of n_trials trial positions with every atom of the OTHER guest molecules,
// pair rule evaluated on d^2 as InteractionRule(::Å^2) does (interactions.jl:391-409): Lennard-Jones from
// r^2, hard sphere, and through sqrt: Buckingham, the real-space Ewald term CCF q1 q2 erfc(alpha r)/r
// (CoulombEwaldDirect) and bare Coulomb.  Distances use unsafe_periodic_distance2! (utils.jl:294-302: wrap,
// no image search) with un-contracted arithmetic, so the `d2 < cutoff2` decisions match the oracle bit for
// bit; the sum runs in the reference's order (molecule atom outer, environment atom inner).
//
// Round 2: the pair loop no longer runs the literal arithmetic for every pair (two un-fused 3x3 products and a
// floor-wrap, 47 FP64 operations, for a sum in which 5 of 6 pairs lie outside the cutoff).  Fractional coordinates are
// formed once per atom; a pair costs a subtraction, a rint-wrap and one fused 3x3 product (21 operations).  Only pairs
// whose d^2 lands within 1e-9 Å^2 of the cutoff are re-measured with the literal sequence, which then decides `d2 <
// cutoff2` bit for bit as before; in-cutoff pairs use the branch-free exp / erfcx of fastmath.cuh where their argument
// lies in the fitted range (1e-13 relative), libm elsewhere.  Cells narrower than twice the cutoff keep the literal loop.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace ceg {

struct PairEnergyDev {
    int flags;
    int _pad;
    double hs_r2;
    double lj_eps, lj_sig2, lj_shift;
    double bA, bB, bC, b_shift;
    double ew_alpha, ew_qq;       // qq = CCF * q1 * q2
    double coul_qq;
};

__device__ __forceinline__ double pair_energy(const PairEnergyDev& p, double r2) {
    double e = 0.0;
    const int fl = p.flags;
    // kind order of InteractionRuleSum: HardSphere, CoulombEwaldDirect, Coulomb, LennardJones, Buckingham
    if (fl & CEG_PAIR_HS) e += (r2 < p.hs_r2) ? CUDART_INF : 0.0;
    if (fl & (CEG_PAIR_EWALD | CEG_PAIR_COULOMB | CEG_PAIR_BUCK)) {
        const double r = sqrt(r2);
        if (fl & CEG_PAIR_EWALD) e += p.ew_qq * erfc(p.ew_alpha * r) / r;
        if (fl & CEG_PAIR_COULOMB) e += p.coul_qq / r;
        if (fl & CEG_PAIR_LJ) {
            const double t = p.lj_sig2 / r2;
            const double x6 = t * t * t;
            e += 4.0 * p.lj_eps * x6 * (x6 - 1.0) - p.lj_shift;
        }
        if (fl & CEG_PAIR_BUCK) {
            const double r6 = (r * r) * (r * r) * (r * r);
            e += (p.bA * exp(-p.bB * r) - p.bC / r6) - p.b_shift;
        }
    } else if (fl & CEG_PAIR_LJ) {
        const double t = p.lj_sig2 / r2;
        const double x6 = t * t * t;
        e += 4.0 * p.lj_eps * x6 * (x6 - 1.0) - p.lj_shift;
    }
    return e;
}

// rule(d^2) on the hot path: same terms and order as pair_energy, transcendental functions from fastmath.cuh
__device__ __forceinline__ double pair_energy_fast(const PairEnergyDev& p, double r2) {
    if (!(r2 > 1e-300)) return pair_energy(p, r2);      // coincident atoms: the library functions' Inf / NaN semantics
    double e = 0.0;
    const int fl = p.flags;
    if (fl & CEG_PAIR_HS) e += (r2 < p.hs_r2) ? CUDART_INF : 0.0;
    if (fl & (CEG_PAIR_EWALD | CEG_PAIR_COULOMB | CEG_PAIR_BUCK)) {
        const double ir = rsqrt_fast(r2);
        const double r = r2 * ir;
        if (fl & CEG_PAIR_EWALD) {
            const double x = p.ew_alpha * r;
            const double c = (x >= ERFCX_XMIN && x <= ERFCX_XMAX) ? exp_acc(-(x * x)) * erfcx_fast(x) : erfc(x);
            e += (p.ew_qq * c) * ir;
        }
        if (fl & CEG_PAIR_COULOMB) e += p.coul_qq * ir;
        if (fl & CEG_PAIR_LJ) {
            const double t = p.lj_sig2 * (ir * ir);
            const double x6 = t * t * t;
            e += 4.0 * p.lj_eps * x6 * (x6 - 1.0) - p.lj_shift;
        }
        if (fl & CEG_PAIR_BUCK) {
            const double ir2 = ir * ir;
            const double br = p.bB * r;
            const double ex = br < 700.0 ? exp_acc(-br) : exp(-br);
            e += (p.bA * ex - p.bC * (ir2 * ir2 * ir2)) - p.b_shift;
        }
    } else if (fl & CEG_PAIR_LJ) {
        const double t = p.lj_sig2 / r2;
        const double x6 = t * t * t;
        e += 4.0 * p.lj_eps * x6 * (x6 - 1.0) - p.lj_shift;
    }
    return e;
}

struct GuestParams {
    CellDev cell;
    double cutoff2;
    double cut_lo, cut_hi;          // cutoff2 -/+ GUEST_BAND (fast kernel)
    int fast_ok;                    // cell at least twice the cutoff wide: the wrapped image is the only one in reach
    int n_species;
    const PairEnergyDev* table;     // n_species x n_species
    int n_env;
    const double* env_xyz;          // n_env x 3
    const int* env_species;
    const int* env_mol;
    int64_t n_trials;
    int n_atoms;
    const int* mol_species;         // n_atoms
    int skip_mol;
    const double* trial_xyz;        // n_trials x n_atoms x 3
    double* out;
};

constexpr int GUEST_ENV_TILE = 256;
constexpr double GUEST_BAND = 1e-9;     // Å^2 around the cutoff inside which the literal sequence decides
// Every lane of a warp follows its own trial placement, so "this pair is inside the cutoff" (about one pair in six) is
// true for SOME lane in nearly every iteration: evaluated in place, the ~150-instruction rule(d^2) ran at one sixth lane
// efficiency and took 90 % of the kernel (ncu, round 2: 7.7 ms).  In-cutoff pairs are therefore queued per thread
// (d^2 and species, shared memory) and the warp evaluates them together whenever one lane's queue is full.  A thread's
// pairs are still summed in their original order.
constexpr int GUEST_QUEUE_LIT = 8;

__global__ void __launch_bounds__(128, 8) guest_energy_kernel(const GuestParams P) {
    __shared__ double s_xyz[GUEST_ENV_TILE][3];
    __shared__ double s_frac[GUEST_ENV_TILE][3];
    __shared__ int s_sp[GUEST_ENV_TILE];
    __shared__ int s_mol[GUEST_ENV_TILE];
    __shared__ double s_qd2[GUEST_QUEUE_LIT][128];
    __shared__ int s_qsp[GUEST_QUEUE_LIT][128];
    __shared__ CellDev s_cell;
    const int tid = threadIdx.x;
    if (tid == 0) s_cell = P.cell;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + tid;
    const bool valid = i < P.n_trials;
    double energy = 0.0;
    const double* M = P.cell.mat;
    const double* IM = P.cell.invmat;
    const double cut_lo = P.cutoff2 - GUEST_BAND, cut_hi = P.cutoff2 + GUEST_BAND;
    // The reference loops molecule atom (outer) x environment atom (inner); the environment is staged
    // through shared memory in tiles, so for more than GUEST_ENV_TILE atoms the outer/inner order is
    // tile-blocked (a different, still deterministic, summation order).
    for (int base = 0; base < P.n_env; base += GUEST_ENV_TILE) {
        const int cnt = min(GUEST_ENV_TILE, P.n_env - base);
        __syncthreads();
        for (int t = tid; t < cnt; t += blockDim.x) {
            const double x = P.env_xyz[3 * (base + t)], y = P.env_xyz[3 * (base + t) + 1], z = P.env_xyz[3 * (base + t) + 2];
            s_xyz[t][0] = x; s_xyz[t][1] = y; s_xyz[t][2] = z;
#pragma unroll
            for (int d = 0; d < 3; d++) s_frac[t][d] = fma(IM[d], x, fma(IM[3 + d], y, IM[6 + d] * z));
            s_sp[t] = P.env_species[base + t];
            s_mol[t] = P.env_mol[base + t];
        }
        __syncthreads();
        for (int k2 = 0; k2 < P.n_atoms; k2++) {       // (all lanes walk the loop: the queue is drained by warp votes)
            double x2 = 0.0, y2 = 0.0, z2 = 0.0;
            if (valid) {
                const double* p2 = P.trial_xyz + ((int64_t)i * P.n_atoms + k2) * 3;
                x2 = p2[0]; y2 = p2[1]; z2 = p2[2];
            }
            const double fx = fma(IM[0], x2, fma(IM[3], y2, IM[6] * z2));
            const double fy = fma(IM[1], x2, fma(IM[4], y2, IM[7] * z2));
            const double fz = fma(IM[2], x2, fma(IM[5], y2, IM[8] * z2));
            const PairEnergyDev* row = P.table + (size_t)P.mol_species[k2] * P.n_species;
            int nq = 0;
            auto drain = [&]() {
                for (int j = 0; __any_sync(0xffffffffu, j < nq); j++)
                    if (j < nq) energy += pair_energy_fast(row[s_qsp[j][tid]], s_qd2[j][tid]);
                nq = 0;
            };
            for (int l1 = 0; l1 < cnt; l1++) {
                bool in = false;
                double d2 = 0.0;
                if (valid && s_mol[l1] != P.skip_mol) {
                    bool literal = !P.fast_ok;
                    if (!literal) {
                        double ax = fx - s_frac[l1][0], ay = fy - s_frac[l1][1], az = fz - s_frac[l1][2];
                        ax -= rint(ax); ay -= rint(ay); az -= rint(az);
                        const double vx = fma(M[0], ax, fma(M[3], ay, M[6] * az));
                        const double vy = fma(M[1], ax, fma(M[4], ay, M[7] * az));
                        const double vz = fma(M[2], ax, fma(M[5], ay, M[8] * az));
                        d2 = fma(vx, vx, fma(vy, vy, vz * vz));
                        in = d2 <= cut_lo;
                        literal = !in && d2 < cut_hi;
                    }
                    if (literal) {
                        // buffer = pos2 - pos1; unsafe_periodic_distance2! (utils.jl:294-302), un-contracted: decides exactly
                        double b[3] = {xsub(x2, s_xyz[l1][0]), xsub(y2, s_xyz[l1][1]), xsub(z2, s_xyz[l1][2])};
                        double f[3];
                        matvec_x(s_cell.invmat, b, f);
#pragma unroll
                        for (int d = 0; d < 3; d++) {
                            const double diff = xadd(f[d], 0.5);
                            f[d] = xsub(xsub(diff, floor(diff)), 0.5);
                        }
                        matvec_x(s_cell.mat, f, b);
                        d2 = norm2_x(b);
                        in = d2 < P.cutoff2;
                    }
                }
                if (in) {
                    s_qd2[nq][tid] = d2;
                    s_qsp[nq][tid] = s_sp[l1];
                    nq++;
                }
                if (__any_sync(0xffffffffu, nq == GUEST_QUEUE_LIT)) drain();
            }
            drain();
        }
    }
    if (valid) P.out[i] = energy;
}

// ---- fast variant (cell at least twice the cutoff wide: the wrapped image is the only one in reach) -------------------
// ncu of the kernel above (round 2, 10^5 trial CO2 x 3000 guests): 137 warp instructions per pair test of which 28 FP64,
// 31 % occupancy (one thread per trial placement = 10^5 threads), a vote and two divergence barriers per test.  Here
//  * a trial placement is shared by `tpt` threads, thread p taking molecule atoms p, p + tpt, ... (3 x the threads for
//    CO2); the partial sums are added in thread order by the first of them — a fixed order, not the reference's
//    (atom-major over ALL pairs), within the 1e-9 tolerance like the tile-blocked order above;
//  * environment atoms are staged as 32-byte {fractional xyz, species, molecule} records (two broadcast LDS.128);
//  * the cell matrix lives in registers, four tests are issued per loop trip and the queue is checked once per trip;
//  * a queue entry carries the pair-table index, so the queue survives the change of molecule atom and tile and is
//    drained only when a lane runs out of room.
struct alignas(16) GuestEnvAtom {
    double fx, fy, fz;
    int sp, mol;
};
constexpr int GUEST_THREADS = 128;
#ifndef CEG_GUEST_QUEUE
#define CEG_GUEST_QUEUE 16
#endif
constexpr int GUEST_QUEUE = CEG_GUEST_QUEUE;      // in-cutoff pairs a thread can hold before the warp evaluates them
#ifndef CEG_GUEST_UNROLL
#define CEG_GUEST_UNROLL 4
#endif
constexpr int GUEST_UNROLL = CEG_GUEST_UNROLL;
#ifndef CEG_GUEST_MINB
#define CEG_GUEST_MINB 6
#endif

// The rule(d^2) code is ~150 instructions and wants ~60 registers of its own: kept out of line so that the test loop
// around it keeps the cell matrix and the atom's fractional coordinates in registers (inlined, the loop re-read them
// from local memory and the constant bank for every pair: 137 -> 104 instructions per test were still mostly that).
// qd2 / qix: this thread's queue column (stride GUEST_THREADS).
__device__ __noinline__ double guest_drain(const PairEnergyDev* __restrict__ table, const double* qd2, const int* qix, int nq,
                                           double energy) {
    for (int j = 0; __any_sync(0xffffffffu, j < nq); j++)
        if (j < nq) energy += pair_energy_fast(table[qix[j * GUEST_THREADS]], qd2[j * GUEST_THREADS]);
    return energy;
}

// TRI: the cell matrix is upper triangular (a along x, b in the xy plane — every cell the RASPA loader produces): the three
// products with its zeros are dropped.
template <bool TRI>
__global__ void __launch_bounds__(GUEST_THREADS, CEG_GUEST_MINB) guest_energy_fast_kernel(const GuestParams P, const int tpt) {
    __shared__ GuestEnvAtom s_env[GUEST_ENV_TILE];
    __shared__ double s_qd2[GUEST_QUEUE + 1][GUEST_THREADS];       // + 1: the slot an out-of-cutoff pair is written to
    __shared__ int s_qix[GUEST_QUEUE + 1][GUEST_THREADS];
    __shared__ double s_part[GUEST_THREADS];
    __shared__ CellDev s_cell;
    const int tid = threadIdx.x;
    if (tid == 0) s_cell = P.cell;
    const int trials_per_cta = GUEST_THREADS / tpt;
    const int lt = tid / tpt, part = tid - lt * tpt;
    const int64_t i = (int64_t)blockIdx.x * trials_per_cta + lt;
    const bool valid = lt < trials_per_cta && i < P.n_trials;
    const double m0 = P.cell.mat[0], m1 = P.cell.mat[1], m2 = P.cell.mat[2], m3 = P.cell.mat[3], m4 = P.cell.mat[4],
                 m5 = P.cell.mat[5], m6 = P.cell.mat[6], m7 = P.cell.mat[7], m8 = P.cell.mat[8];
    const double* IM = P.cell.invmat;
    const double cut_lo = P.cut_lo, cut_hi = P.cut_hi;
    const int skip_mol = P.skip_mol;
    const int k_trips = (P.n_atoms + tpt - 1) / tpt;
    double* const qd2 = &s_qd2[0][tid];
    int* const qix = &s_qix[0][tid];
    double energy = 0.0;
    int nq = 0;
    for (int base = 0; base < P.n_env; base += GUEST_ENV_TILE) {
        const int cnt = min(GUEST_ENV_TILE, P.n_env - base);
        const int cnt_pad = (cnt + GUEST_UNROLL - 1) / GUEST_UNROLL * GUEST_UNROLL;     // <= GUEST_ENV_TILE (a multiple of it)
        __syncthreads();
        for (int t = tid; t < cnt_pad; t += GUEST_THREADS) {
            GuestEnvAtom e;
            if (t < cnt) {
                const double x = P.env_xyz[3 * (base + t)], y = P.env_xyz[3 * (base + t) + 1], z = P.env_xyz[3 * (base + t) + 2];
                e.fx = fma(IM[0], x, fma(IM[3], y, IM[6] * z));
                e.fy = fma(IM[1], x, fma(IM[4], y, IM[7] * z));
                e.fz = fma(IM[2], x, fma(IM[5], y, IM[8] * z));
                e.sp = P.env_species[base + t];
                e.mol = P.env_mol[base + t];
            } else {                    // padding up to the unroll factor: a NaN distance fails every comparison
                e.fx = e.fy = e.fz = CUDART_NAN;
                e.sp = 0;
                e.mol = skip_mol;
            }
            s_env[t] = e;
        }
        __syncthreads();
        for (int kk = 0; kk < k_trips; kk++) {            // every lane walks every loop: the queue is drained by warp votes
            const int k2 = part + kk * tpt;
            const bool active = valid && k2 < P.n_atoms;
            double x2 = CUDART_NAN, y2 = CUDART_NAN, z2 = CUDART_NAN;      // an idle lane measures NaN distances
            int rowbase = 0;
            if (active) {
                const double* p2 = P.trial_xyz + ((int64_t)i * P.n_atoms + k2) * 3;
                x2 = p2[0]; y2 = p2[1]; z2 = p2[2];
                rowbase = P.mol_species[k2] * P.n_species;
            }
            const double fx = fma(IM[0], x2, fma(IM[3], y2, IM[6] * z2));
            const double fy = fma(IM[1], x2, fma(IM[4], y2, IM[7] * z2));
            const double fz = fma(IM[2], x2, fma(IM[5], y2, IM[8] * z2));
            for (int l0 = 0; l0 < cnt_pad; l0 += GUEST_UNROLL) {
                bool band_any = false;
#pragma unroll
                for (int u = 0; u < GUEST_UNROLL; u++) {
                    const GuestEnvAtom e = s_env[l0 + u];
                    double ax = fx - e.fx, ay = fy - e.fy, az = fz - e.fz;
                    ax -= rint(ax); ay -= rint(ay); az -= rint(az);
                    const double vx = fma(m0, ax, fma(m3, ay, m6 * az));
                    const double vy = TRI ? fma(m4, ay, m7 * az) : fma(m1, ax, fma(m4, ay, m7 * az));
                    const double vz = TRI ? m8 * az : fma(m2, ax, fma(m5, ay, m8 * az));
                    const double d2 = fma(vx, vx, fma(vy, vy, vz * vz));
                    const bool ok = e.mol != skip_mol;
                    const bool in = ok && d2 <= cut_lo;
                    band_any |= ok && d2 > cut_lo && d2 < cut_hi;
                    qd2[nq * GUEST_THREADS] = d2;                     // unconditional store, conditional advance
                    qix[nq * GUEST_THREADS] = rowbase + e.sp;
                    nq += in ? 1 : 0;
                }
                if (__any_sync(0xffffffffu, band_any)) {
                    // a pair within 1e-9 A^2 of the cutoff (rare): unsafe_periodic_distance2! (utils.jl:294-302), un-contracted,
                    // decides — and supplies the distance — for the band pairs of this trip
                    energy = guest_drain(P.table, qd2, qix, nq, energy);      // room for up to GUEST_UNROLL more entries
                    nq = 0;
                    for (int u = 0; u < GUEST_UNROLL; u++) {
                        const GuestEnvAtom e = s_env[l0 + u];
                        double ax = fx - e.fx, ay = fy - e.fy, az = fz - e.fz;
                        ax -= rint(ax); ay -= rint(ay); az -= rint(az);
                        const double vx = fma(m0, ax, fma(m3, ay, m6 * az));
                        const double vy = TRI ? fma(m4, ay, m7 * az) : fma(m1, ax, fma(m4, ay, m7 * az));
                        const double vz = TRI ? m8 * az : fma(m2, ax, fma(m5, ay, m8 * az));
                        const double d2f = fma(vx, vx, fma(vy, vy, vz * vz));
                        if (e.mol != skip_mol && d2f > cut_lo && d2f < cut_hi) {
                            const double* pe = P.env_xyz + 3 * (int64_t)(base + l0 + u);
                            double b[3] = {xsub(x2, pe[0]), xsub(y2, pe[1]), xsub(z2, pe[2])};
                            double f[3];
                            matvec_x(s_cell.invmat, b, f);
#pragma unroll
                            for (int d = 0; d < 3; d++) {
                                const double diff = xadd(f[d], 0.5);
                                f[d] = xsub(xsub(diff, floor(diff)), 0.5);
                            }
                            matvec_x(s_cell.mat, f, b);
                            const double d2 = norm2_x(b);
                            if (d2 < P.cutoff2) {
                                // the pair was left out above; appended here it keeps the thread's pair set, the order of its
                                // sum moves by at most GUEST_UNROLL - 1 places
                                qd2[nq * GUEST_THREADS] = d2;
                                qix[nq * GUEST_THREADS] = rowbase + e.sp;
                                nq++;
                            }
                        }
                    }
                }
                if (__any_sync(0xffffffffu, nq > GUEST_QUEUE - GUEST_UNROLL - 1)) {
                    energy = guest_drain(P.table, qd2, qix, nq, energy);
                    nq = 0;
                }
            }
        }
    }
    energy = guest_drain(P.table, qd2, qix, nq, energy);
    s_part[tid] = energy;
    __syncthreads();
    if (valid && part == 0) {
        double e = 0.0;
        for (int p = 0; p < tpt; p++) e += s_part[tid + p];
        P.out[i] = e;
    }
}

}  // namespace ceg
