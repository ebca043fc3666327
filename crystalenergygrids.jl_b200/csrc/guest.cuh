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
constexpr int GUEST_QUEUE = 8;

__global__ void __launch_bounds__(128, 8) guest_energy_kernel(const GuestParams P) {
    __shared__ double s_xyz[GUEST_ENV_TILE][3];
    __shared__ double s_frac[GUEST_ENV_TILE][3];
    __shared__ int s_sp[GUEST_ENV_TILE];
    __shared__ int s_mol[GUEST_ENV_TILE];
    __shared__ double s_qd2[GUEST_QUEUE][128];
    __shared__ int s_qsp[GUEST_QUEUE][128];
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
                if (__any_sync(0xffffffffu, nq == GUEST_QUEUE)) drain();
            }
            drain();
        }
    }
    if (valid) P.out[i] = energy;
}

}  // namespace ceg
