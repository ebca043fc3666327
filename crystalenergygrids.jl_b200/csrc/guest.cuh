// Guest–guest pair energies for batched Monte-Carlo trial evaluation (sm_100a).
//
// Replaces single_contribution_vdw_noneighbour / compute_vdw_noneighbour (energy.jl:339-427): the energy of
// one molecule placed at each of n_trials trial positions with every atom of the OTHER guest molecules,
// pair rule evaluated on d^2 as InteractionRule(::Å^2) does (interactions.jl:391-409): Lennard-Jones from
// r^2, hard sphere, and through sqrt: Buckingham, the real-space Ewald term CCF q1 q2 erfc(alpha r)/r
// (CoulombEwaldDirect) and bare Coulomb.  Distances use unsafe_periodic_distance2! (utils.jl:294-302: wrap,
// no image search) with un-contracted arithmetic, so the `d2 < cutoff2` decisions match the oracle bit for
// bit; the sum runs in the reference's order (molecule atom outer, environment atom inner).
#pragma once
#include "common.cuh"

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

struct GuestParams {
    CellDev cell;
    double cutoff2;
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

constexpr int GUEST_ENV_TILE = 1024;

__global__ void __launch_bounds__(128) guest_energy_kernel(const GuestParams P) {
    __shared__ double s_xyz[GUEST_ENV_TILE][3];
    __shared__ int s_sp[GUEST_ENV_TILE];
    __shared__ int s_mol[GUEST_ENV_TILE];
    __shared__ CellDev s_cell;
    if (threadIdx.x == 0) s_cell = P.cell;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = i < P.n_trials;
    double energy = 0.0;
    // The reference loops molecule atom (outer) x environment atom (inner); the environment is staged
    // through shared memory in tiles, so for more than GUEST_ENV_TILE atoms the outer/inner order is
    // tile-blocked (a different, still deterministic, summation order).
    for (int base = 0; base < P.n_env; base += GUEST_ENV_TILE) {
        const int cnt = min(GUEST_ENV_TILE, P.n_env - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
            s_xyz[t][0] = P.env_xyz[3 * (base + t)];
            s_xyz[t][1] = P.env_xyz[3 * (base + t) + 1];
            s_xyz[t][2] = P.env_xyz[3 * (base + t) + 2];
            s_sp[t] = P.env_species[base + t];
            s_mol[t] = P.env_mol[base + t];
        }
        __syncthreads();
        if (!valid) continue;
        for (int k2 = 0; k2 < P.n_atoms; k2++) {
            const double* p2 = P.trial_xyz + ((int64_t)i * P.n_atoms + k2) * 3;
            const double x2 = p2[0], y2 = p2[1], z2 = p2[2];
            const PairEnergyDev* row = P.table + (size_t)P.mol_species[k2] * P.n_species;
            for (int l1 = 0; l1 < cnt; l1++) {
                if (s_mol[l1] == P.skip_mol) continue;
                // buffer = pos2 - pos1; unsafe_periodic_distance2! (utils.jl:294-302)
                double b[3] = {xsub(x2, s_xyz[l1][0]), xsub(y2, s_xyz[l1][1]), xsub(z2, s_xyz[l1][2])};
                double f[3];
                matvec_x(s_cell.invmat, b, f);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double diff = xadd(f[d], 0.5);
                    f[d] = xsub(xsub(diff, floor(diff)), 0.5);
                }
                matvec_x(s_cell.mat, f, b);
                const double d2 = norm2_x(b);
                if (d2 < P.cutoff2) energy += pair_energy(row[s_sp[l1]], d2);
            }
        }
    }
    if (valid) P.out[i] = energy;
}

}  // namespace ceg
