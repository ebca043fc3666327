// Small-batch (latency) path of energy_point (sm_100a).
//
// A serial Monte-Carlo chain asks for ONE placement at a time (montecarlo.jl:490-504 + compute_ewald): the batched
// kernels then cost five API calls and two launches for a few microseconds of work (round 1: 69 us per call).  Here a
// call is one kernel launch and nothing else: inputs and outputs live in a pinned host block that is MAPPED into the
// device address space, so the kernel reads the positions and writes the energies over PCIe itself, and the host
// polls the result slots instead of synchronising a stream: every energy travels in one 16-byte store together with
// the call's generation number, so a slot whose generation matches holds its value — no completion counter, no
// system-scope fence between "data" and "flag" on the critical path.
//
// One CTA (16 warps) per placement: VdW grid terms (warp 0, one lane per atom, grids.jl:311-318), Coulomb grid terms
// (warp 2), block test (warp 1) and the reciprocal-space term (ewald.jl:555-577, the k-list spread over the CTA) run side by
// side — power tables by sincospi in shared memory as in reciprocal_kernel, fixed-order reduction.
#pragma once
#include "ewald.cuh"
#include "interp.cuh"

namespace ceg {

constexpr int SMALL_THREADS = 512;
constexpr int SMALL_MAX_ATOMS = 32;       // one lane per atom for the grid terms and the block test

constexpr int SMALL_INLINE_DOUBLES = 160;   // inputs up to this size travel inside the kernel parameters

// The two halves of atom_grid_terms for grids without fused pair records (the latency kernel evaluates them on
// different warps; same operations, same bits)
__device__ __forceinline__ double atom_vdw_term(const EnergyPointsParams& P, int gi, const double* p) {
    if (!P.vdw_present[gi]) return 0.0;
    Stencil st;
    make_stencil(P.vdw[gi].uc, P.vdw[gi].gs, p, st);
    return interp_any(P.vdw[gi], st);
}
__device__ __forceinline__ double atom_coulomb_term(const EnergyPointsParams& P, double qa, const double* p) {
    if (!P.has_coulomb) return 0.0;
    Stencil st;
    make_stencil(P.coulomb.uc, P.coulomb.gs, p, st);
    return qa * interp_any(P.coulomb, st);
}

struct SmallParams {
    EnergyPointsParams E;            // grids, block mask, n_atoms (its atomsidx / charges / xyz pointers are not used)
    // inputs, as doubles: [atomsidx as int32, padded to 8 bytes][charges n_atoms][xyz n_pos x n_atoms x 3] — either in
    // the mapped host block (`in`) or, when small, by value in `inl` (no PCIe read on the critical path)
    const double* in;
    int n_inline;
    int n_idx8;
    double inl[SMALL_INLINE_DOUBLES];
    int has_ewald;
    double invmat[9];
    int kx, ky, kz;
    int64_t n_k;
    const KVecDev* kv;
    const double2* rest;             // null: the framework's S_f inside kv; else the structure factor to interact with (mapped host)
    double energy_net_charges, static_contribution;
    double2* out;                    // mapped host: per placement {vdw, gen}, {coulomb, gen} (gen = bits of the call's number)
    unsigned long long gen;
};

__global__ void __launch_bounds__(SMALL_THREADS) energy_points_small_kernel(const __grid_constant__ SmallParams S) {
    extern __shared__ double2 s_tab[];
    __shared__ double s_v[SMALL_MAX_ATOMS], s_c[SMALL_MAX_ATOMS], s_q[SMALL_MAX_ATOMS];
    __shared__ double s_fa[SMALL_THREADS / 32], s_aa[SMALL_THREADS / 32];
    __shared__ int s_blocked;
    const EnergyPointsParams& P = S.E;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = blockIdx.x;
    const double* in = S.n_inline ? S.inl : S.in;             // __grid_constant__: the parameter block is addressable
    const int* idx = reinterpret_cast<const int*>(in);
    const double* q_in = in + S.n_idx8;
    const double* mol = q_in + P.n_atoms + (int64_t)3 * P.n_atoms * i;
    const bool recip = S.has_ewald && S.n_k > 0;
    if (tid == 0) s_blocked = 0;
    if (tid < P.n_atoms) s_q[tid] = q_in[tid];
    __syncthreads();
    // The three parts of energy_point are independent until the final sum and run side by side: warp 0 interpolates
    // (one lane per atom), warp 1 tests the blocking mask, every warp then fills the Ewald phase tables.
    if ((warp == 0 || warp == 2) && lane < P.n_atoms) {
        const double p[3] = {mol[3 * lane], mol[3 * lane + 1], mol[3 * lane + 2]};
        const int gi = idx[lane];
        if (P.pair_data[gi]) {          // fused {VdW, Coulomb} records: one stencil gives both
            if (warp == 0) atom_grid_terms(P, gi, s_q[lane], p, s_v[lane], s_c[lane]);
        } else if (warp == 0) {         // the two grids of an atom are independent lookups: side by side
            s_v[lane] = atom_vdw_term(P, gi, p);
        } else {
            s_c[lane] = atom_coulomb_term(P, s_q[lane], p);
        }
    }
    if (warp == 1 && lane < P.n_atoms && P.has_block) {
        const double p[3] = {mol[3 * lane], mol[3 * lane + 1], mol[3 * lane + 2]};
        if (atom_blocked(P.block, p)) s_blocked = 1;
    }
    double fa = 0.0, aa = 0.0;
    if (recip) {
        const int nx = S.kx + 1, ny = 2 * S.ky + 1, nz = 2 * S.kz + 1;
        const int per_atom = nx + ny + nz;
        for (int t = tid; t < per_atom * P.n_atoms; t += SMALL_THREADS) {
            const int a = t / per_atom, u = t - a * per_atom;
            const double x = mol[3 * a], y = mol[3 * a + 1], z = mol[3 * a + 2];
            int axis, m;
            if (u < nx) { axis = 0; m = u; }
            else if (u < nx + ny) { axis = 1; m = u - nx - S.ky; }
            else { axis = 2; m = u - nx - ny - S.kz; }
            double f = S.invmat[axis] * x + S.invmat[3 + axis] * y + S.invmat[6 + axis] * z;   // ewald.jl:358
            f -= floor(f);
            double sn, cs;
            sincospi(2.0 * red_prod((double)m, f), &sn, &cs);
            s_tab[t] = make_double2(cs, sn);
        }
        __syncthreads();
        for (int64_t k = tid; k < S.n_k; k += SMALL_THREADS) {
            const KVecDev kv = S.kv[k];
            double sre = 0.0, sim = 0.0;
            for (int a = 0; a < P.n_atoms; a++) {
                const double2* ta = s_tab + a * per_atom;
                const double2 ex = ta[kv.ix], ey = ta[nx + S.ky + kv.jy], ez = ta[nx + ny + S.kz + kv.jz];
                const double yzr = ey.x * ez.x - ey.y * ez.y, yzi = ey.x * ez.y + ey.y * ez.x;
                const double qa = s_q[a];
                sre += qa * (ex.x * yzr - ex.y * yzi);
                sim += qa * (ex.x * yzi + ex.y * yzr);
            }
            const double2 sf = S.rest ? S.rest[k] : make_double2(kv.sf_re, kv.sf_im);
            fa += kv.kf * (sf.x * sre + sf.y * sim);
            aa += kv.kf * (sre * sre + sim * sim);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fa += __shfl_xor_sync(0xffffffffu, fa, o);
            aa += __shfl_xor_sync(0xffffffffu, aa, o);
        }
        if (lane == 0) { s_fa[warp] = fa; s_aa[warp] = aa; }
    }
    __syncthreads();
    if (tid == 0) {
        double vdw = 0.0, coul = 0.0;
        if (s_blocked) {
            vdw = 1e100;     // grids.jl:313
        } else {
            for (int a = 0; a < P.n_atoms; a++) { vdw += s_v[a]; coul += s_c[a]; }
            if (recip) {
                double f = 0.0, q = 0.0;
                for (int w = 0; w < SMALL_THREADS / 32; w++) { f += s_fa[w]; q += s_aa[w]; }
                coul += 2.0 * (f + S.energy_net_charges) + (q + S.static_contribution);   // ewald.jl:572-576
            }
        }
        const double g = __longlong_as_double((long long)S.gen);
        S.out[2 * i] = make_double2(vdw, g);            // one 16-byte store each: value and generation arrive together
        S.out[2 * i + 1] = make_double2(coul, g);
        __threadfence_system();                         // push them out now rather than at kernel exit
    }
}

// The floor of the latency path on a given box: a kernel that only raises the completion flag in mapped host memory
// (ceg_measure_launch_roundtrip: launch -> flag visible to the polling CPU).
__global__ void flag_only_kernel(volatile unsigned long long* flag, unsigned long long gen) {
    *flag = gen;
    __threadfence_system();
}

}  // namespace ceg
