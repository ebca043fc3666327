/*
 * libceg_cuda — C ABI of the B200 (sm_100a) energy-grid hot path.
 *
 * The reference (CrystalEnergyGrids.jl) is pure Julia and has no FFI of its own; the drop-in
 * boundary is its Julia function surface.  Each entry point below replaces the
 * inner loops of the reference function cited next to it; the Julia-side `ccall` stubs a
 * maintainer would add are in INTEGRATION.md and julia/CEGCuda.jl, and the Python host mirror
 * (crystalenergygrids.jl_b200/) binds the same symbols through ctypes.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns 0 on success, non-zero on error, and
 *    ceg_last_error() then returns a thread-local message (the host shim raises it, like the
 *    reference's `error(...)`).
 *  - units are stripped by the caller: Å, K, e.  Every physical constant (GRID_TO_KELVIN,
 *    COULOMBIC_CONVERSION_FACTOR, thresholds) is an ARGUMENT; the device code carries none.
 *  - 3x3 matrices are 9 doubles in column-major order (m[3*c+r] = component r of cell vector c),
 *    i.e. the memory layout of Julia's SMatrix{3,3,Float64,9}.
 *  - host buffers are owned by the caller (Julia Arrays under GC.@preserve); device memory is
 *    owned by the library behind opaque handles.  Entry points are blocking and re-entrant
 *    (one mutex + one stream set per handle), because the reference calls interpolate_grid and
 *    compute_ewald from concurrent tasks (montecarlo.jl:573-578).
 *  - functions suffixed `_dev` take DEVICE pointers (on the handle's device) and a cudaStream_t
 *    passed as void*; they are asynchronous on that stream and exist so that a caller which keeps
 *    trial positions resident in HBM avoids the PCIe round trip.
 *  - there is no CPU fallback anywhere: without a usable CUDA device every call fails.
 */
#ifndef CEG_CUDA_H
#define CEG_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEG_ABI_VERSION 2

/* Periodic cell used for minimum-image distances (utils.jl:146-155, 210-246).
 * `ortho` and `safemin2` are the outputs of prepare_periodic_distance_computations. */
typedef struct ceg_cell_t {
    double mat[9];
    double invmat[9];
    int32_t ortho;
    int32_t _pad;
    double safemin2;
} ceg_cell_t;

/* GridCoordinatesSetup (coordinates.jl:15-41): dims = nodes-1 per axis. */
typedef struct ceg_gridspec_t {
    int32_t dims[3];
    int32_t _pad;
    double size[3];
    double shift[3];
    double delta[3];
} ceg_gridspec_t;

/* One (probe species, framework species) entry of ForceField.interactions as seen by
 * derivativesGrid (interactions.jl:432-472, 599-610).  Terms are summed in the reference's
 * kind order: HardSphere, Lennard-Jones, Buckingham. */
#define CEG_PAIR_HS   1
#define CEG_PAIR_LJ   2
#define CEG_PAIR_BUCK 4
typedef struct ceg_pair_t {
    int32_t flags;     /* OR of CEG_PAIR_*; 0 = NoInteraction; <0 = UndefinedInteraction */
    int32_t _pad;
    double hs_r;       /* R1+R2 (Å): value = Inf where r2 < hs_r^2 */
    double lj_eps, lj_sigma, lj_shift;
    double buck_A, buck_B, buck_C, buck_shift;
} ceg_pair_t;

/* One (species, species) entry of ForceField.interactions as seen by the ENERGY call rule(d^2)
 * (interactions.jl:367-409) between guest atoms: ceg_pair_t's terms plus the real-space Ewald term
 * CoulombEwaldDirect (raspa.jl:688-700) and bare Coulomb.  qq = COULOMBIC_CONVERSION_FACTOR * q1 * q2. */
#define CEG_PAIR_EWALD   8
#define CEG_PAIR_COULOMB 16
typedef struct ceg_pair_energy_t {
    int32_t flags;
    int32_t _pad;
    double hs_r;
    double lj_eps, lj_sigma, lj_shift;
    double buck_A, buck_B, buck_C, buck_shift;
    double ewald_alpha, ewald_qq;
    double coulomb_qq;
} ceg_pair_energy_t;

typedef struct ceg_probe_s* ceg_probe_h;   /* ProbeSystem(s)            probes.jl:12-63   */
typedef struct ceg_grid_s*  ceg_grid_h;    /* device-resident EnergyGrid grids.jl:10-16   */
typedef struct ceg_ewald_s* ceg_ewald_h;   /* EwaldFramework            ewald.jl:40-50    */
typedef struct ceg_block_s* ceg_block_h;   /* BlockFile                 coordinates.jl:83 */
typedef struct ceg_guest_s* ceg_guest_h;   /* guests of a SimulationStep energy.jl:43-60  */
typedef struct ceg_gridpair_s* ceg_gridpair_h; /* {VdW grid of a species, Coulomb grid} fused on the device */

const char* ceg_last_error(void);
int ceg_abi_version(void);
int ceg_device_count(int32_t* n);
/* name (<=255 chars), SM count, memory bytes, SM clock kHz of a device */
int ceg_device_info(int32_t device, char* name256, int32_t* sm_count, int64_t* mem_bytes, int32_t* clock_khz);

/* Large device scratch buffers are cached per process between calls/handles (cudaMalloc/cudaFree cost
 * up to hundreds of ms); this releases the cache back to the driver. */
int ceg_trim_device_cache(void);

/* Number of kernels this library has launched in this process (all handles, all threads). */
int ceg_kernel_launch_count(int64_t* n);
/* FP64 FMA-chain microbenchmark: sustained DFMA throughput of the device in TFLOP/s (2 flop per FMA).
 * The roofline denominator of the construction kernels (MEASURED_PEAKS.json has no FP64 entry). */
int ceg_measure_fp64_peak(int32_t device, double* tflops);
/* Same chains with three per-thread register operands per DFMA (what real kernels issue): on B200 the
 * register-file read bandwidth caps this form ~30 % below the uniform-operand figure. */
int ceg_measure_fp64_peak_regops(int32_t device, double* tflops);
/* Median round trip (µs) of one EMPTY kernel launch through the single-placement latency path's machinery (launch ->
 * completion flag in mapped host memory seen by the polling CPU): the box's floor for ceg_energy_points with one
 * placement; bench.py reports it next to the measured latencies. */
int ceg_measure_launch_roundtrip(int32_t device, int32_t reps, double* median_us);

/* Registers per thread of the main kernels as loaded on `device`: which = 0 fused one-probe + Coulomb build kernel,
 * 1 interpolation (brick layout), 2 fused VdW + Coulomb pair interpolation, 3 reciprocal rows kernel (one atom). */
int ceg_kernel_registers(int32_t device, int32_t which, int32_t* regs);
/* Plain device buffers for chaining the _dev entry points from a host language without a CUDA runtime binding
 * (build_dev -> gather -> upload_dev); ceg_device_download copies one back (e.g. to write its `.grid` file). */
int ceg_device_malloc(int32_t device, int64_t bytes, void** out);
int ceg_device_free(void* ptr);
int ceg_device_download(int32_t device, const void* d_src, int64_t bytes, void* host_dst);
int ceg_device_synchronize(int32_t device);

/* Self-test hook for the branch-free device math of the construction kernels (fastmath.cuh):
 * out[6*i..] = exp_fast(-x[i]), erfcx_fast(x[i]), rsqrt_fast(x[i]), rcp_fast(x[i]), exp_acc(-x[i]), exp_fast(x[i]), x[i] > 0. */
int ceg_selftest_math(int32_t device, int64_t n, const double* x, double* out);

/* Pin / unpin a caller-owned host buffer so that slab write-back and query upload overlap compute. */
int ceg_host_register(void* ptr, int64_t bytes);
int ceg_host_unregister(void* ptr);

/* ------------------------------------------------------------------------------------------
 * Grid construction.  Replaces ProbeSystem (probes.jl:21-63), compute_derivatives_vdw /
 * compute_derivatives_ewald (probes.jl:71-117), derivativesGrid (interactions.jl:432-472),
 * derivatives_ewald (ewald.jl:299-312), _set_gridpoint! and the @threads loops of
 * create_grid_vdw / create_grid_coulomb (grids.jl:118-185).
 *
 * One handle fuses n_probes VdW probes (table row p) and, if want_coulomb, the Ewald real-space
 * grid: n_grids = n_probes + want_coulomb, Coulomb last.  xyz/type/charge describe the
 * REPLICATED supercell (N_sc atoms) and `supercell` is its cell.
 * ------------------------------------------------------------------------------------------ */
int ceg_probe_create(int32_t device, const ceg_cell_t* supercell,
                     int64_t n_atoms, const double* xyz /* n_atoms x 3 */, const int32_t* type /* n_atoms */,
                     const double* charge /* n_atoms, may be NULL if !want_coulomb */,
                     int32_t n_types, int32_t n_probes, const ceg_pair_t* table /* n_probes x n_types */,
                     double cutoff2, int32_t want_coulomb, double alpha,
                     ceg_probe_h* out);
int ceg_probe_free(ceg_probe_h h);
int ceg_probe_num_grids(ceg_probe_h h, int32_t* n_grids);

/* (value, d1[3], d2[3], d3) at arbitrary positions, before _set_gridpoint!:
 * out[(g*n_pos + i)*8 + c].  Literal per-pair reference algorithm (A3) on the device. */
int ceg_compute_derivatives(ceg_probe_h h, int64_t n_pos, const double* xyz /* n_pos x 3 */, double* out);

/* Build the x-slab [i0, i1) of every grid of the handle (i in 0..dims[0], the reference's
 * @threads axis, grids.jl:144,171).  out_f32[g] points at the FULL host array of grid g in the
 * reference/.grid order [8][nx][ny][nz] (z fastest); only planes i0..i1-1 are written.
 * lambda/thr per grid kind as in grids.jl:142-143,169-170.
 * raw_f64 (optional, may be NULL): pre-quantisation derivatives of the slab,
 * raw_f64[((g*8 + c)*(i1-i0) + (i-i0))*ny*nz + j*nz + k] — used by the parity tests. */
int ceg_grid_build(ceg_probe_h h, const ceg_gridspec_t* gs, int32_t i0, int32_t i1,
                   double lambda_vdw, double thr_vdw, double lambda_coul, double thr_coul,
                   float* const* out_f32 /* n_grids host pointers, entries may be NULL */, double* raw_f64);

/* Same, leaving the slab on the device: d_out[g] = device buffer [8][i1-i0][ny][nz] floats.
 * Asynchronous on `stream`.  Returns the kernel's algorithmic pair counts through
 * ceg_probe_last_stats. */
int ceg_grid_build_dev(ceg_probe_h h, const ceg_gridspec_t* gs, int32_t i0, int32_t i1,
                       double lambda_vdw, double thr_vdw, double lambda_coul, double thr_coul,
                       float* const* d_out, void* stream);

/* Work counters of the last build on this handle (after a sync): number of grid points,
 * number of (point, atom image) pairs tested, number inside the cutoff. */
int ceg_probe_last_stats(ceg_probe_h h, int64_t* n_points, int64_t* pairs_tested, int64_t* pairs_in_cutoff);

/* Phases of the last ceg_grid_build on this handle, in ms: out4 = { host time spent building image lists, device span
 * first-to-last kernel, device tail between the last kernel and the last slab copy, host time of the whole call }. */
int ceg_probe_last_timing(ceg_probe_h h, double* out4);

/* Gather the x-slabs of one grid that were built on several devices of this process into a full grid on dst_device
 * (peer copies over NVLink) — the device-resident form of the hand-off the reference makes through `.grid` files
 * (raspa.jl:430-438).  Part p: planes [i0[p], i1[p]) as [8][i1-i0][ny][nz] floats on devices[p] (ceg_grid_build_dev
 * output); d_full: [8][nx][ny][nz] floats on dst_device, ready for ceg_grid_upload_dev.  Blocking. */
int ceg_grid_gather(int32_t n_parts, const int32_t* devices, const float* const* d_slabs, const int32_t* i0,
                    const int32_t* i1, const ceg_gridspec_t* gs, int32_t dst_device, float* d_full);

/* ------------------------------------------------------------------------------------------
 * Interpolation.  Replaces interpolate_grid (grids.jl:212-273) + offsetpoint
 * (coordinates.jl:58-66).  `grid` is the host array AFTER parse_grid (values already multiplied
 * by GRID_TO_KELVIN, grids.jl:78), reference order [8][nx][ny][nz]; it is re-laid out on the
 * device as node-major records.  is_vdw selects the >5e6 K corner test returning 1e100
 * (grids.jl:245-248).
 * ------------------------------------------------------------------------------------------ */
int ceg_grid_upload(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs,
                    const float* grid, int32_t is_vdw, ceg_grid_h* out);
/* The same from a grid that is already on `device` (e.g. the output of ceg_grid_build_dev, or slabs gathered from
 * other GPUs): d_grid is [8][nx][ny][nz] f32; every value is converted as Float32(value * scale) with the product in
 * double — scale = GRID_TO_KELVIN reproduces parse_grid (grids.jl:78) on a grid still in `.grid` units, 1 leaves a
 * grid already in K untouched.  The source may be freed once the call returns. */
int ceg_grid_upload_dev(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, const float* d_grid,
                        double scale, int32_t is_vdw, ceg_grid_h* out);
/* The same from the x-slabs of a sharded build as they sit on `device` after an all-gather (NCCL, MPI, peer copies):
 * part r holds planes [i0[r], i1[r]) in a [8][stride[r]][ny][nz] f32 buffer (stride = NULL: stride[r] = i1[r] - i0[r]);
 * the parts (at most 64) tile 0..nx in order.  The full grid is never assembled: one pass reads the parts and writes the records. */
int ceg_grid_upload_parts_dev(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, int32_t n_parts,
                              const float* const* d_parts, const int32_t* i0, const int32_t* i1, const int64_t* stride,
                              double scale, int32_t is_vdw, ceg_grid_h* out);
/* ---- build with the hand-off fused in (replaces the `.grid` file exchange of raspa.jl:430-438 between the
 * construction and the interpolation stage; single- and multi-GPU) ----
 * ceg_grid_create_empty: an interpolation grid on `device` whose node records are still to be written (layout: the
 * one ceg_grid_upload would choose, NODE or BLOCK); ceg_grid_records: its record storage (device pointer, bytes,
 * layout).  ceg_grid_build_records runs the build kernel of planes [i0, i1) and writes every node straight to its
 * place in the record storage of n_dest grids, converted as Float32(value * scale[g]) (parse_grid, grids.jl:78; one
 * scale per grid of the pass: probes in order, Coulomb last): d_records[t * n_grids + g] = records of grid g at
 * destination t — a handle on the probe's device, on another device of this process, or a ceg_ipc_open mapping of
 * another process's handle (peer stores over NVLink).  A build sharded over the GPUs of a node thus leaves a complete
 * replica on every GPU when the kernels of all ranks have finished (synchronise the ranks before interpolating):
 * no all-gather, no re-layout pass, no `.grid`-order intermediate.  d_out (may be NULL) additionally receives the
 * `.grid`-order slab [8][i1-i0][ny][nz] per grid, for callers that also write the file.  Asynchronous on `stream`
 * like ceg_grid_build_dev.  Results are bit-identical to build -> `.grid` -> parse_grid -> ceg_grid_upload. */
int ceg_grid_create_empty(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, int32_t is_vdw,
                          ceg_grid_h* out);
int ceg_grid_records(ceg_grid_h grid, void** d_records, int64_t* bytes, int32_t* layout);
int ceg_grid_build_records(ceg_probe_h probe, const ceg_gridspec_t* gs, int32_t i0, int32_t i1, double lambda_vdw,
                           double thr_vdw, double lambda_coul, double thr_coul, int32_t n_dest,
                           void* const* d_records, int32_t layout, const double* scale, float* const* d_out,
                           void* stream);
/* Device buffers across processes (CUDA IPC, one process per GPU): the owner exports a 64-byte handle + the offset
 * of d_ptr inside its allocation; a peer process opens it on its own device (peer access enabled) and closes it
 * before the owner frees the buffer. */
int ceg_ipc_export(int32_t device, const void* d_ptr, uint8_t* handle64, int64_t* offset);
int ceg_ipc_open(int32_t device, const uint8_t* handle64, int64_t offset, void** d_ptr);
int ceg_ipc_close(int32_t device, void* d_ptr);
int ceg_grid_free(ceg_grid_h h);
/* Device record layout used by subsequent ceg_grid_upload calls: 0 = node records (1x memory),
 * 1 = (y,z) quads (4x), 2 = whole 8-corner cells (8x, one aligned 256 B block per query),
 * 3 = z pairs (2x), 4 = 2x2x2-node bricks of 256 B (1x memory),
 * -1 = default (bricks for grids between 96 MB and 1.5 GB, node records otherwise: the fastest measured on B200). */
int ceg_set_interp_layout(int32_t layout);
int ceg_grid_layout(ceg_grid_h h, int32_t* layout);
int ceg_interpolate(ceg_grid_h h, int64_t n, const double* xyz /* n x 3 Cartesian Å */, double* out /* n, K */);
int ceg_interpolate_dev(ceg_grid_h h, int64_t n, const double* d_xyz, double* d_out, void* stream);
/* Several grids (e.g. the VdW grid of an atom and the Coulomb grid, grids.jl:316-318) at the same
 * positions: positions are uploaded once; outs[g] receives n energies of grids[g]. */
int ceg_interpolate_multi(int32_t n_grids, const ceg_grid_h* grids, int64_t n, const double* xyz, double* const* outs);
/* Fused pair: energy_point interpolates, per atom, its species' VdW grid and the Coulomb grid at the same position
 * (grids.jl:316-318).  ceg_grid_pair_create interleaves two resident grids that share one GridCoordinatesSetup into
 * 64-byte node records {VdW, Coulomb} (2x the memory of one grid, in addition to the two handles, which stay valid and
 * must outlive the pair); a lookup then fetches only fully used 64-byte units and shares the position, offsetpoint and
 * stencil between the two interpolations.  Results are bit-identical to ceg_interpolate on each grid.  While the pair
 * exists, ceg_energy_points / ceg_energy_lattice use it whenever they are given this (VdW, Coulomb) couple of handles. */
int ceg_grid_pair_create(ceg_grid_h vdw, ceg_grid_h coulomb, ceg_gridpair_h* out);
int ceg_grid_pair_free(ceg_gridpair_h h);
int ceg_interpolate_pair(ceg_gridpair_h h, int64_t n, const double* xyz, double* out_vdw, double* out_coul);
int ceg_interpolate_pair_dev(ceg_gridpair_h h, int64_t n, const double* d_xyz, double* d_out_vdw, double* d_out_coul,
                             void* stream);

/* Interpolation sweep over positions generated on the device (the 10^9-query sweep of C5): query
 * first + t (t < n) sits at mat * (u0,u1,u2), u_d = splitmix64(seed, 3 (first + t) + d) in [0,1), mat = 9 doubles
 * column-major.  Returns the sum and the minimum of the unblocked energies and the number of blocked (1e100)
 * queries, reduced on the device in a fixed order; keep (may be NULL) receives the first n_keep energies;
 * kernel_ms (may be NULL) the device time of the sweep.  Replaces a loop of interpolate_grid (grids.jl:212-273)
 * over random trial positions. */
int ceg_interpolate_sweep(ceg_grid_h h, int64_t first, int64_t n, uint64_t seed, const double* mat,
                          double* sum, double* min, int64_t* n_blocked, int64_t n_keep, double* keep, float* kernel_ms);

/* ------------------------------------------------------------------------------------------
 * Ewald reciprocal space.  ceg_ewald_create replaces setup_Eik + ewald_main_loop! for the
 * framework (ewald.jl:109-185, 264-271): S_f(k) = sum_sites q exp(2 pi i (ix fx + jy fy + jz fz))
 * over the n_sites supercell sites given in SUPERCELL-fractional coordinates.
 * kvec[3*m..] = (ix, jy, jz) of k-vector m in kindices order (ewald.jl:213-236); kfactor as
 * ewald.jl:247-259.  S_f_reim (n_k x 2) is returned to the caller (StoreRigidChargeFramework).
 * ceg_ewald_reciprocal replaces move_one_system! + compute_ewald(ctx) (ewald.jl:350-364,555-577)
 * for n_pos independent placements of one rigid set of n_atoms charges:
 *   out = 2*(sum_k kf Re(S_f conj S_a) + energy_net_charges) + (sum_k kf |S_a|^2 + static_contribution)
 * ------------------------------------------------------------------------------------------ */
int ceg_ewald_create(int32_t device, const double* invmat /* supercell inverse, 9 doubles */,
                     const int32_t kmax[3], int64_t n_k, const int32_t* kvec, const double* kfactor,
                     int64_t n_sites, const double* frac /* n_sites x 3 */, const double* charge,
                     ceg_ewald_h* out, double* S_f_reim);
/* The same handle from a structure factor computed elsewhere (the StoreRigidChargeFramework of an unchanged
 * initialize_ewald): S_f_reim (n_k x 2) is taken as given. */
int ceg_ewald_create_from_sf(int32_t device, const double* invmat, const int32_t kmax[3], int64_t n_k, const int32_t* kvec,
                             const double* kfactor, const double* S_f_reim, ceg_ewald_h* out);
int ceg_ewald_free(ceg_ewald_h h);
int ceg_ewald_reciprocal(ceg_ewald_h h, int64_t n_pos, int32_t n_atoms,
                         const double* xyz /* n_pos x n_atoms x 3 Cartesian Å */, const double* q /* n_atoms */,
                         double energy_net_charges, double static_contribution, double* out /* n_pos, K */);
/* single_contribution_ewald (ewald.jl:704-738) for n_pos trial placements of one molecule — the Ewald half of a
 * batched Monte-Carlo trial evaluation (montecarlo.jl:563-606).  rest_reim (n_k x 2, host) = framework + all
 * other adsorbates: StoreRigidChargeFramework + sums[:,1] - sums[:,ij+1].
 *   out = 2 sum_k kf Re(rest conj S_single) + sum_k kf |S_single|^2 */
int ceg_ewald_single_contribution(ceg_ewald_h h, const double* rest_reim, int64_t n_pos, int32_t n_atoms,
                                  const double* xyz, const double* q, double* out);
/* (asynchronous on `stream`; the charges of molecules of up to 32 atoms travel by value with the launch, so calls on
 * different streams never share state; larger molecules use a handle-owned buffer and must then share one stream) */
int ceg_ewald_reciprocal_dev(ceg_ewald_h h, int64_t n_pos, int32_t n_atoms, const double* d_xyz, const double* q_host,
                             double energy_net_charges, double static_contribution, double* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Guest-guest pair energies: single_contribution_vdw_noneighbour (energy.jl:407-427) for n_trials
 * placements of one molecule (n_atoms atoms of species mol_species) against the n_env atoms of the other
 * guests; atoms whose molecule id equals skip_mol (the molecule being moved) are ignored.  `cell` is the
 * simulation cell (distances by unsafe_periodic_distance2!, utils.jl:294-302: wrap only).
 * out[t] = sum over pairs with d2 < cutoff2 of rule(d2).  Together with ceg_energy_points (framework
 * grids) and ceg_ewald_single_contribution (reciprocal part) this evaluates movement_energy
 * (montecarlo.jl:563-606) for a batch of trial moves.
 * ------------------------------------------------------------------------------------------ */
/* The same with the environment resident on the device: create once per simulation, evaluate batches of trial
 * placements (ceg_guest_env_energy = the arguments of ceg_guest_energy after the environment), move atoms
 * [first, first + n) after an accepted step (ceg_guest_env_update). */
int ceg_guest_env_create(int32_t device, const ceg_cell_t* cell, double cutoff2, int32_t n_species,
                         const ceg_pair_energy_t* table, int64_t n_env, const double* env_xyz,
                         const int32_t* env_species, const int32_t* env_mol, ceg_guest_h* out);
int ceg_guest_env_free(ceg_guest_h h);
int ceg_guest_env_update(ceg_guest_h h, int64_t first, int64_t n, const double* xyz /* n x 3 */);
int ceg_guest_env_energy(ceg_guest_h h, int64_t n_trials, int32_t n_atoms, const int32_t* mol_species, int32_t skip_mol,
                         const double* trial_xyz, double* out);
/* Device time (ms, CUDA events) of the kernel launched by the last ceg_guest_env_energy call on this handle. */
int ceg_guest_env_last_kernel_ms(ceg_guest_h h, double* ms);
int ceg_guest_energy(int32_t device, const ceg_cell_t* cell, double cutoff2, int32_t n_species,
                     const ceg_pair_energy_t* table /* n_species x n_species */, int64_t n_env, const double* env_xyz,
                     const int32_t* env_species, const int32_t* env_mol, int64_t n_trials, int32_t n_atoms,
                     const int32_t* mol_species, int32_t skip_mol, const double* trial_xyz, double* out);

/* ------------------------------------------------------------------------------------------
 * Blocking mask (coordinates.jl:83-101) and fused batched evaluation: energy_point over n_pos
 * placements (grids.jl:311-327), the inner loop of energy_grid (grids.jl:394-419).
 * mask: nx*ny*nz bytes, index (a*ny + b)*nz + c with (a,b,c) = round(offsetpoint)-1.
 * For each placement: blocked -> (1e100, 0); else vdw = sum_i interp(vdw[atomsidx[i]], pos_i),
 * coulomb = sum_i q_i interp(coulomb, pos_i) + reciprocal.
 * vdw[g] may be NULL (zero grid, ewald_precision == -Inf); coulomb/ewald/block may be NULL.
 * ------------------------------------------------------------------------------------------ */
int ceg_block_upload(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs,
                     const uint8_t* mask, ceg_block_h* out);
int ceg_block_free(ceg_block_h h);
/* parse_blockfile (coordinates.jl:112-167): mask[(a*ny+b)*nz+c] = 1 if node (a,b,c) lies within radius of a
 * blocking sphere under the reference's periodic distance.  centers: n x 3 Cartesian Å, already passed
 * through inverse_offsetpoint(offsetpoint(.)) by the caller (coordinates.jl:131-134); radius2 = radius^2. */
int ceg_block_from_spheres(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, int32_t n_spheres,
                           const double* centers, const double* radius2, uint8_t* mask);
/* BlockFile(g::EnergyGrid) (grids.jl:188-204): voxel-dilated mask of the nodes whose VdW value exceeds
 * `threshold` (5e6 K); values = channel 0 of the grid, [nx][ny][nz]. */
int ceg_block_from_values(int32_t device, const ceg_gridspec_t* gs, const float* values, float threshold, uint8_t* mask);
/* Latency: ceg_energy_points, ceg_interpolate, ceg_ewald_reciprocal and ceg_ewald_single_contribution calls with at most 256 placements
 * of at most 32 atoms (single Monte-Carlo steps, montecarlo.jl:490-504) are ONE kernel launch with no copies: small
 * inputs travel inside the launch parameters, results come back through mapped pinned memory as {value, generation}
 * slots the host polls (csrc/latency.cuh); larger batches take the throughput kernels.  Same results within 1e-9. */
int ceg_energy_points(int32_t n_vdw, const ceg_grid_h* vdw, ceg_grid_h coulomb, ceg_ewald_h ewald, ceg_block_h block,
                      int64_t n_pos, int32_t n_atoms, const int32_t* atomsidx /* n_atoms, 0-based into vdw[] */,
                      const double* charges /* n_atoms */, const double* xyz /* n_pos x n_atoms x 3 */,
                      double energy_net_charges, double static_contribution,
                      double* out_vdw /* n_pos */, double* out_coulomb /* n_pos */);
/* energy_grid (grids.jl:346-424) with the lattice generated on the device: nums = (numA, numB, numC), steps[3 d + c] =
 * component c of axis_d / num_d, rotpos = n_rot x n_atoms x 3 rotated molecule offsets.  Point (iA, iB, iC) of rotation
 * r: atoms at (iA stepA + iB stepB) + iC stepC + rotpos[r][a] (grids.jl:396).  out[r][iA][iB][iC] = vdw + coulomb of
 * energy_point at that placement (1e100 when blocked).  Other arguments as ceg_energy_points. */
int ceg_energy_lattice(int32_t n_vdw, const ceg_grid_h* vdw, ceg_grid_h coulomb, ceg_ewald_h ewald, ceg_block_h block,
                       const int32_t* nums, const double* steps, int32_t n_rot, int32_t n_atoms, const double* rotpos,
                       const int32_t* atomsidx, const double* charges, double energy_net_charges,
                       double static_contribution, double* out /* n_rot x numA x numB x numC */);

#ifdef __cplusplus
}
#endif
#endif /* CEG_CUDA_H */
