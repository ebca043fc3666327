/*
 * ORACLE — test infrastructure, not product code.
 *
 * A plain-C, FP64, CPU restatement of the reference's hot-path algorithm (CrystalEnergyGrids.jl,
 * pure Julia, which cannot run in this environment: no Julia, no network).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (libceg_cuda.so) never does.
 *
 * Parity pin: oracle/pin_goldens.py + tests/test_oracle_goldens.py check it against the
 * known-answer energies of the reference's own test-suite (test/runtests.jl:24-56,165-169,
 * 269-272); the generated inputs live in tests/golden/.
 *
 * Compile WITHOUT floating-point contraction (-ffp-contract=off): every expression below keeps
 * the reference's operation order.  Each function cites the reference file:line it follows
 * (paths relative to /root/reference/src).
 *
 * Third-party arithmetic the reference delegates (versions unpinned by the reference):
 * SpecialFunctions.erfc/erf -> libm erfc/erf here; Base.cispi -> sincos(2*pi*x) on the fractional
 * part; Base.exp/sqrt -> libm.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ceg_cuda.h"

#define M(m, r, c) ((m)[3 * (c) + (r)]) /* column-major 3x3 */

static inline void matvec(const double* m, const double* v, double* out) {
    /* mul!(out, m, v): row r = m[r,1]*v1 + m[r,2]*v2 + m[r,3]*v3 */
    for (int r = 0; r < 3; r++) out[r] = M(m, r, 0) * v[0] + M(m, r, 1) * v[1] + M(m, r, 2) * v[2];
}
static inline double norm2(const double* v) { /* utils.jl:199-205 */
    double r2 = 0.0;
    r2 += v[0] * v[0];
    r2 += v[1] * v[1];
    r2 += v[2] * v[2];
    return r2;
}

/* utils.jl:210-246 periodic_distance2_fromcartesian! / periodic_distance2!.
 * `buffer` holds the Cartesian difference on entry and the image vector the reference leaves
 * in it on exit (including the stale last-tried image when no closer image is found). */
double ceg_o_periodic_distance2(double* buffer, const double* mat, const double* invmat, int ortho, double safemin2) {
    double f[3];
    matvec(invmat, buffer, f);
    for (int i = 0; i < 3; i++) {
        double diff = f[i] + 0.5;
        f[i] = diff - floor(diff) - 0.5;
    }
    matvec(mat, f, buffer);
    double ref2 = norm2(buffer);
    if (ortho || ref2 <= safemin2) return ref2;
    for (int i = 0; i < 3; i++) {
        f[i] += 1;
        matvec(mat, f, buffer);
        double n2 = norm2(buffer);
        if (n2 < ref2) return n2;
        f[i] -= 2;
        matvec(mat, f, buffer);
        n2 = norm2(buffer);
        if (n2 < ref2) return n2;
        f[i] += 1;
    }
    return ref2;
}

/* interactions.jl:432-472 derivativesGrid(rule, d2) for each kind, summed as
 * derivativesGrid(::InteractionRuleSum) does (interactions.jl:599-610) in kind order
 * HardSphere < LennardJones < Buckingham.  out = (U - shift, DU, D2U, D3U). */
void ceg_o_derivatives_pair(const ceg_pair_t* p, double r2, double* out) {
    double value = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    int nterms = 0;
    if (p->flags & CEG_PAIR_HS) {
        double s = p->hs_r;
        double v = (r2 < s * s) ? INFINITY : 0.0;
        value += v - 0.0;
        nterms++;
    }
    if (p->flags & CEG_PAIR_LJ) {
        double eps = p->lj_eps, sig = p->lj_sigma;
        double t = (sig * sig) / r2;
        double x6 = t * t * t;
        double r4 = r2 * r2;
        double v = 4 * eps * x6 * (x6 - 1);
        double p1 = 24 * eps * (x6 * (1 - 2 * x6)) / r2;
        double p2 = 96 * eps * (x6 * (7 * x6 - 2)) / r4;
        double p3 = 384 * eps * (x6 * (5 - 28 * x6)) / (r4 * r4);
        value += v - p->lj_shift;
        d1 += p1; d2 += p2; d3 += p3;
        nterms++;
    }
    if (p->flags & CEG_PAIR_BUCK) {
        double A = p->buck_A, B = p->buck_B, C = p->buck_C;
        double r4 = r2 * r2;
        double r = sqrt(r2);
        double r6 = r4 * r2;
        double x6 = C / r6;
        double xe = A * exp(-B * r);
        double v = xe - x6;
        double p1 = -B * xe / r + 6 * x6 / r2;
        double p2 = -48 * x6 / r4 + B * xe * (1 + B * r) / (r2 * r);
        double p3 = -(3 * B * r + B * B * r2 + 3) * B * xe * r / r6 + 480 * C / (r6 * r6);
        value += v - p->buck_shift;
        d1 += p1; d2 += p2; d3 += p3;
        nterms++;
    }
    (void)nterms;
    out[0] = value; out[1] = d1; out[2] = d2; out[3] = d3;
}

/* ewald.jl:299-312 derivatives_ewald */
void ceg_o_derivatives_ewald(double alpha, double charge, double r2, double* out) {
    double r = sqrt(r2);
    double r3 = r2 * r;
    double r5 = r3 * r2;
    double r2a2 = r2 * (alpha * alpha);
    double er2a2 = 2 * alpha * r * exp(-r2a2) / sqrt(M_PI);
    double erfar = erfc(alpha * r);
    out[0] = charge * erfar / r;
    out[1] = -charge * (er2a2 + erfar) / r3;
    out[2] = charge * (er2a2 * (3 + 2 * r2a2) + 3 * erfar) / r5;
    out[3] = charge * (-er2a2 * (15 + 10 * r2a2 + 4 * (r2a2 * r2a2)) - 15 * erfar) / (r5 * r2);
}

static inline void accumulate(double* acc, const double* d, const double* b) {
    /* probes.jl:85-89 / 110-114 */
    acc[0] += d[0];
    acc[1] += d[1] * b[0];
    acc[2] += d[1] * b[1];
    acc[3] += d[1] * b[2];
    double d13 = b[0] * b[2];
    acc[4] += d[2] * (b[0] * b[1]);
    acc[5] += d[2] * d13;
    acc[6] += d[2] * (b[1] * b[2]);
    acc[7] += d[3] * d13 * b[1];
}

/* probes.jl:71-92 compute_derivatives_vdw: `row` = pair-table row of the probe (n_types). */
void ceg_o_compute_derivatives_vdw(const ceg_cell_t* cell, int64_t n_atoms, const double* xyz, const int32_t* type,
                                   const ceg_pair_t* row, double cutoff2, const double* pos, double* out) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < n_atoms; i++) {
        double b[3] = {pos[0] - xyz[3 * i], pos[1] - xyz[3 * i + 1], pos[2] - xyz[3 * i + 2]};
        double d2 = ceg_o_periodic_distance2(b, cell->mat, cell->invmat, cell->ortho, cell->safemin2);
        if (d2 >= cutoff2) continue;
        const ceg_pair_t* p = &row[type[i]];
        if (p->flags == 0) continue; /* NoInteraction: adds exact zeros */
        double d[4];
        ceg_o_derivatives_pair(p, d2, d);
        accumulate(acc, d, b);
    }
    memcpy(out, acc, sizeof(acc));
}

/* probes.jl:94-117 compute_derivatives_ewald */
void ceg_o_compute_derivatives_ewald(const ceg_cell_t* cell, int64_t n_atoms, const double* xyz, const double* charge,
                                     double alpha, double cutoff2, const double* pos, double* out) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double smallest = INFINITY;
    for (int64_t i = 0; i < n_atoms; i++) {
        double b[3] = {pos[0] - xyz[3 * i], pos[1] - xyz[3 * i + 1], pos[2] - xyz[3 * i + 2]};
        double d2 = ceg_o_periodic_distance2(b, cell->mat, cell->invmat, cell->ortho, cell->safemin2);
        if (d2 >= cutoff2) continue;
        smallest = fmin(smallest, d2);
        double d[4];
        ceg_o_derivatives_ewald(alpha, charge[i], d2, d);
        accumulate(acc, d, b);
    }
    if (smallest < 1.0) acc[0] = INFINITY;
    memcpy(out, acc, sizeof(acc));
}

/* Batched front-end: out[(g*n_pos + i)*8 + c], grids = n_probes VdW rows then Coulomb. */
void ceg_o_compute_derivatives(const ceg_cell_t* cell, int64_t n_atoms, const double* xyz, const int32_t* type,
                               const double* charge, int32_t n_types, int32_t n_probes, const ceg_pair_t* table,
                               double cutoff2, int32_t want_coulomb, double alpha, int64_t n_pos, const double* pos,
                               double* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_pos; i++) {
        for (int g = 0; g < n_probes; g++)
            ceg_o_compute_derivatives_vdw(cell, n_atoms, xyz, type, table + (size_t)g * n_types, cutoff2, pos + 3 * i,
                                          out + ((size_t)g * n_pos + i) * 8);
        if (want_coulomb)
            ceg_o_compute_derivatives_ewald(cell, n_atoms, xyz, charge, alpha, cutoff2, pos + 3 * i,
                                            out + ((size_t)n_probes * n_pos + i) * 8);
    }
}

static inline double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

/* grids.jl:118-135 _set_gridpoint!: der = (value, d1[3], d2[3], d3) -> 8 floats */
void ceg_o_set_gridpoint(const double* delta, double lambda, double thr, const double* der, float* out8) {
    double v = der[0], a1 = der[1], a2 = der[2], a3 = der[3], b1 = der[4], b2 = der[5], b3 = der[6], c = der[7];
    if (v > thr) {
        v = 2 * thr;
        a1 = clampd(a1, -thr, thr); a2 = clampd(a2, -thr, thr); a3 = clampd(a3, -thr, thr);
        b1 = b2 = b3 = 0.0;
        c = 0.0;
    }
    out8[0] = (float)(v * lambda);
    out8[1] = (float)(a1 * delta[0] * lambda);
    out8[2] = (float)(a2 * delta[1] * lambda);
    out8[3] = (float)(a3 * delta[2] * lambda);
    out8[4] = (float)(b1 * (delta[0] * delta[1]) * lambda);
    out8[5] = (float)(b2 * (delta[0] * delta[2]) * lambda);
    out8[6] = (float)(b3 * (delta[1] * delta[2]) * lambda);
    out8[7] = (float)(c * (delta[0] * delta[1] * delta[2]) * lambda);
}

/* grids.jl:137-185 create_grid_vdw / create_grid_coulomb loops (coordinates.jl:72-76 abc_to_xyz),
 * for x planes i0, i0+istride, ... < i1.  Outputs use the reference order [8][nx][ny][nz];
 * raw (optional) = [g][8][n_planes][ny][nz] doubles before _set_gridpoint!.
 * The reference makes one pass per grid; the sum over atoms is restated per grid exactly. */
void ceg_o_grid_build(const ceg_cell_t* cell, int64_t n_atoms, const double* xyz, const int32_t* type,
                      const double* charge, int32_t n_types, int32_t n_probes, const ceg_pair_t* table, double cutoff2,
                      int32_t want_coulomb, double alpha, const ceg_gridspec_t* gs, int32_t i0, int32_t i1,
                      int32_t istride, double lambda_vdw, double thr_vdw, double lambda_coul, double thr_coul,
                      float* const* out_f32, double* raw) {
    const int64_t nx = gs->dims[0] + 1, ny = gs->dims[1] + 1, nz = gs->dims[2] + 1;
    const int64_t npts = nx * ny * nz;
    const int n_grids = n_probes + (want_coulomb ? 1 : 0);
    const int64_t n_planes = (i1 - i0 + istride - 1) / istride;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int64_t ip = 0; ip < n_planes; ip++) {
        for (int64_t j = 0; j < ny; j++) {
            int64_t i = i0 + ip * istride;
            for (int64_t k = 0; k < nz; k++) {
                double pos[3] = {i * gs->size[0] / gs->dims[0] + gs->shift[0],
                                 j * gs->size[1] / gs->dims[1] + gs->shift[1],
                                 k * gs->size[2] / gs->dims[2] + gs->shift[2]};
                for (int g = 0; g < n_grids; g++) {
                    double der[8];
                    float q8[8];
                    int coul = (g == n_probes);
                    if (coul)
                        ceg_o_compute_derivatives_ewald(cell, n_atoms, xyz, charge, alpha, cutoff2, pos, der);
                    else
                        ceg_o_compute_derivatives_vdw(cell, n_atoms, xyz, type, table + (size_t)g * n_types, cutoff2,
                                                      pos, der);
                    if (raw)
                        for (int c = 0; c < 8; c++)
                            raw[(((size_t)g * 8 + c) * n_planes + ip) * ny * nz + j * nz + k] = der[c];
                    if (out_f32 && out_f32[g]) {
                        ceg_o_set_gridpoint(gs->delta, coul ? lambda_coul : lambda_vdw, coul ? thr_coul : thr_vdw, der, q8);
                        for (int c = 0; c < 8; c++) out_f32[g][(size_t)c * npts + (i * ny + j) * nz + k] = q8[c];
                    }
                }
            }
        }
    }
}

/* coordinates.jl:58-66 offsetpoint (wrap_atom then affine map), 1-based grid coordinates */
static inline void offsetpoint(const ceg_cell_t* uc, const ceg_gridspec_t* gs, const double* p, double* s) {
    double abc[3], w[3];
    matvec(uc->invmat, p, abc);
    for (int i = 0; i < 3; i++) abc[i] = abc[i] - floor(abc[i]);
    matvec(uc->mat, abc, w);
    for (int i = 0; i < 3; i++) s[i] = (w[i] - gs->shift[i]) * gs->dims[i] / gs->size[i] + 1;
}

/* grids.jl:212-258 interpolate_grid, higher-order branch, with the 64x64 matrix `coeff`
 * (row-major: a[m] = sum_n coeff[m*64+n] X[n]; constants.jl:24-89).  grid = [8][nx][ny][nz] floats in K. */
void ceg_o_interpolate(const ceg_cell_t* uc, const ceg_gridspec_t* gs, const float* grid, int32_t is_vdw,
                       const double* coeff, int64_t n, const double* xyz, double* out) {
    const int64_t nx = gs->dims[0] + 1, ny = gs->dims[1] + 1, nz = gs->dims[2] + 1;
    const int64_t npts = nx * ny * nz;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < n; q++) {
        double s[3];
        offsetpoint(uc, gs, xyz + 3 * q, s);
        int64_t p0[3], p1[3];
        double r[3];
        const int64_t nn[3] = {nx, ny, nz};
        for (int i = 0; i < 3; i++) {
            p0[i] = (int64_t)floor(s[i]);
            p1[i] = p0[i] + (p0[i] != nn[i]);
            r[i] = s[i] - (double)p0[i];
        }
        double X[64];
        for (int c = 0; c < 8; c++)
            for (int iz = 0; iz < 2; iz++)
                for (int iy = 0; iy < 2; iy++)
                    for (int ix = 0; ix < 2; ix++) {
                        int64_t x = (ix ? p1[0] : p0[0]) - 1, y = (iy ? p1[1] : p0[1]) - 1, z = (iz ? p1[2] : p0[2]) - 1;
                        /* X order of grids.jl:227-244: index = ix(x1) + 2*iy + 4*iz within the channel */
                        X[8 * c + ix + 2 * iy + 4 * iz] = (double)grid[(size_t)c * npts + (x * ny + y) * nz + z];
                    }
        if (is_vdw) {
            int blocked = 0;
            for (int m = 0; m < 8; m++) blocked |= (X[m] > 5e6);
            if (blocked) { out[q] = 1e100; continue; }
        }
        double a[64];
        for (int m = 0; m < 64; m++) {
            double acc = 0.0;
            for (int nn2 = 0; nn2 < 64; nn2++) acc += coeff[m * 64 + nn2] * X[nn2];
            a[m] = acc;
        }
        double rx2 = r[0] * r[0], ry2 = r[1] * r[1], rz2 = r[2] * r[2];
        double rxs[4] = {1.0, r[0], rx2, rx2 * r[0]};
        double rys[4] = {1.0, r[1], ry2, ry2 * r[1]};
        double rzs[4] = {1.0, r[2], rz2, rz2 * r[2]};
        double ret = 0.0;
        for (int k = 0; k < 4; k++)
            for (int j = 0; j < 4; j++)
                for (int i = 0; i < 4; i++) ret += a[i + 4 * j + 16 * k] * rxs[i] * rys[j] * rzs[k];
        out[q] = ret;
    }
}

/* Grid coordinates used by the block lookup: round(Int, offsetpoint) (coordinates.jl:97-101),
 * returned 0-based; Julia's round is half-to-even = rint. */
void ceg_o_block_lookup(const ceg_cell_t* uc, const ceg_gridspec_t* gs, const uint8_t* mask, int64_t n,
                        const double* xyz, uint8_t* out) {
    const int64_t ny = gs->dims[1] + 1, nz = gs->dims[2] + 1;
    for (int64_t q = 0; q < n; q++) {
        double s[3];
        offsetpoint(uc, gs, xyz + 3 * q, s);
        int64_t a = (int64_t)rint(s[0]) - 1, b = (int64_t)rint(s[1]) - 1, c = (int64_t)rint(s[2]) - 1;
        out[q] = mask[(a * ny + b) * nz + c];
    }
}

/* ---- Ewald reciprocal space ---------------------------------------------------------------- */
typedef struct { double re, im; } cplx;
static inline cplx cmul(cplx a, cplx b) { cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cplx cispi2(double x) {
    /* cispi(2x): Julia reduces the argument exactly; do the same through the fractional part */
    double t = 2.0 * x;
    double fr = t - 2.0 * floor(t / 2.0); /* in [0,2) */
    cplx r;
    if (fr == 0.0) { r.re = 1; r.im = 0; }
    else if (fr == 0.5) { r.re = 0; r.im = 1; }
    else if (fr == 1.0) { r.re = -1; r.im = 0; }
    else if (fr == 1.5) { r.re = 0; r.im = -1; }
    else { r.re = cos(M_PI * fr); r.im = sin(M_PI * fr); }
    return r;
}
/* ewald.jl:69-86 make_line_pos! / make_line_neg!: tab[0..kmax] (pos) or tab[-kmax..kmax] (neg) */
static void make_line_pos(cplx* tab, int kmax, cplx e) {
    tab[0].re = 1; tab[0].im = 0;
    if (kmax >= 1) tab[1] = e;
    for (int i = 2; i <= kmax; i++) tab[i] = cmul(tab[i - 1], e);
}
static void make_line_neg(cplx* tab /* centre */, int kmax, cplx e) {
    cplx ce = {e.re, -e.im};
    if (kmax >= 1) tab[-1] = ce;
    for (int i = 2; i <= kmax; i++) tab[-i] = cmul(tab[-i + 1], ce);
    make_line_pos(tab, kmax, e);
}

/* ewald.jl:148-185 ewald_main_loop! over sites given in supercell-fractional coordinates
 * (setup_Eik, ewald.jl:109-145: cispi(2*f) + recurrences).  rows = kindices entries
 * (jy, jz, ix_first, ix_last, rangeidx) as 5 int32 each.  sums[n_k] complex (re,im pairs). */
void ceg_o_structure_factor(const int32_t* kmax, int64_t n_rows, const int32_t* rows, int64_t n_k, int64_t n_sites,
                            const double* frac, const double* charge, double* sums_reim) {
    int kx = kmax[0], ky = kmax[1], kz = kmax[2];
    cplx* ex = (cplx*)malloc(sizeof(cplx) * (kx + 1));
    cplx* ey = (cplx*)malloc(sizeof(cplx) * (2 * ky + 1));
    cplx* ez = (cplx*)malloc(sizeof(cplx) * (2 * kz + 1));
    cplx* sums = (cplx*)sums_reim;
    for (int64_t m = 0; m < n_k; m++) { sums[m].re = 0; sums[m].im = 0; }
    for (int64_t s = 0; s < n_sites; s++) {
        make_line_pos(ex, kx, cispi2(frac[3 * s]));
        make_line_neg(ey + ky, ky, cispi2(frac[3 * s + 1]));
        make_line_neg(ez + kz, kz, cispi2(frac[3 * s + 2]));
        double c = charge[s];
        for (int64_t rr = 0; rr < n_rows; rr++) {
            const int32_t* row = rows + 5 * rr;
            /* c*Eiky*Eikz evaluated left to right: (c*Eiky)*Eikz */
            cplx t = {c * ey[ky + row[0]].re, c * ey[ky + row[0]].im};
            cplx cyz = cmul(t, ez[kz + row[1]]);
            int64_t base = row[4];
            for (int ix = row[2]; ix <= row[3]; ix++) {
                cplx v = cmul(ex[ix], cyz);
                sums[base + (ix - row[2])].re += v.re;
                sums[base + (ix - row[2])].im += v.im;
            }
        }
    }
    free(ex); free(ey); free(ez);
}

/* ewald.jl:555-577 compute_ewald(ctx) for one placement of n_atoms charges at Cartesian xyz:
 * move_one_system! (ewald.jl:350-364) + ewald_main_loop! + the k-sum. */
double ceg_o_ewald_reciprocal(const double* invmat, const int32_t* kmax, int64_t n_rows, const int32_t* rows,
                              int64_t n_k, const double* kfactor, const double* Sf_reim, int32_t n_atoms,
                              const double* xyz, const double* q, double energy_net_charges,
                              double static_contribution) {
    double* frac = (double*)malloc(sizeof(double) * 3 * n_atoms);
    for (int a = 0; a < n_atoms; a++) matvec(invmat, xyz + 3 * a, frac + 3 * a);
    double* Sa = (double*)malloc(sizeof(double) * 2 * n_k);
    ceg_o_structure_factor(kmax, n_rows, rows, n_k, n_atoms, frac, q, Sa);
    double fa = 0.0, aa = 0.0;
    for (int64_t m = 0; m < n_k; m++) {
        double t = kfactor[m];
        double ref = Sf_reim[2 * m], imf = Sf_reim[2 * m + 1], rea = Sa[2 * m], ima = Sa[2 * m + 1];
        fa += t * (ref * rea + imf * ima);
        aa += t * (rea * rea + ima * ima);
    }
    free(frac); free(Sa);
    double host = 2 * (fa + energy_net_charges);
    double ads = aa + static_contribution;
    return host + ads;
}

void ceg_o_ewald_reciprocal_batch(const double* invmat, const int32_t* kmax, int64_t n_rows, const int32_t* rows,
                                  int64_t n_k, const double* kfactor, const double* Sf_reim, int64_t n_pos,
                                  int32_t n_atoms, const double* xyz, const double* q, double energy_net_charges,
                                  double static_contribution, double* out) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n_pos; i++)
        out[i] = ceg_o_ewald_reciprocal(invmat, kmax, n_rows, rows, n_k, kfactor, Sf_reim, n_atoms,
                                        xyz + (size_t)3 * n_atoms * i, q, energy_net_charges, static_contribution);
}

/* ---- guest-guest pair energies ------------------------------------------------------------------- */
/* interactions.jl:391-409 rule(d2) (and :367-387 through sqrt), summed in the kind order of
 * InteractionRuleSum: HardSphere, CoulombEwaldDirect, Coulomb, LennardJones, Buckingham. */
static double pair_energy(const ceg_pair_energy_t* p, double r2) {
    double e = 0.0;
    if (p->flags & CEG_PAIR_HS) e += (r2 < p->hs_r * p->hs_r) ? INFINITY : 0.0;
    if (p->flags & CEG_PAIR_EWALD) { double r = sqrt(r2); e += p->ewald_qq * erfc(p->ewald_alpha * r) / r; }
    if (p->flags & CEG_PAIR_COULOMB) { double r = sqrt(r2); e += p->coulomb_qq / r; }
    if (p->flags & CEG_PAIR_LJ) {
        double t = (p->lj_sigma * p->lj_sigma) / r2;
        double x6 = t * t * t;
        e += 4 * p->lj_eps * x6 * (x6 - 1) - p->lj_shift;
    }
    if (p->flags & CEG_PAIR_BUCK) {
        double r = sqrt(r2);
        e += (p->buck_A * exp(-p->buck_B * r) - p->buck_C / pow(r, 6)) - p->buck_shift;
    }
    return e;
}

/* energy.jl:407-427 single_contribution_vdw_noneighbour for n_trials placements (rigid molecule). */
void ceg_o_guest_energy(const ceg_cell_t* cell, double cutoff2, int32_t n_species, const ceg_pair_energy_t* table,
                        int64_t n_env, const double* env_xyz, const int32_t* env_species, const int32_t* env_mol,
                        int64_t n_trials, int32_t n_atoms, const int32_t* mol_species, int32_t skip_mol,
                        const double* trial_xyz, double* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t t = 0; t < n_trials; t++) {
        double energy = 0.0;
        for (int k2 = 0; k2 < n_atoms; k2++) {
            const double* p2 = trial_xyz + (t * n_atoms + k2) * 3;
            for (int64_t l1 = 0; l1 < n_env; l1++) {
                if (env_mol[l1] == skip_mol) continue;
                /* utils.jl:294-302 unsafe_periodic_distance2! */
                double b[3] = {p2[0] - env_xyz[3 * l1], p2[1] - env_xyz[3 * l1 + 1], p2[2] - env_xyz[3 * l1 + 2]};
                double f[3];
                matvec(cell->invmat, b, f);
                for (int d = 0; d < 3; d++) { double diff = f[d] + 0.5; f[d] = diff - floor(diff) - 0.5; }
                matvec(cell->mat, f, b);
                double d2 = norm2(b);
                if (d2 < cutoff2) energy += pair_energy(&table[(size_t)mol_species[k2] * n_species + env_species[l1]], d2);
            }
        }
        out[t] = energy;
    }
}

int ceg_o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm runs under torch.distributed.run, which exports OMP_NUM_THREADS=1 to its workers:
 * the timing harness sets the team size explicitly (all cores of the process' affinity mask). */
void ceg_o_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
