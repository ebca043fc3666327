// Shared device/host helpers of libceg_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/ceg_cuda.h"

namespace ceg {

// ---- errors ---------------------------------------------------------------------------------
std::string& last_error();
int fail(const std::string& msg);
int fail_cuda(cudaError_t e, const char* what, const char* file, int line);

#define CEG_CK(call)                                                              \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return ::ceg::fail_cuda(_e, #call, __FILE__, __LINE__); \
    } while (0)

// every kernel launch of the library goes through this counter (ceg_kernel_launch_count)
void count_launch(int n = 1);

// Process-wide cache of device allocations (per device): cudaMalloc/cudaFree of the large scratch
// buffers cost up to hundreds of ms (cudaFree synchronises the device and unmaps), which would
// dominate a create-build-free cycle of a probe handle.  Buffers of >= 1 MiB are parked here on
// release and handed out again to the next request of a similar size; ceg_trim_device_cache() frees them.
void* pool_acquire(size_t bytes, size_t* got);
void pool_release(void* p, size_t bytes);
void pool_trim();

// RAII device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) pool_release(p, bytes);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t reserve(size_t n) {  // grow-only
        if (n <= bytes && p) return cudaSuccess;
        release();
        if (n == 0) n = 16;
        size_t got = 0;
        p = pool_acquire(n, &got);
        if (p) { bytes = got; return cudaSuccess; }
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {          // out of memory: drop the cache and retry once
            cudaGetLastError();
            pool_trim();
            e = cudaMalloc(&p, n);
        }
        if (e == cudaSuccess) bytes = n; else p = nullptr;
        return e;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- device structs ----------------------------------------------------------------------------
struct CellDev {
    double mat[9];     // column-major
    double invmat[9];
    int ortho;
    double safemin2;
};

struct GridSpecDev {
    int dims[3];
    int n[3];          // nodes = dims+1
    double size[3], shift[3], delta[3];
};

inline CellDev to_dev(const ceg_cell_t& c) {
    CellDev d;
    for (int i = 0; i < 9; i++) { d.mat[i] = c.mat[i]; d.invmat[i] = c.invmat[i]; }
    d.ortho = c.ortho;
    d.safemin2 = c.safemin2;
    return d;
}
inline GridSpecDev to_dev(const ceg_gridspec_t& g) {
    GridSpecDev d;
    for (int i = 0; i < 3; i++) {
        d.dims[i] = g.dims[i];
        d.n[i] = g.dims[i] + 1;
        d.size[i] = g.size[i];
        d.shift[i] = g.shift[i];
        d.delta[i] = g.delta[i];
    }
    return d;
}

// ---- arithmetic with the reference's operation order (no FMA contraction) ----------------------
// Used wherever a comparison decides a mask (cutoff, hard sphere, 1 Å, cell index), so that the
// device and the oracle take bit-identical decisions.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// mul!(out, m, v): row r = m[r,1]*v1 + m[r,2]*v2 + m[r,3]*v3, left to right
__device__ __forceinline__ void matvec_x(const double* __restrict__ m, const double* v, double* out) {
#pragma unroll
    for (int r = 0; r < 3; r++) out[r] = xadd(xadd(xmul(m[r], v[0]), xmul(m[3 + r], v[1])), xmul(m[6 + r], v[2]));
}
__device__ __forceinline__ double norm2_x(const double* v) {
    return xadd(xadd(xmul(v[0], v[0]), xmul(v[1], v[1])), xmul(v[2], v[2]));
}

// One grid node as the interpolation kernels read it: the eight channels of `.grid` in 32 bytes (one DRAM sector).
struct alignas(32) Rec8 {
    float4 lo, hi;         // f, fx, fy, fz | fxy, fxz, fyz, fxyz
};
// Place of node (x, y, z) among the records: node-major, or 2x2x2-node bricks of 256 bytes (interp.cuh).
__device__ __forceinline__ int64_t record_index(bool block, int x, int y, int z, int ny, int nz) {
    if (!block) return ((int64_t)x * ny + y) * nz + z;
    const int64_t nby = (ny + 1) >> 1, nbz = (nz + 1) >> 1;
    return ((((int64_t)(x >> 1) * nby + (y >> 1)) * nbz + (z >> 1)) << 3) + (((x & 1) * 2 + (y & 1)) * 2 + (z & 1));
}
// parse_grid's unit conversion of one `.grid` value (grids.jl:78: Float32(value * GRID_TO_KELVIN), the product formed in
// double); scale == 1 leaves the value as it is.
__device__ __forceinline__ float record_value(float f, double scale) {
    return scale == 1.0 ? f : (float)xmul((double)f, scale);
}

// utils.jl:210-246 periodic_distance2_fromcartesian!, literally: `b` = Cartesian difference on
// entry, the vector the reference leaves in its buffer on exit (stale last-tried image included).
__device__ __noinline__ double periodic_distance2_ref(double* b, const CellDev& c) {
    double f[3];
    matvec_x(c.invmat, b, f);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double diff = xadd(f[i], 0.5);
        f[i] = xsub(xsub(diff, floor(diff)), 0.5);
    }
    matvec_x(c.mat, f, b);
    double ref2 = norm2_x(b);
    if (c.ortho || ref2 <= c.safemin2) return ref2;
    for (int i = 0; i < 3; i++) {
        f[i] = xadd(f[i], 1.0);
        matvec_x(c.mat, f, b);
        double n2 = norm2_x(b);
        if (n2 < ref2) return n2;
        f[i] = xsub(f[i], 2.0);
        matvec_x(c.mat, f, b);
        n2 = norm2_x(b);
        if (n2 < ref2) return n2;
        f[i] = xadd(f[i], 1.0);
    }
    return ref2;
}

// coordinates.jl:58-66 offsetpoint: wrap into the unit cell, then map to 1-based grid coordinates.
__device__ __forceinline__ void offsetpoint_x(const CellDev& uc, const GridSpecDev& gs, const double* p, double* s) {
    double abc[3], w[3];
    matvec_x(uc.invmat, p, abc);
#pragma unroll
    for (int i = 0; i < 3; i++) abc[i] = xsub(abc[i], floor(abc[i]));
    matvec_x(uc.mat, abc, w);
#pragma unroll
    for (int i = 0; i < 3; i++)
        s[i] = xadd(xdiv(xmul(xsub(w[i], gs.shift[i]), (double)gs.dims[i]), gs.size[i]), 1.0);
}

}  // namespace ceg
