// Random-gather microbenchmark (experiment): 8 x 32-byte loads per thread over a footprint F.
//  mode 0: 8 independent random sectors anywhere in F            (8 pages/thread)
//  mode 1: 8 random sectors inside ONE random 2 MB page           (1 page/thread)
//  mode 2: 4 random aligned 64-byte units anywhere in F           (4 pages, no over-fetch)
//  mode 3: 4 random MISaligned sector pairs (offset 32 B)         (4 pages, 64 B-granularity over-fetch)
//  mode 4: the tricubic stencil on a node-major [x][y][z] grid of 32-byte records (the kernel's pattern)
//  mode 5: the same stencil on a grid stored as 4x4x4-node bricks of 2 KB (one DRAM page per brick)
//  mode 6: the same stencil on 2x2x2-node blocks of 256 B
//  mode 7: 8 random sectors inside one 2 MB page chosen PER WARP (all lanes of a load share the page)
//  mode 8: the same with a 64 KB region per warp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__device__ __forceinline__ float ld32(const float* p) { float a,b,c,d,e,f,g,h; asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a),"=f"(b),"=f"(c),"=f"(d),"=f"(e),"=f"(f),"=f"(g),"=f"(h) : "l"(p)); return a+b+c+d+e+f+g+h; }
__global__ void __launch_bounds__(256) k(const float* base, uint64_t n_sectors, int mode, int64_t n, float* out, uint64_t seed) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        uint64_t r = mix(seed + q);
        float acc = 0;
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) { uint64_t s = mix(r + i) % n_sectors; acc += ld32(base + s * 8); }
        } else if (mode == 1) {
            uint64_t page = (r % (n_sectors >> 16)) << 16;     // 65536 sectors = 2 MB
#pragma unroll
            for (int i = 0; i < 8; i++) { uint64_t s = page + (mix(r + i) & 0xffff); acc += ld32(base + s * 8); }
        } else if (mode >= 7) {
            const uint64_t rw = mix(seed + (uint64_t)(q >> 5) * 1315423911ull);
            const int bits = mode == 7 ? 16 : 11;                 // sectors per region: 2 MB or 64 KB
            const uint64_t region = (rw % (n_sectors >> bits)) << bits;
#pragma unroll
            for (int i = 0; i < 8; i++) { uint64_t s = region + (mix(r + i) & ((1ull << bits) - 1)); acc += ld32(base + s * 8); }
        } else if (mode >= 4) {
            // virtual grid nx x 204 x 192 nodes filling the footprint
            const uint64_t ny = 204, nz = 192, nx = n_sectors / (ny * nz);
            const uint64_t x = mix(r + 1) % (nx - 1), y = mix(r + 2) % (ny - 1), z = mix(r + 3) % (nz - 1);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint64_t cx = x + (c >> 2), cy = y + ((c >> 1) & 1), cz = z + (c & 1);
                uint64_t s;
                if (mode == 4) s = (cx * ny + cy) * nz + cz;
                else if (mode == 5) s = (((cx >> 2) * (ny / 4) + (cy >> 2)) * (nz / 4) + (cz >> 2)) * 64 + (((cx & 3) * 4 + (cy & 3)) * 4 + (cz & 3));
                else s = (((cx >> 1) * (ny / 2) + (cy >> 1)) * (nz / 2) + (cz >> 1)) * 8 + (((cx & 1) * 2 + (cy & 1)) * 2 + (cz & 1));
                acc += ld32(base + s * 8);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint64_t s = (mix(r + i) % (n_sectors - 4)) & ~1ull;
                if (mode == 3) s += 1;
                acc += ld32(base + s * 8) + ld32(base + s * 8 + 8);
            }
        }
        out[q] = acc;
    }
}
int main() {
    const int64_t n = 10000000;
    float* out; cudaMalloc(&out, n * 4);
    for (size_t mb : {256, 1024, 4080}) {
        float* base; size_t bytes = mb << 20; cudaMalloc(&base, bytes); cudaMemset(base, 0, bytes);
        for (int mode : {0, 1, 7, 8, 4}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0); k<<<148 * 16, 256>>>(base, bytes / 32, mode, n, out, rep * 7919); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            printf("F %5zu MB mode %d: %.3f ms  %.1f G sectors/s  %.0f GB/s useful\n", mb, mode, best, n * 8 / best / 1e6, n * 256.0 / best / 1e6);
        }
        cudaFree(base);
    }
    return 0;
}
