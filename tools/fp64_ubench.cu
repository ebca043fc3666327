// FP64 pipe microbenchmark (experiment, not product): DFMA throughput vs operand kind, ILP and warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, bool REGOPS>
__global__ void k(double* out, int iters, double a, double b) {
    double x[ILP], y[ILP], z[ILP];
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x + i; y[i] = REGOPS ? 0.999999 + 1e-9 * threadIdx.x * (i + 1) : a; z[i] = REGOPS ? 1e-9 * (threadIdx.x + i) : b; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], y[i], z[i]);
    }
    double s = 0;
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, bool REGOPS>
void run(int warps_per_sm, double* buf) {
    int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm;
    int blocks_per_sm = (32 * warps_per_sm + threads - 1) / threads;
    int blocks = 148 * blocks_per_sm, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<ILP, REGOPS><<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    double fma_per_sm_clk = (double)iters * 8 * ILP * threads * blocks_per_sm / (best * 1e-3 * 1.965e9);
    printf("ILP %d regops %d warps/SM %2d: %7.3f ms  %6.2f TFLOP/s  %5.1f DFMA/clk/SM  (per-warp %.2f cyc/DFMA)\n", ILP, (int)REGOPS,
           warps_per_sm, best, 2.0 * iters * 8 * ILP * threads * blocks / (best * 1e-3) / 1e12, fma_per_sm_clk,
           (double)(best * 1e-3 * 1.965e9) / ((double)iters * 8 * ILP));
}
int main() {
    double* buf; cudaMalloc(&buf, sizeof(double) * 148 * 4 * 1024);
    for (int w : {4, 8, 16, 32, 64}) { run<1, false>(w, buf); run<1, true>(w, buf); }
    for (int w : {4, 16}) { run<2, true>(w, buf); run<4, true>(w, buf); run<8, true>(w, buf); run<8, false>(w, buf); }
    return 0;
}
