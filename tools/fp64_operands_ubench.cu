// FP64 pipe microbenchmark (experiment, not product): issue cost of FP64 instructions by operand kind.
//   mode 0: DFMA reg,const,const   1: DFMA reg,reg,const   2: DFMA reg,reg,reg   3: DMUL reg,reg   4: DADD reg,reg
//   mode 5: DMUL reg,const         6: DFMA reg,reg,reg with one operand shared by the ILP chains (reuse cache)
//   mode 7: DFMA reg,const,reg (Horner step with the coefficient in the middle)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int MODE>
__global__ void k(double* out, int iters, double a, double b) {
    double x[ILP], y[ILP], z[ILP];
    for (int i = 0; i < ILP; i++) { x[i] = 1.0 + 1e-3 * (threadIdx.x + i); y[i] = 0.999999 + 1e-9 * threadIdx.x * (i + 1); z[i] = 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) x[i] = fma(x[i], a, b);
                if (MODE == 1) x[i] = fma(x[i], y[i], b);
                if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);
                if (MODE == 3) x[i] = x[i] * y[i];
                if (MODE == 4) x[i] = x[i] + z[i];
                if (MODE == 5) x[i] = x[i] * a;
                if (MODE == 6) x[i] = fma(x[i], y[0], z[i]);
                if (MODE == 7) x[i] = fma(x[i], a, z[i]);
            }
    }
    double s = 0;
    for (int i = 0; i < ILP; i++) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int MODE>
void run(int warps_per_sm, double* buf, double clk) {
    int threads = 32 * warps_per_sm > 512 ? 512 : 32 * warps_per_sm;
    int blocks_per_sm = (32 * warps_per_sm + threads - 1) / threads;
    int blocks = 148 * blocks_per_sm, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<ILP, MODE><<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    // cycles per warp-instruction per SMSP = elapsed cycles / (instructions per warp * warps per SMSP)
    double cyc = best * 1e-3 * clk / ((double)iters * 8 * ILP * (warps_per_sm / 4.0));
    printf("mode %d ILP %d warps/SM %2d: %7.3f ms  %.2f cycles per warp-instruction per SMSP\n", MODE, ILP, warps_per_sm, best, cyc);
}
int main() {
    double* buf; cudaMalloc(&buf, sizeof(double) * 148 * 4 * 1024);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("clock %.0f MHz (attribute; the achieved clock may differ)\n", clk / 1e6);
    for (int w : {16, 32}) {
        run<4, 0>(w, buf, clk); run<4, 1>(w, buf, clk); run<4, 2>(w, buf, clk); run<4, 3>(w, buf, clk);
        run<4, 4>(w, buf, clk); run<4, 5>(w, buf, clk); run<4, 6>(w, buf, clk); run<4, 7>(w, buf, clk);
    }
    return 0;
}
