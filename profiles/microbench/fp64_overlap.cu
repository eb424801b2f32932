// Do DMMA (tensor) and DFMA (scalar FP64) share an execution pipe on B200?
// Even warps issue DMMA, odd warps issue DFMA; compare mixed time with each alone.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mixed(double* out, int iters, int mode) {
    // mode 0: all DMMA, 1: all DFMA, 2: even warps DMMA / odd warps DFMA
    const int warp = threadIdx.x >> 5;
    const bool do_mma = (mode == 0) || (mode == 2 && ((warp >> 2) & 1) == 0);  // every SMSP gets both kinds
    double c[8][2], f[8];
    for (int i = 0; i < 8; i++) { c[i][0] = c[i][1] = 0.0; f[i] = i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001;
    if (do_mma) {
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    } else {
        for (int it = 0; it < iters * 8; it++)  // x8: same stand-alone duration as the DMMA warps
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = fma(f[i], b, a);
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + f[i];
    if (s == 12345.678) out[0] = s;
}
int main() {
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000, warps = 16, ctas = 148;
    for (int mode = 0; mode < 3; mode++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        mixed<<<ctas, warps * 32>>>(out, iters, mode); cudaDeviceSynchronize();
        cudaEventRecord(e0); mixed<<<ctas, warps * 32>>>(out, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mma_w = (mode == 0) ? warps : (mode == 2 ? warps / 2 : 0), fma_w = (mode == 1) ? warps : (mode == 2 ? warps / 2 : 0);
        double tf_mma = ctas * mma_w * iters * 8.0 * 512 / ms * 1e-9, tf_fma = ctas * fma_w * 32 * (double)iters * 8 * 8 * 2 / ms * 1e-9;
        printf("mode %d: %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  sum %.2f\n", mode, ms, tf_mma, tf_fma, tf_mma + tf_fma);
    }
    return 0;
}
