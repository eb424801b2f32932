// Microbenchmark: FP64 issue peaks on B200 (sm_100a).
//   dmma : mma.sync.aligned.m8n8k4.f64 with NACC independent accumulators per warp
//   dfma : fma.rn.f64 with NACC independent chains per thread
// Prints achieved TFLOP/s for several (warps per CTA, CTAs per SM) shapes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d L2 %d MB smem/SM %zu\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.sharedMemPerMultiprocessor);
    double* out; CK(cudaMalloc(&out, 8));
    int nsm = p.multiProcessorCount;
    const int iters = 20000;
    int warps_list[] = {4, 8, 16, 32};
    int ctas_list[] = {1, 2};
    for (int wi = 0; wi < 4; wi++) for (int ci = 0; ci < 2; ci++) {
        int warps = warps_list[wi], ctas = ctas_list[ci];
        if (warps * ctas > 64) continue;
        {
            float ms = time_it([&] { dmma_kernel<8><<<nsm * ctas, warps * 32>>>(out, iters, 1.0, 1.0); });
            double fl = (double)nsm * ctas * warps * iters * 8 * 512.0;
            printf("DMMA  nacc=8  warps/CTA=%2d CTAs/SM=%d : %8.3f ms  %7.2f TFLOP/s\n", warps, ctas, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma_kernel<16><<<nsm * ctas, warps * 32>>>(out, iters, 1.0, 1.0); });
            double fl = (double)nsm * ctas * warps * iters * 16 * 512.0;
            printf("DMMA  nacc=16 warps/CTA=%2d CTAs/SM=%d : %8.3f ms  %7.2f TFLOP/s\n", warps, ctas, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dfma_kernel<8><<<nsm * ctas, warps * 32>>>(out, iters, 1.0000001, 1e-9); });
            double fl = (double)nsm * ctas * warps * 32 * (double)iters * 8 * 2.0;
            printf("DFMA  nacc=8  warps/CTA=%2d CTAs/SM=%d : %8.3f ms  %7.2f TFLOP/s\n", warps, ctas, ms, fl / ms * 1e-9);
        }
    }
    // latency: single warp, 1 accumulator chain
    {
        float ms = time_it([&] { dmma_kernel<1><<<1, 32>>>(out, iters, 1.0, 1.0); });
        printf("DMMA dependent-chain latency: %.1f ns per mma (x clock for cycles)\n", ms * 1e6 / iters);
        float ms2 = time_it([&] { dfma_kernel<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9); });
        printf("DFMA dependent-chain latency: %.1f ns per fma\n", ms2 * 1e6 / iters);
    }
    return 0;
}
