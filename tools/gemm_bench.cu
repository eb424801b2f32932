// Development aid: throughput of gemm.cu's NT DMMA GEMM as a function of the contraction length K
// (rank-K update of the lower triangle of an n x n matrix).  Build: see tools/gemm_bench.sh
#include <cstdio>
#include <cstdlib>
#include "../gptools_b200/csrc/internal.h"
int main(int argc, char** argv) {
    int T = argc > 1 ? atoi(argv[1]) : 96;  // tiles per side
    long n = (long)T * 128;
    double *C, *P;
    cudaMalloc(&C, n * n * 8);
    cudaMalloc(&P, n * 1024 * 8);
    cudaMemset(C, 0, n * n * 8);
    cudaMemset(P, 0, n * 1024 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int Ks[] = {128, 256, 384, 512, 1024};
    for (int lower = 1; lower >= 0; lower--)
        for (int ki = 0; ki < 5; ki++) {
            int K = Ks[ki];
            GemmParams g;
            g.C = C; g.ldc = n; g.A = P; g.lda = K; g.B = P; g.ldb = K;
            g.tiles_m = T; g.tiles_n = T; g.K = K; g.alpha = -1.0; g.beta = 1.0; g.lower_only = lower; g.kbegin_row = 0;
            launch_gemm_nt(g, 0);
            cudaEventRecord(e0);
            int reps = 4;
            for (int r = 0; r < reps; r++) launch_gemm_nt(g, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double tiles = lower ? (double)T * (T + 1) / 2 : (double)T * T;
            double fl = tiles * 2.0 * 128 * 128 * K * reps;
            printf("T=%d lower=%d K=%4d: %.3f ms/launch  %.2f TFLOP/s\n", T, lower, K, ms / reps, fl / (ms * 1e-3) * 1e-12);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
