// FP64 tensor-core GEMM for the blocked single-matrix algorithms:  C = beta*C + alpha * A * B^T.
//
// Both operands are row-major with the contraction index contiguous ("NT"), which is the only
// form the blocked Cholesky / trtri / lauum / T K T^T / triangular-solve paths need (symmetric
// operands and pre-transposed inverses make every product NT).
//
// CTA tile 128x64x16, 128 threads = 4 warps (2 x 2), warp tile 64x32 = 8x4 DMMA.8x8x4 tiles,
// double-buffered cp.async pipeline (61 KB), 166 registers, THREE CTAs per SM: the read-modify-write epilogue of one
// CTA (an L2/HBM round trip with the accumulators pinned in registers), its pipeline fill and its barriers overlap
// the main loops of the other two.  A 128x128 CTA alone on the SM spent ~20% of a K=128 rank update in the
// epilogue; thread-level parallelism across CTAs beat a deeper pipeline (profiles/r01f_gemm_notes.txt).  Shared rows are padded to 20 doubles: the fragment read
// A[g][k0+t] then maps the 16 lanes of a half-warp to 16 distinct 8-byte banks (g*20+t mod 16
// = 4g+t), i.e. conflict-free LDS.64 for both operands.
// Roofline: FP64 tensor pipe (measured cuBLAS Dgemm 35.5 TFLOP/s on this pool's B200).
#include "common.cuh"
#include "internal.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 2, LDS_ = BK + 4;
constexpr int THREADS = 128;
constexpr size_t SMEM_BYTES = (size_t)STAGES * (BM + BN) * LDS_ * sizeof(double);

__global__ void __launch_bounds__(THREADS, 3) gemm_nt_kernel(GemmParams p) {
    extern __shared__ __align__(16) double smem[];
    const int bm = blockIdx.y, bn = blockIdx.x;
    if (p.lower_only && bn > 2 * bm + 1) return;  // bn counts 64-wide half tiles; diagonal 128-blocks are full
    const int kstart = p.kbegin_row ? bm * BM : 0;
    const int kend = (p.kend_row && (bm + 1) * BM < p.K) ? (bm + 1) * BM : p.K;
    const int nk = (kend - kstart) / BK;
    const long zb = blockIdx.z;  // batch index
    const double* __restrict__ Ag = p.A + zb * p.strideA + (long)bm * BM * p.lda + kstart;
    const double* __restrict__ Bg = p.B + zb * p.strideB + (long)bn * BN * p.ldb + kstart;
    double* __restrict__ Cg = p.C + zb * p.strideC;
    double* sA = smem;
    double* sB = smem + STAGES * BM * LDS_;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;

    auto load_stage = [&](int stage, int kc) {
        double* a_dst = sA + stage * BM * LDS_;
        double* b_dst = sB + stage * BN * LDS_;
#pragma unroll
        for (int i = tid; i < BM * (BK / 2); i += THREADS) {
            const int r = i >> 3, c = (i & 7) * 2;
            cp_async16(a_dst + r * LDS_ + c, Ag + (long)r * p.lda + kc * BK + c);
        }
#pragma unroll
        for (int i = tid; i < BN * (BK / 2); i += THREADS) {
            const int r = i >> 3, c = (i & 7) * 2;
            cp_async16(b_dst + r * LDS_ + c, Bg + (long)r * p.ldb + kc * BK + c);
        }
    };

    // The C tile is read only in the epilogue; all CTAs of a wave reach it together, which turns the 128 KB
    // read-modify-write per tile into an HBM burst with the tensor pipe idle.  Pull the tile into L2 now so
    // the fetch overlaps the main loop (the operands are L2 resident panels).
    if (p.beta != 0.0) {
        const char* cbase = reinterpret_cast<const char*>(Cg + (long)bm * BM * p.ldc + (long)bn * BN);
#pragma unroll
        for (int i = tid; i < BM * 4; i += THREADS)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(cbase + (long)(i >> 2) * p.ldc * 8 + (i & 3) * 128));
    }

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kc = 0; kc < nk; kc++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kc + STAGES - 1;
        if (nxt < nk) load_stage(nxt % STAGES, nxt);
        cp_async_commit();
        const double* a_s = sA + (kc % STAGES) * BM * LDS_ + (wm * 64 + g) * LDS_ + t;
        const double* b_s = sB + (kc % STAGES) * BN * LDS_ + (wn * 32 + g) * LDS_ + t;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a_s[i * 8 * LDS_ + kk * 4];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = b_s[j * 8 * LDS_ + kk * 4];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each lane owns C[g][2t], C[g][2t+1] of every 8x8 tile -> one 16-byte store
    const long row0 = (long)bm * BM + wm * 64 + g;
    const long col0 = (long)bn * BN + wn * 32 + 2 * t;
    if (p.beta != 0.0) {
        // read-modify-write in batches of 8 independent 16-byte loads per lane: issued back to back they cost one
        // L2 round trip per batch instead of one per element (stores may alias the loads as far as nvcc knows)
#pragma unroll
        for (int i0 = 0; i0 < 8; i0 += 2) {
            double2 old[2][4];
#pragma unroll
            for (int ii = 0; ii < 2; ii++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    old[ii][j] = __ldcg(reinterpret_cast<const double2*>(Cg + (row0 + (i0 + ii) * 8) * p.ldc + col0 + j * 8));
#pragma unroll
            for (int ii = 0; ii < 2; ii++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double2 v;
                    v.x = p.alpha * acc[i0 + ii][j][0] + p.beta * old[ii][j].x;
                    v.y = p.alpha * acc[i0 + ii][j][1] + p.beta * old[ii][j].y;
                    *reinterpret_cast<double2*>(Cg + (row0 + (i0 + ii) * 8) * p.ldc + col0 + j * 8) = v;
                }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double2 v;
                v.x = p.alpha * acc[i][j][0];
                v.y = p.alpha * acc[i][j][1];
                *reinterpret_cast<double2*>(Cg + (row0 + i * 8) * p.ldc + col0 + j * 8) = v;
            }
    }
}

}  // namespace

// rows of a 128-column output (tiles_n == 1) that fill the machine with whole waves of CTAs
long gemm_rows_per_wave_n128(int num_sms) { return (long)num_sms * 3 /* CTAs per SM */ * 128 / (128 / BN); }

void launch_gemm_nt(const GemmParams& p, cudaStream_t s) {
    cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (p.tiles_m <= 0 || p.tiles_n <= 0) return;
    dim3 grid(p.tiles_n * (128 / BN), p.tiles_m, p.batch > 0 ? p.batch : 1);
    gemm_nt_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(p);
}
