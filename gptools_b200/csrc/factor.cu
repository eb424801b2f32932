// Single-matrix blocked Cholesky building blocks (the GEMM-shaped parts live in gemm.cu):
//   potrf_diag   : 128x128 diagonal block factorisation + explicit inverse of the block (so that the
//                  panel solve, the triangular solves of predict and trtri all become DMMA GEMMs),
//                  the block's share of z = L^{-1} y, its log-determinant share and LAPACK-style info
//   panel_gemv   : right-looking update of the remaining right-hand side
//   backsolve    : alpha = L^{-T} z, one block per launch
//   grad_reduce  : 1/2 tr((a a^T - S) dK_p) with the dK tiles regenerated on the fly (never stored)
// Replaces scipy.linalg.cholesky / cho_solve in compute_K_L_alpha_ll (gaussian_process.py:1452-1504).
#include "common.cuh"
#include "internal.h"

namespace {

constexpr int NB = GPT_NB;
constexpr int LDB = NB + 4;  // stride == 4 (mod 16): conflict-free DMMA fragment loads from the block
constexpr int NBLK = NB / 8;
constexpr size_t POTRF_SMEM = ((size_t)NB * LDB + 2 * NB) * sizeof(double);

// ---- 128x128 diagonal block: blocked in-place Gauss-Jordan sweep on 8x8 sub-blocks ------------------------------
// The first version of this kernel eliminated one column per __syncthreads with scalar FP64 (207 us per block,
// and it sits on the critical path of the blocked Cholesky: 13% of an M = 24576 factorisation, all of an M = 4000
// one).  Here everything except the sixteen 8x8 pivot factorisations is an 8x8x8 product = two DMMA.8x8x4
// (same scheme as batched4.cu::potrf_inv_tile, tools/tile_model.py: blocked_gj_inverse_factor).
// In-place layout of V while pivot block J is processed:
//   (I,K), K > J : Schur complement;  (I,J): panel L_IJ (also written to A), later Y_IJ = -L_IJ Xp_J;
//   K < J: Y_IK (rows > J) / X_JK (rows <= J, final inverse).
__device__ __forceinline__ double* blk8(double* V, int I, int K) { return V + (8 * I) * LDB + 8 * K; }

// C (8x8, in place) = sc * C + sa * A * B^T (nt) or sc * C + sa * A * B (nn); one warp.  G != null: also to global.
template <bool NN>
__device__ __forceinline__ void blk_mma(double* C, const double* A, const double* B, double sa, double sc, int g,
                                        int t, double* G = nullptr, long ldg = 0) {
    const double a0 = sa * A[g * LDB + t], a1 = sa * A[g * LDB + 4 + t];
    double b0, b1;
    if (NN) {
        b0 = B[t * LDB + g];
        b1 = B[(4 + t) * LDB + g];
    } else {
        b0 = B[g * LDB + t];
        b1 = B[g * LDB + 4 + t];
    }
    double2 c = make_double2(0.0, 0.0);
    if (sc != 0.0) {
        c = *reinterpret_cast<const double2*>(C + g * LDB + 2 * t);
        c.x *= sc;
        c.y *= sc;
    }
    dmma884(c.x, c.y, a0, b0);
    dmma884(c.x, c.y, a1, b1);
    *reinterpret_cast<double2*>(C + g * LDB + 2 * t) = c;
    if (G) *reinterpret_cast<double2*>(G + g * ldg + 2 * t) = c;
}

// Pivot block P (8x8 SPD, lower valid): in place -> chol(P)^{-1} (zeros above the diagonal); chol(P) itself goes
// to the global block G (zeros above).  Every lane of one warp redundantly, fully unrolled in registers.
// Returns sum(log L_ii).
__device__ __forceinline__ double pivot8(double* P, double* G, long ldg, int lane, int row0, int* s_info) {
    double p[8][8], lc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) p[i][j] = P[i * LDB + j];
    double dsave[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double d = p[j][j];
        if (!(d > 0.0)) {
            if (lane == 0 && *s_info == 0) *s_info = row0 + j + 1;
            d = 1.0;
        }
        dsave[j] = d;
        const double rinv = __drcp_rn(d);
        double w[8];
#pragma unroll
        for (int c = 0; c < 8; c++) w[c] = (c > j) ? p[c][j] : ((c < j) ? p[j][c] : 0.0);
#pragma unroll
        for (int r = j + 1; r < 8; r++) lc[r][j] = w[r];  // unscaled column j of the Cholesky factor
#pragma unroll
        for (int r = j + 1; r < 8; r++) {
            const double mult = w[r] * rinv;
#pragma unroll
            for (int c = 0; c <= r; c++) {
                if (c == j) p[r][c] = -mult;
                else p[r][c] -= mult * w[c];
            }
        }
    }
    double rs[8];
#pragma unroll
    for (int i = 0; i < 8; i++) rs[i] = 1.0 / sqrt(dsave[i]);
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double xval = (j < i) ? p[i][j] * rs[i] : ((j == i) ? rs[i] : 0.0);
            const double lval = (j < i) ? lc[i][j] * rs[j] : ((j == i) ? dsave[i] * rs[i] : 0.0);
            if (lane == ((i * 8 + j) & 31)) {
                P[i * LDB + j] = xval;
                G[i * ldg + j] = lval;
            }
        }
    }
    // one log per lane (lanes 0..7), summed over the warp: sum(log L_ii) = 1/2 sum(log d_i)
    double mine = 1.0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (lane == i) mine = dsave[i];
    return 0.5 * warp_sum(log(mine));
}

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, long lda, double* __restrict__ inv,
                                                            double* __restrict__ yk, double* __restrict__ logdet_part,
                                                            int* __restrict__ info, int row0) {
    extern __shared__ __align__(16) double sm[];
    double* V = sm;                  // NB x LDB
    double* yv = sm + NB * LDB;      // right-hand side block
    __shared__ int s_info;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        V[r * LDB + c] = (c <= r) ? A[(long)r * lda + c] : 0.0;
    }
    if (tid < NB && yk != nullptr) yv[tid] = yk[tid];
    if (tid == 0) s_info = 0;
    __syncthreads();
    // the factor is a GEMM operand later: the blocks above the block diagonal must be clean zeros
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if ((c >> 3) > (r >> 3)) A[(long)r * lda + c] = 0.0;
    }
    double logacc = 0.0;  // warp 7
    if (warp == 7) logacc = pivot8(blk8(V, 0, 0), A, lda, lane, row0, &s_info);
    __syncthreads();

#pragma unroll 1
    for (int J = 0; J < NBLK; J++) {
        const double* Xp = blk8(V, J, J);
        // ---- phase A: multiply by the pivot inverses.  Items:
        //   [0, J)            : X_JK = Xp_J Y_JK (K < J); for K = J-1 first Y_{J,J-1} = -L_{J,J-1} Xp_{J-1}
        //   [J, J+n)          : L_IJ = V_IJ Xp_J^T (I > J), the final factor -> also to A
        //   [J+n, J+2n)       : Y_{I,J-1} = -L_{I,J-1} Xp_{J-1} (I > J), the deferred last stage of step J-1
        const int n = NBLK - 1 - J;
        const int nitems = J + n + (J > 0 ? n : 0);
        for (int item = warp; item < nitems; item += 8) {
            if (item < J) {
                const int K = item;
                if (K == J - 1) {
                    blk_mma<true>(blk8(V, J, K), blk8(V, J, K), blk8(V, J - 1, J - 1), -1.0, 0.0, g, t);
                    __syncwarp();
                }
                blk_mma<true>(blk8(V, J, K), Xp, blk8(V, J, K), 1.0, 0.0, g, t);
            } else if (item < J + n) {
                const int I = J + 1 + (item - J);
                blk_mma<false>(blk8(V, I, J), blk8(V, I, J), Xp, 1.0, 0.0, g, t, A + (long)(8 * I) * lda + 8 * J, lda);
            } else {
                const int I = J + 1 + (item - J - n);
                blk_mma<true>(blk8(V, I, J - 1), blk8(V, I, J - 1), blk8(V, J - 1, J - 1), -1.0, 0.0, g, t);
            }
        }
        __syncthreads();
        if (J + 1 == NBLK) break;
        // ---- phase B: warp 7 finalises and factors the next pivot block while the others apply the rank-8 update
        if (warp == 7) {
            blk_mma<false>(blk8(V, J + 1, J + 1), blk8(V, J + 1, J), blk8(V, J + 1, J), -1.0, 1.0, g, t);
            __syncwarp();
            logacc += pivot8(blk8(V, J + 1, J + 1), A + (long)(8 * (J + 1)) * lda + 8 * (J + 1), lda, lane,
                             row0 + 8 * (J + 1), &s_info);
        } else {
            int cnt = 0;
            for (int I = J + 1; I < NBLK; I++) {
                for (int K = J + 1; K <= I; K++) {
                    if (I == J + 1) continue;  // (J+1, J+1) belongs to warp 7
                    if ((cnt++ % 7) == warp)
                        blk_mma<false>(blk8(V, I, K), blk8(V, I, J), blk8(V, K, J), -1.0, 1.0, g, t);
                }
                for (int K = 0; K < J; K++)
                    if ((cnt++ % 7) == warp)
                        blk_mma<true>(blk8(V, I, K), blk8(V, I, J), blk8(V, J, K), -1.0, 1.0, g, t);
            }
        }
        __syncthreads();
    }

    // V now holds X = L^{-1} in its lower block triangle (diagonal blocks carry their own zeros)
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        inv[r * NB + c] = ((c >> 3) <= (r >> 3)) ? V[r * LDB + c] : 0.0;
    }
    if (yk != nullptr) {
        // z_k = X y_k : two lanes per row
        const int r = tid >> 1, h = tid & 1;
        double sacc = 0.0;
        for (int c = h; c <= r; c += 2) sacc += V[r * LDB + c] * yv[c];
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        if (h == 0) yk[r] = sacc;
    }
    if (tid == 7 * 32) {
        *logdet_part = logacc;
        if (s_info != 0) atomicCAS(info, 0, s_info);
    }
}


__global__ void __launch_bounds__(256) panel_gemv_kernel(const double* __restrict__ P, int rows,
                                                         const double* __restrict__ zk, double* __restrict__ y) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const double* row = P + (long)warp * NB;
    double s = 0.0;
#pragma unroll
    for (int c = lane; c < NB; c += 32) s += row[c] * zk[c];
    s = warp_sum(s);
    if (lane == 0) y[warp] -= s;
}

__global__ void __launch_bounds__(256) backsolve_step_kernel(const double* __restrict__ L, long ld, int k,
                                                             const double* __restrict__ inv_k,
                                                             double* __restrict__ z, double* __restrict__ alpha) {
    __shared__ double ak[NB];
    const int tid = threadIdx.x;
    if (tid < NB) {
        double s = 0.0;
        for (int i = tid; i < NB; i++) s += inv_k[i * NB + tid] * z[k * NB + i];
        ak[tid] = s;
        if (blockIdx.x == 0) alpha[k * NB + tid] = s;
    }
    __syncthreads();
    const int c = blockIdx.x * 256 + tid;
    if (c < k * NB) {
        const double* col = L + (long)k * NB * ld + c;
        double s = 0.0;
#pragma unroll 4
        for (int r = 0; r < NB; r++) s += col[(long)r * ld] * ak[r];
        z[c] -= s;
    }
}

__global__ void transpose_kernel(double* __restrict__ out, long ldo, const double* __restrict__ in, long ldi,
                                 int rows, int cols) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int ir = by + r, ic = bx + threadIdx.x;
        tile[r][threadIdx.x] = (ir < rows && ic < cols) ? in[(long)ir * ldi + ic] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int orow = bx + r, ocol = by + threadIdx.x;
        if (orow < cols && ocol < rows) out[(long)orow * ldo + ocol] = tile[threadIdx.x][r];
    }
}

__global__ void copy2d_kernel(double* __restrict__ out, long ldo, const double* __restrict__ in, long ldi,
                              int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < cols && r < rows) out[(long)r * ldo + c] = in[(long)r * ldi + c];
}

__global__ void add_diag_kernel(double* __restrict__ A, long lda, const double* __restrict__ d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(long)i * lda + i] += d[i];
}

__global__ void fill_kernel(double* __restrict__ p, long n, double v) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void identity_pad_kernel(double* __restrict__ A, long lda, int n_valid, int n_pad) {
    // rows/cols >= n_valid: zero, with ones on the diagonal
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= n_pad || r >= n_pad) return;
    if (r >= n_valid || c >= n_valid) A[(long)r * lda + c] = (r == c) ? 1.0 : 0.0;
}

constexpr int GT = 64;

__global__ void __launch_bounds__(256) grad_reduce_kernel(GradReduceParams p) {
    __shared__ double sx_r[GT * GPT_MAX_DIM], sx_c[GT * GPT_MAX_DIM];
    __shared__ int32_t sn_r[GT * GPT_MAX_DIM], sn_c[GT * GPT_MAX_DIM];
    __shared__ double red[8][GPT_MAX_PARAMS];
    const int bx = blockIdx.x, by = blockIdx.y;
    const int cta = by * gridDim.x + bx;
    const int tid = threadIdx.x;
    double acc[GPT_MAX_PARAMS];
#pragma unroll
    for (int q = 0; q < GPT_MAX_PARAMS; q++) acc[q] = 0.0;
    if (bx <= by) {
        const int D = p.cp.D;
        const int r0 = by * GT, c0 = bx * GT;
        for (int i = tid; i < GT * D; i += 256) {
            const int r = i / D, d = i - r * D;
            const bool okr = (r0 + r) < p.N, okc = (c0 + r) < p.N;
            sx_r[r * GPT_MAX_DIM + d] = okr ? p.X[(long)(r0 + r) * D + d] : 0.0;
            sn_r[r * GPT_MAX_DIM + d] = okr ? p.n[(long)(r0 + r) * D + d] : 0;
            sx_c[r * GPT_MAX_DIM + d] = okc ? p.X[(long)(c0 + r) * D + d] : 0.0;
            sn_c[r * GPT_MAX_DIM + d] = okc ? p.n[(long)(c0 + r) * D + d] : 0;
        }
        __syncthreads();
        const int ty = tid >> 4, tx = tid & 15;
#pragma unroll 1
        for (int a = 0; a < 4; a++) {
            const int lr = ty + 16 * a, r = r0 + lr;
#pragma unroll 1
            for (int b = 0; b < 4; b++) {
                const int lc = tx + 16 * b, c = c0 + lc;
                if (r >= p.N || c > r) continue;
                double w = -p.S[(long)r * p.lds + c];
                if (p.a) w += p.a[r] * p.a[c];
                if (r == c) w *= 0.5;
                if (p.cp.kid == GPT_KERNEL_SE) {
                    double dk[2 + GPT_MAX_DIM];
                    se_cov_all(p.cp, sx_r + lr * GPT_MAX_DIM, sn_r + lr * GPT_MAX_DIM, sx_c + lc * GPT_MAX_DIM,
                               sn_c + lc * GPT_MAX_DIM, dk);
                    for (int q = 0; q < p.nidx; q++) acc[q] += w * dk[1 + p.idx[q]];
                } else {
                    for (int q = 0; q < p.nidx; q++)
                        acc[q] += w * cov_eval(p.cp, sx_r + lr * GPT_MAX_DIM, sn_r + lr * GPT_MAX_DIM,
                                               sx_c + lc * GPT_MAX_DIM, sn_c + lc * GPT_MAX_DIM, p.idx[q]);
                }
            }
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = 0; q < p.nidx; q++) {
        const double s = warp_sum(acc[q]);
        if (lane == 0) red[warp][q] = s;
    }
    __syncthreads();
    if (tid < p.nidx) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += red[w][tid];
        p.partials[(long)cta * p.nidx + tid] = s;
    }
}

__global__ void grad_final_kernel(const double* __restrict__ partials, int nctas, int nidx, double* __restrict__ out) {
    // fixed-order (deterministic) sum of the per-CTA partials
    const int q = blockIdx.x;
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nctas; i += 256) s += partials[(long)i * nidx + q];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[q] = sh[0];
}

__global__ void trace_sumsq_kernel(const double* __restrict__ A, long lda, const double* __restrict__ v, int n,
                                   double* __restrict__ out) {
    __shared__ double sh[2][256];
    double tr = 0.0, ss = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        tr += A[(long)i * lda + i];
        ss += v[i] * v[i];
    }
    sh[0][threadIdx.x] = tr;
    sh[1][threadIdx.x] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sh[0][0];
        out[1] = sh[1][0];
    }
}

}  // namespace

void launch_potrf_diag(double* Ablk, long lda, double* inv, double* yk, double* logdet_part, int* info, int row0,
                       int /*nvalid*/, cudaStream_t s) {
    cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
    potrf_diag_kernel<<<1, 256, POTRF_SMEM, s>>>(Ablk, lda, inv, yk, logdet_part, info, row0);
}

void launch_panel_gemv(const double* P, int rows, const double* zk, double* y, cudaStream_t s) {
    if (rows <= 0) return;
    panel_gemv_kernel<<<(rows + 7) / 8, 256, 0, s>>>(P, rows, zk, y);
}

void launch_backsolve_step(const double* L, long ld, int k, const double* inv_k, double* z, double* alpha,
                           cudaStream_t s) {
    int grid = (k * NB + 255) / 256;
    if (grid < 1) grid = 1;
    backsolve_step_kernel<<<grid, 256, 0, s>>>(L, ld, k, inv_k, z, alpha);
}

void launch_transpose(double* out, long ldo, const double* in, long ldi, int rows, int cols, cudaStream_t s) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(out, ldo, in, ldi, rows, cols);
}

void launch_copy2d(double* out, long ldo, const double* in, long ldi, int rows, int cols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((cols + 255) / 256, rows);
    copy2d_kernel<<<grid, 256, 0, s>>>(out, ldo, in, ldi, rows, cols);
}

void launch_add_diag(double* A, long lda, const double* d, int n, cudaStream_t s) {
    add_diag_kernel<<<(n + 255) / 256, 256, 0, s>>>(A, lda, d, n);
}

void launch_fill(double* p, long n, double v, cudaStream_t s) {
    if (n <= 0) return;
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    fill_kernel<<<(int)blocks, 256, 0, s>>>(p, n, v);
}

void launch_set_identity_pad(double* A, long lda, int n_valid, int n_pad, cudaStream_t s) {
    dim3 grid((n_pad + 255) / 256, n_pad);
    identity_pad_kernel<<<grid, 256, 0, s>>>(A, lda, n_valid, n_pad);
}

void launch_grad_reduce(const GradReduceParams& p, cudaStream_t s) {
    const int nt = (p.N + GT - 1) / GT;
    dim3 grid(nt, nt);
    grad_reduce_kernel<<<grid, 256, 0, s>>>(p);
    grad_final_kernel<<<p.nidx, 256, 0, s>>>(p.partials, nt * nt, p.nidx, p.out);
}

void launch_trace_and_sumsq(const double* A, long lda, const double* v, int n, double* out, cudaStream_t s) {
    trace_sumsq_kernel<<<1, 256, 0, s>>>(A, lda, v, n, out);
}
