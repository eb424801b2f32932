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
constexpr int LDB = NB + 1;
constexpr size_t POTRF_SMEM = ((size_t)NB * LDB + 3 * NB) * sizeof(double);

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, long lda, double* __restrict__ inv,
                                                            double* __restrict__ yk, double* __restrict__ logdet_part,
                                                            int* __restrict__ info, int row0) {
    extern __shared__ __align__(16) double sm[];
    double* Lb = sm;                 // NB x LDB; lower: L, strict upper: (L^{-1})^T
    double* dvec = sm + NB * LDB;    // pivots d_j, then sqrt(d_j)
    double* xdiag = dvec + NB;       // 1 / L_jj
    double* yv = xdiag + NB;
    __shared__ int s_info;
    const int tid = threadIdx.x;

    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        Lb[r * LDB + c] = (c <= r) ? A[(long)r * lda + c] : 0.0;
    }
    if (tid == 0) s_info = 0;
    __syncthreads();

    // right-looking elimination on the unscaled columns: a_ic -= a_ij a_cj / d_j  (i >= c > j)
    const int ta = tid >> 4, tb = tid & 15;
    for (int j = 0; j < NB - 1; j++) {
        double d = Lb[j * LDB + j];
        if (!(d > 0.0)) {
            if (tid == 0 && s_info == 0) s_info = row0 + j + 1;
            d = 1.0;
        }
        const double invd = 1.0 / d;
        for (int i = j + 1 + ta; i < NB; i += 16) {
            const double lij = Lb[i * LDB + j] * invd;
            for (int c = j + 1 + tb; c <= i; c += 16) Lb[i * LDB + c] -= lij * Lb[c * LDB + j];
        }
        __syncthreads();
    }
    if (tid < NB) {
        double d = Lb[tid * LDB + tid];
        if (!(d > 0.0)) {
            // pivots j < NB-1 were already flagged (in order) inside the loop; only the last is new here
            if (tid == NB - 1 && s_info == 0) s_info = row0 + NB;
            d = 1.0;
        }
        dvec[tid] = sqrt(d);
        xdiag[tid] = 1.0 / dvec[tid];
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if (c < r) Lb[r * LDB + c] *= xdiag[c];
        else if (c == r) Lb[r * LDB + c] = dvec[r];
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        A[(long)r * lda + c] = (c <= r) ? Lb[r * LDB + c] : 0.0;  // clean lower-triangular block (a GEMM operand later)
    }
    if (tid < 32) {
        double s = 0.0;
        for (int j = tid; j < NB; j += 32) s += log(dvec[j]);
        s = warp_sum(s);
        if (tid == 0) {
            *logdet_part = s;
            if (s_info != 0) atomicCAS(info, 0, s_info);
        }
    }

    // X = L^{-1}, column j owned by the lane pair (2j, 2j+1); X^T kept in the strict upper triangle.
    {
        const int j = tid >> 1, h = tid & 1;
        const double xjj = xdiag[j];
        for (int i = 1; i < NB; i++) {
            double s = 0.0;
            if (i > j) {
                for (int m = j + h; m < i; m += 2) {
                    const double x = (m == j) ? xjj : Lb[j * LDB + m];
                    s += Lb[i * LDB + m] * x;
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (i > j && h == 0) Lb[j * LDB + i] = -s * xdiag[i];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        inv[r * NB + c] = (c < r) ? Lb[c * LDB + r] : ((c == r) ? xdiag[r] : 0.0);
    }
    if (yk != nullptr) {
        if (tid < NB) yv[tid] = yk[tid];
        __syncthreads();
        if (tid < NB) {
            double s = xdiag[tid] * yv[tid];
            for (int c = 0; c < tid; c++) s += Lb[c * LDB + tid] * yv[c];
            yk[tid] = s;
        }
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
