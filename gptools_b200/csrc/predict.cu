// Prediction epilogues (gaussian_process.py:965-1006): mean = K*^T alpha, var = diag(K**) - |L^{-1} K*|^2
// computed per test point without ever forming the M* x M* covariance the reference builds.
// The K* tiles come from assemble.cu, the triangular solve is a sequence of DMMA GEMMs (api.cu);
// these kernels are the HBM-bound row reductions that finish the job.
#include "common.cuh"
#include "internal.h"
#include "se_fast.cuh"

namespace {

using namespace sefast;

__global__ void __launch_bounds__(256) rowdot_kernel(const double* __restrict__ Kst, long ld, int rows, int n,
                                                     const double* __restrict__ alpha, double* __restrict__ mean) {
    const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* r = Kst + (long)row * ld;
    double s = 0.0;
    for (int c = lane; c < n; c += 32) s += r[c] * alpha[c];
    s = warp_sum(s);
    if (lane == 0) mean[row] = s;
}

__global__ void __launch_bounds__(256) row_var_kernel(const double* __restrict__ V, long ld, int rows, int n,
                                                      const double* __restrict__ kss, double* __restrict__ var) {
    const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* r = V + (long)row * ld;
    double s = 0.0;
    for (int c = lane; c < n; c += 32) s += r[c] * r[c];
    s = warp_sum(s);
    if (lane == 0) var[row] = kss[row] - s;
}

__global__ void prior_diag_kernel(CovParams cp, const double* __restrict__ Xs, const int32_t* __restrict__ ns,
                                  int rows, double* __restrict__ kss) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double x[GPT_MAX_DIM];
    int32_t m[GPT_MAX_DIM];
    for (int d = 0; d < cp.D; d++) {
        x[d] = Xs[(long)r * cp.D + d];
        m[d] = ns[(long)r * cp.D + d];
    }
    kss[r] = cov_eval(cp, x, m, x, m, -1);
}

// ---- fused predictive mean: mean[s] = sum_i k(X_i, x*_s; n_i, n*_s) u_i, K* generated on the fly and never stored ----
// One thread per test point, training points (coordinates, orders, weight u = alpha or T^T alpha) staged through
// shared memory 128 at a time and read as broadcasts; grid.y splits the training set so that small test sets still
// fill the machine (deterministic two-pass reduction).  Bound: FP64 ALU (exp + polynomial per pair); the
// "equivalent bytes" 8 N M* of the K* the reference materialises are never moved.
constexpr int PM_THREADS = 128;
constexpr int PM_STAGE = 128;

template <int FD, int KIND = 0>
__global__ void __launch_bounds__(PM_THREADS) predict_mean_kernel(CovParams cp, const double* __restrict__ X,
                                                                  const int32_t* __restrict__ n,
                                                                  const double* __restrict__ u, int N,
                                                                  const double* __restrict__ Xs,
                                                                  const int32_t* __restrict__ ns, int Ms, int per_split,
                                                                  int low_order, double* __restrict__ partial) {
    __shared__ double sx[PM_STAGE * GPT_MAX_DIM];
    __shared__ int32_t sn[PM_STAGE * GPT_MAX_DIM];
    __shared__ double su[PM_STAGE];
    __shared__ double etab[64];  // 2^(j/64) for exp_nonpos_tab
    if (threadIdx.x < 64) etab[threadIdx.x] = GPT_EXP2_64[threadIdx.x];  // published by the first barrier below
    const int D = cp.D;
    const int s = blockIdx.x * PM_THREADS + threadIdx.x;
    const bool live = s < Ms;
    const int i0 = blockIdx.y * per_split;
    const int i1 = min(N, i0 + per_split);
    double xs[GPT_MAX_DIM];
    int32_t ms[GPT_MAX_DIM];
#pragma unroll
    for (int d = 0; d < GPT_MAX_DIM; d++) {
        xs[d] = (live && d < D) ? Xs[(size_t)s * D + d] : 0.0;
        ms[d] = (live && d < D) ? ns[(size_t)s * D + d] : 0;
    }
    double acc = 0.0;
    for (int base = i0; base < i1; base += PM_STAGE) {
        const int cnt = min(PM_STAGE, i1 - base);
        __syncthreads();
        for (int idx = threadIdx.x; idx < PM_STAGE * D; idx += PM_THREADS) {
            const int r = idx / D, d = idx - r * D;
            const bool ok = r < cnt;
            sx[r * GPT_MAX_DIM + d] = ok ? X[(size_t)(base + r) * D + d] : 0.0;
            sn[r * GPT_MAX_DIM + d] = ok ? n[(size_t)(base + r) * D + d] : 0;
        }
        if (threadIdx.x < PM_STAGE) su[threadIdx.x] = (threadIdx.x < cnt) ? u[base + threadIdx.x] : 0.0;
        __syncthreads();
        if constexpr (FD == 0) {
#pragma unroll 1
            for (int r = 0; r < cnt; r++)
                acc += cov_eval(cp, sx + r * GPT_MAX_DIM, sn + r * GPT_MAX_DIM, xs, ms, -1) * su[r];
        } else if constexpr (KIND == 1) {
            // Matern 5/2 (points with at most one first derivative): branch-free, all 128 staged rows (u = 0 beyond cnt)
            const M52Hoist<FD> hm = m52_hoist<FD>(cp, etab);
            PointReg<FD> pj;
#pragma unroll
            for (int d = 0; d < FD; d++) {
                pj.x[d] = xs[d];
                pj.n[d] = ms[d];
            }
#pragma unroll 8
            for (int r = 0; r < PM_STAGE; r++) {
                PointReg<FD> pi;
#pragma unroll
                for (int d = 0; d < FD; d++) {
                    pi.x[d] = sx[r * GPT_MAX_DIM + d];
                    pi.n[d] = sn[r * GPT_MAX_DIM + d];
                }
                acc = fma(m52_value_low<FD>(hm, pi, pj), su[r], acc);
            }
        } else {
            SEHoist<FD> h = se_hoist<FD>(cp);
            h.etab = etab;
            PointReg<FD> pj;
#pragma unroll
            for (int d = 0; d < FD; d++) {
                pj.x[d] = xs[d];
                pj.n[d] = ms[d];
            }
            if (low_order) {
                // staged rows beyond cnt carry u = 0 and finite coordinates: evaluate all 128 branch-free
#pragma unroll 8
                for (int r = 0; r < PM_STAGE; r++) {
                    PointReg<FD> pi;
#pragma unroll
                    for (int d = 0; d < FD; d++) {
                        pi.x[d] = sx[r * GPT_MAX_DIM + d];
                        pi.n[d] = sn[r * GPT_MAX_DIM + d];
                    }
                    acc = fma(se_value_low<FD, true>(h, pi, pj), su[r], acc);
                }
            } else {
#pragma unroll 2
                for (int r = 0; r < cnt; r++) {
                    PointReg<FD> pi;
#pragma unroll
                    for (int d = 0; d < FD; d++) {
                        pi.x[d] = sx[r * GPT_MAX_DIM + d];
                        pi.n[d] = sn[r * GPT_MAX_DIM + d];
                    }
                    acc += se_value<FD>(h, pi, pj) * su[r];
                }
            }
        }
    }
    if (live) partial[(size_t)blockIdx.y * Ms + s] = acc;
}

__global__ void sum_partials_kernel(const double* __restrict__ partial, int nsplit, int Ms, double* __restrict__ mean) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Ms) return;
    double a = 0.0;
    for (int k = 0; k < nsplit; k++) a += partial[(size_t)k * Ms + s];  // fixed order: deterministic
    mean[s] = a;
}

}  // namespace

int predict_mean_nsplit(int N, int Ms) {
    const int bx = (Ms + PM_THREADS - 1) / PM_THREADS;
    int nsplit = (4 * 148 + bx - 1) / bx;
    const int max_split = (N + 4 * PM_STAGE - 1) / (4 * PM_STAGE);
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    return nsplit;
}

void launch_predict_mean_fused(const CovParams& cp, const double* X, const int32_t* n, const double* u, int N,
                               const double* Xs, const int32_t* ns, int Ms, int low_order, double* partial,
                               double* mean, cudaStream_t s) {
    if (Ms <= 0) return;
    const int nsplit = predict_mean_nsplit(N, Ms);
    int per_split = (N + nsplit - 1) / nsplit;
    per_split = (per_split + PM_STAGE - 1) / PM_STAGE * PM_STAGE;
    dim3 grid((Ms + PM_THREADS - 1) / PM_THREADS, nsplit);
    if (cp.kid == GPT_KERNEL_SE && cp.D == 1)
        predict_mean_kernel<1><<<grid, PM_THREADS, 0, s>>>(cp, X, n, u, N, Xs, ns, Ms, per_split, low_order, partial);
    else if (cp.kid == GPT_KERNEL_SE && cp.D == 2)
        predict_mean_kernel<2><<<grid, PM_THREADS, 0, s>>>(cp, X, n, u, N, Xs, ns, Ms, per_split, low_order, partial);
    else if (cp.kid == GPT_KERNEL_SE && cp.D == 3)
        predict_mean_kernel<3><<<grid, PM_THREADS, 0, s>>>(cp, X, n, u, N, Xs, ns, Ms, per_split, low_order, partial);
    else if (cp.kid == GPT_KERNEL_MATERN52 && cp.D == 1 && low_order)
        predict_mean_kernel<1, 1><<<grid, PM_THREADS, 0, s>>>(cp, X, n, u, N, Xs, ns, Ms, per_split, low_order, partial);
    else if (cp.kid == GPT_KERNEL_MATERN52 && cp.D == 2 && low_order)
        predict_mean_kernel<2, 1><<<grid, PM_THREADS, 0, s>>>(cp, X, n, u, N, Xs, ns, Ms, per_split, low_order, partial);
    else
        predict_mean_kernel<0><<<grid, PM_THREADS, 0, s>>>(cp, X, n, u, N, Xs, ns, Ms, per_split, low_order, partial);
    sum_partials_kernel<<<(Ms + 255) / 256, 256, 0, s>>>(partial, nsplit, Ms, mean);
}

void launch_rowdot(const double* Kst, long ld, int rows, int n, const double* alpha, double* mean, cudaStream_t s) {
    if (rows <= 0) return;
    rowdot_kernel<<<(rows + 7) / 8, 256, 0, s>>>(Kst, ld, rows, n, alpha, mean);
}

void launch_row_var(const double* V, long ld, int rows, int n, const double* kss, double* var, cudaStream_t s) {
    if (rows <= 0) return;
    row_var_kernel<<<(rows + 7) / 8, 256, 0, s>>>(V, ld, rows, n, kss, var);
}

void launch_prior_diag(const CovParams& cp, const double* Xs, const int32_t* ns, int rows, double* kss,
                       cudaStream_t s) {
    if (rows <= 0) return;
    prior_diag_kernel<<<(rows + 127) / 128, 128, 0, s>>>(cp, Xs, ns, rows, kss);
}
