// Prediction epilogues (gaussian_process.py:965-1006): mean = K*^T alpha, var = diag(K**) - |L^{-1} K*|^2
// computed per test point without ever forming the M* x M* covariance the reference builds.
// The K* tiles come from assemble.cu, the triangular solve is a sequence of DMMA GEMMs (api.cu);
// these kernels are the HBM-bound row reductions that finish the job.
#include "common.cuh"
#include "internal.h"

namespace {

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

}  // namespace

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
