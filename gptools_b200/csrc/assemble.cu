// Covariance assembly: K tiles generated from the kernels' derivative-order closed forms
// (covfn.cuh) straight into the padded, tile-aligned layout the blocked factorisation consumes.
// Replaces GaussianProcess.compute_Kij (gaussian_process.py:1535-1605), which materialises four
// (Mi*Mj, D) pair-list temporaries and calls the Python kernel on them.
//
// One CTA = one 64x64 output tile, 256 threads.  The 64 row points and 64 column points of the tile
// (coordinates + derivative orders) are staged once in shared memory; every thread then produces a
// 4x4 set of entries with columns interleaved by 16 so that each warp store covers two full
// 128-byte row segments.  Bound: HBM write (8 B per entry) / FP64 ALU for exp + polynomial.
#include "common.cuh"
#include "internal.h"

namespace {

constexpr int TS = 64;

__global__ void __launch_bounds__(256) assemble_kernel(AssembleParams p) {
    __shared__ double sxr[TS * GPT_MAX_DIM];
    __shared__ double sxc[TS * GPT_MAX_DIM];
    __shared__ int32_t snr[TS * GPT_MAX_DIM];
    __shared__ int32_t snc[TS * GPT_MAX_DIM];
    const int D = p.cp.D;
    const int r0 = blockIdx.y * TS, c0 = blockIdx.x * TS;
    const int tid = threadIdx.x;
    for (int i = tid; i < TS * D; i += 256) {
        const int r = i / D, d = i - r * D;
        const bool okr = (r0 + r) < p.Mr, okc = (c0 + r) < p.Mc;
        sxr[r * GPT_MAX_DIM + d] = okr ? p.Xr[(long)(r0 + r) * D + d] : 0.0;
        snr[r * GPT_MAX_DIM + d] = okr ? p.nr[(long)(r0 + r) * D + d] : 0;
        sxc[r * GPT_MAX_DIM + d] = okc ? p.Xc[(long)(c0 + r) * D + d] : 0.0;
        snc[r * GPT_MAX_DIM + d] = okc ? p.nc[(long)(c0 + r) * D + d] : 0;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll 1
    for (int a = 0; a < 4; a++) {
        const int lr = ty + 16 * a;
        const int r = r0 + lr;
        if (r >= p.rows_pad) continue;
#pragma unroll 1
        for (int b = 0; b < 4; b++) {
            const int lc = tx + 16 * b;
            const int c = c0 + lc;
            if (c >= p.cols_pad) continue;
            double v = 0.0;
            if (r < p.Mr && c < p.Mc) {
                if (p.swap_roles)
                    v = cov_eval(p.cp, sxc + lc * GPT_MAX_DIM, snc + lc * GPT_MAX_DIM, sxr + lr * GPT_MAX_DIM,
                                 snr + lr * GPT_MAX_DIM, p.hyper_deriv);
                else
                    v = cov_eval(p.cp, sxr + lr * GPT_MAX_DIM, snr + lr * GPT_MAX_DIM, sxc + lc * GPT_MAX_DIM,
                                 snc + lc * GPT_MAX_DIM, p.hyper_deriv);
                if (p.symmetric && r == c) {
                    v += p.diag_const;
                    if (p.diag_add) v += p.diag_add[r];
                }
            } else if (p.symmetric && p.pad_identity && r == c) {
                v = 1.0;
            }
            p.out[(long)r * p.ldo + c] = v;
        }
    }
}

__global__ void cov_pairs_kernel(CovParams cp, int hyper_deriv, long npairs, const double* __restrict__ Xi,
                                 const double* __restrict__ Xj, const int32_t* __restrict__ ni,
                                 const int32_t* __restrict__ nj, double* __restrict__ out) {
    const int D = cp.D;
    for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < npairs; q += (long)gridDim.x * blockDim.x) {
        double xi[GPT_MAX_DIM], xj[GPT_MAX_DIM];
        int32_t mi[GPT_MAX_DIM], mj[GPT_MAX_DIM];
        for (int d = 0; d < D; d++) {
            xi[d] = Xi[q * D + d];
            xj[d] = Xj[q * D + d];
            mi[d] = ni[q * D + d];
            mj[d] = nj[q * D + d];
        }
        out[q] = cov_eval(cp, xi, mi, xj, mj, hyper_deriv);
    }
}

}  // namespace

void launch_assemble(const AssembleParams& p, cudaStream_t s) {
    dim3 grid((p.cols_pad + TS - 1) / TS, (p.rows_pad + TS - 1) / TS);
    assemble_kernel<<<grid, 256, 0, s>>>(p);
}

void launch_cov_pairs(const CovParams& cp, int hyper_deriv, long npairs, const double* Xi, const double* Xj,
                      const int32_t* ni, const int32_t* nj, double* out, cudaStream_t s) {
    if (npairs <= 0) return;
    long blocks = (npairs + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cov_pairs_kernel<<<(int)blocks, 256, 0, s>>>(cp, hyper_deriv, npairs, Xi, Xj, ni, nj, out);
}
