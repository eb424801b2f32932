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
#include "se_fast.cuh"

namespace {

constexpr int TS = 64;

// Squared-exponential value tiles, input dimension FD <= 3, derivative orders <= 1 on both sides: register-resident
// branch-free closed forms with the table-based exp (se_fast.cuh) instead of the generic cov_eval -- the generic
// path keeps per-dimension arrays in local memory and branches per entry.
// KIND 0: squared exponential; KIND 1: Matern 5/2 (m52_value_low)
template <int FD, int KIND>
__device__ __forceinline__ void assemble_tile_se_low(const AssembleParams& p, const double* sxr, const int32_t* snr,
                                                     const double* sxc, const int32_t* snc, const double* etab,
                                                     int r0, int c0, int tid) {
    using namespace sefast;
    SEHoist<FD> h = se_hoist<FD>(p.cp);
    h.etab = etab;
    const M52Hoist<FD> hm = m52_hoist<FD>(p.cp, etab);
    const int ty = tid >> 4, tx = tid & 15;
    PointReg<FD> pc[4];
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int d = 0; d < FD; d++) {
            pc[b].x[d] = sxc[(tx + 16 * b) * GPT_MAX_DIM + d];
            pc[b].n[d] = snc[(tx + 16 * b) * GPT_MAX_DIM + d];
        }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int lr = ty + 16 * a;
        const int r = r0 + lr;
        PointReg<FD> pr;
#pragma unroll
        for (int d = 0; d < FD; d++) {
            pr.x[d] = sxr[lr * GPT_MAX_DIM + d];
            pr.n[d] = snr[lr * GPT_MAX_DIM + d];
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int c = c0 + tx + 16 * b;
            // out[r][c] = k(row_r, col_c), or k(col_c, row_r) with swapped roles (the sign depends on which side
            // carries the odd derivative order)
            double v;
            if constexpr (KIND == 1) v = p.swap_roles ? m52_value_low<FD>(hm, pc[b], pr) : m52_value_low<FD>(hm, pr, pc[b]);
            else v = p.swap_roles ? se_value_low<FD, true>(h, pc[b], pr) : se_value_low<FD, true>(h, pr, pc[b]);
            const bool inside = r < p.Mr && c < p.Mc;
            v = inside ? v : 0.0;
            if (p.symmetric && r == c) {
                if (inside) v += p.diag_const + (p.diag_add ? p.diag_add[r] : 0.0);
                else if (p.pad_identity) v = 1.0;
            }
            if (r < p.rows_pad && c < p.cols_pad) p.out[(long)r * p.ldo + c] = v;
        }
    }
}

// Stage the 64 row points and 64 column points of the tile (coordinates + derivative orders) and the exp table.
__device__ __forceinline__ void stage_tile_points(const AssembleParams& p, double* sxr, double* sxc, int32_t* snr,
                                                  int32_t* snc, double* etab, int r0, int c0, int tid) {
    const int D = p.cp.D;
    for (int i = tid; i < TS * D; i += 256) {
        const int r = i / D, d = i - r * D;
        const bool okr = (r0 + r) < p.Mr, okc = (c0 + r) < p.Mc;
        sxr[r * GPT_MAX_DIM + d] = okr ? p.Xr[(long)(r0 + r) * D + d] : 0.0;
        snr[r * GPT_MAX_DIM + d] = okr ? p.nr[(long)(r0 + r) * D + d] : 0;
        sxc[r * GPT_MAX_DIM + d] = okc ? p.Xc[(long)(c0 + r) * D + d] : 0.0;
        snc[r * GPT_MAX_DIM + d] = okc ? p.nc[(long)(c0 + r) * D + d] : 0;
    }
    if (tid < 64) etab[tid] = GPT_EXP2_64[tid];
}

// The SE / orders <= 1 value tiles have their OWN kernel: inside the generic kernel they inherited its 255 registers
// (cov_eval covers Matern with K_nu of real order, Gibbs, Hermite recurrences ...) and ran at one CTA per SM --
// 12% occupancy, 14% of the FP64 pipe, 250 GB/s of output (profiles/r02j_*).  Three CTAs per SM here.
template <int FD, int KIND>
__global__ void __launch_bounds__(256, 3) assemble_se_low_kernel(AssembleParams p) {
    if (p.lower_tiles_only && blockIdx.x > blockIdx.y) return;  // the Cholesky never reads tiles above the diagonal
    __shared__ double sxr[TS * GPT_MAX_DIM];
    __shared__ double sxc[TS * GPT_MAX_DIM];
    __shared__ int32_t snr[TS * GPT_MAX_DIM];
    __shared__ int32_t snc[TS * GPT_MAX_DIM];
    __shared__ double etab[64];
    const int r0 = blockIdx.y * TS, c0 = blockIdx.x * TS;
    stage_tile_points(p, sxr, sxc, snr, snc, etab, r0, c0, threadIdx.x);
    __syncthreads();
    assemble_tile_se_low<FD, KIND>(p, sxr, snr, sxc, snc, etab, r0, c0, threadIdx.x);
}

// Gibbs kernel in one dimension (kernel/gibbs.py:324-423; gibbs_cov_l in covfn.cuh), derivative orders <= 1: the length
// scale l(x) and l'(x) of the 64 row and 64 column points are evaluated ONCE per tile (tanh profile: kernel/gibbs.py:458-461;
// AUX: they arrive as point columns 1 and 2), the entries then cost one reciprocal, one square root and one table exp each.
// Inside the generic kernel every entry paid two tanh, a division, sqrt and the library exp at one CTA per SM: the
// latent covariance of config 5 (4001^2 entries) was most of its 1.7 ms log-likelihood.
template <bool AUX>
__global__ void __launch_bounds__(256, 3) assemble_gibbs_low_kernel(AssembleParams p) {
    if (p.lower_tiles_only && blockIdx.x > blockIdx.y) return;
    __shared__ double sr[TS][3], sc[TS][3];  // x, l(x), l'(x)
    __shared__ int32_t nr[TS], nc[TS];
    __shared__ double etab[64];
    const int r0 = blockIdx.y * TS, c0 = blockIdx.x * TS;
    const int tid = threadIdx.x;
    const int D = p.cp.D;  // 1, or 3 with the auxiliary columns
    if (tid < 2 * TS) {
        const bool row = tid < TS;
        const int l = row ? tid : tid - TS;
        const int gi = (row ? r0 : c0) + l;
        const bool ok = gi < (row ? p.Mr : p.Mc);
        const double* X = row ? p.Xr : p.Xc;
        const int32_t* n = row ? p.nr : p.nc;
        double x = 0.0, lx = 1.0, lx1 = 0.0;
        if (ok) {
            x = X[(long)gi * D];
            if (AUX) {
                lx = X[(long)gi * D + 1];
                lx1 = X[(long)gi * D + 2];
            } else {
                gibbs_tanh_l(p.cp, x, lx, lx1);
            }
        }
        double* dst = row ? sr[l] : sc[l];
        dst[0] = x;
        dst[1] = lx;
        dst[2] = lx1;
        (row ? nr : nc)[l] = ok ? n[(long)gi * D] : 0;
    }
    if (tid < 64) etab[tid] = GPT_EXP2_64[tid];
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int lr = ty + 16 * a;
        const int r = r0 + lr;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int lc = tx + 16 * b;
            const int c = c0 + lc;
            // out[r][c] = k(row_r, col_c), or k(col_c, row_r) with swapped roles
            const double* pa = p.swap_roles ? sc[lc] : sr[lr];
            const double* pb = p.swap_roles ? sr[lr] : sc[lc];
            const int na = p.swap_roles ? nc[lc] : nr[lr], nb = p.swap_roles ? nr[lr] : nc[lc];
            const double lx = pa[1], lx1 = pa[2], ly = pb[1], ly1 = pb[2];
            const double d = pa[0] - pb[0];
            const double S = lx * lx + ly * ly;
            const double iS = (S > 0.0) ? fast_rcp_pos(S) : 1.0 / S;
            const double z = 2.0 * lx * ly * iS;
            const double k00 = ((z > 0.0) ? z * fast_rsqrt_pos(z) : sqrt(z)) * exp_nonpos_tab(-d * d * iS, etab);
            const double Ax = lx1 / (2.0 * lx) - lx * lx1 * iS - 2.0 * d * iS + 2.0 * d * d * lx * lx1 * iS * iS;
            const double Ay = ly1 / (2.0 * ly) - ly * ly1 * iS + 2.0 * d * iS + 2.0 * d * d * ly * ly1 * iS * iS;
            const double dAx = 2.0 * lx * lx1 * ly * ly1 * iS * iS + 2.0 * iS + 4.0 * d * ly * ly1 * iS * iS -
                               4.0 * d * lx * lx1 * iS * iS - 8.0 * d * d * lx * lx1 * ly * ly1 * iS * iS * iS;
            double f = 1.0;
            f = (na && !nb) ? Ax : f;
            f = (!na && nb) ? Ay : f;
            f = (na && nb) ? fma(Ax, Ay, dAx) : f;
            double v = p.cp.sig2 * k00 * f;
            const bool inside = r < p.Mr && c < p.Mc;
            v = inside ? v : 0.0;
            if (p.symmetric && r == c) {
                if (inside) v += p.diag_const + (p.diag_add ? p.diag_add[r] : 0.0);
                else if (p.pad_identity) v = 1.0;
            }
            if (r < p.rows_pad && c < p.cols_pad) p.out[(long)r * p.ldo + c] = v;
        }
    }
}

__global__ void __launch_bounds__(256) assemble_kernel(AssembleParams p) {
    if (p.lower_tiles_only && blockIdx.x > blockIdx.y) return;  // the Cholesky never reads tiles above the diagonal
    __shared__ double sxr[TS * GPT_MAX_DIM];
    __shared__ double sxc[TS * GPT_MAX_DIM];
    __shared__ int32_t snr[TS * GPT_MAX_DIM];
    __shared__ int32_t snc[TS * GPT_MAX_DIM];
    __shared__ double etab[64];
    const int r0 = blockIdx.y * TS, c0 = blockIdx.x * TS;
    const int tid = threadIdx.x;
    stage_tile_points(p, sxr, sxc, snr, snc, etab, r0, c0, tid);
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
    if (p.cp.kid == GPT_KERNEL_SE && p.hyper_deriv < 0 && p.low_order && p.cp.D <= 3) {
        if (p.cp.D == 1) assemble_se_low_kernel<1, 0><<<grid, 256, 0, s>>>(p);
        else if (p.cp.D == 2) assemble_se_low_kernel<2, 0><<<grid, 256, 0, s>>>(p);
        else assemble_se_low_kernel<3, 0><<<grid, 256, 0, s>>>(p);
        return;
    }
    // Matern 5/2 value tiles (its points carry at most one first derivative: kernel/matern.py:545-546)
    if (p.cp.kid == GPT_KERNEL_MATERN52 && p.hyper_deriv < 0 && p.low_order && p.cp.D <= 3) {
        if (p.cp.D == 1) assemble_se_low_kernel<1, 1><<<grid, 256, 0, s>>>(p);
        else if (p.cp.D == 2) assemble_se_low_kernel<2, 1><<<grid, 256, 0, s>>>(p);
        else assemble_se_low_kernel<3, 1><<<grid, 256, 0, s>>>(p);
        return;
    }
    if (p.hyper_deriv < 0 && p.low_order && p.cp.kid == GPT_KERNEL_GIBBS_TANH && p.cp.D == 1) {
        assemble_gibbs_low_kernel<false><<<grid, 256, 0, s>>>(p);
        return;
    }
    if (p.hyper_deriv < 0 && p.low_order && p.cp.kid == GPT_KERNEL_GIBBS_AUX && p.cp.D == 3) {
        assemble_gibbs_low_kernel<true><<<grid, 256, 0, s>>>(p);
        return;
    }
    assemble_kernel<<<grid, 256, 0, s>>>(p);
}

void launch_cov_pairs(const CovParams& cp, int hyper_deriv, long npairs, const double* Xi, const double* Xj,
                      const int32_t* ni, const int32_t* nj, double* out, cudaStream_t s) {
    if (npairs <= 0) return;
    long blocks = (npairs + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cov_pairs_kernel<<<(int)blocks, 256, 0, s>>>(cp, hyper_deriv, npairs, Xi, Xj, ni, nj, out);
}
