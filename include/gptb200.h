/*
 * gptb200.h -- C ABI of libgptb200.so: the B200 (sm_100a) implementation of the gptools
 * GP likelihood / gradient / prediction hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI for this path --
 * it is numpy/scipy called from Python -- so each entry point cites the reference routine whose
 * numerical body it replaces (paths under /root/reference/gptools).  The Python host layer
 * (gptools_b200/) binds these with ctypes; INTEGRATION.md shows the stub a gptools maintainer
 * would add.
 *
 * Conventions
 *   - plain C, caller owns every host buffer; all arrays are C-contiguous;
 *     double = IEEE float64, derivative orders are int32.
 *   - every function returns 0 on success, a negative code on usage / CUDA errors
 *     (message via gpt_last_error), and LAPACK-style positive `info` values are reported
 *     through the `status` out-parameters (order of the first non-positive pivot).
 *   - a handle owns one CUDA device + stream and all device memory; it is not thread safe.
 *   - there is no CPU fallback: without a CUDA device gpt_create fails.
 */
#ifndef GPTB200_H
#define GPTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpt_handle gpt_handle;

/* kernel ids (gptools.kernel classes) */
#define GPT_SE 0          /* SquaredExponentialKernel  kernel/squared_exponential.py:31-174  params [sigma_f, l_1..l_D]      */
#define GPT_MATERN52 1    /* Matern52Kernel            kernel/matern.py:468-555 + src/matern.c  params [sigma_f, l_1..l_D]   */
#define GPT_MATERN 2      /* MaternKernel (nu = p+1/2) kernel/matern.py:251-465             params [sigma_f, nu, l_1..l_D]  */
#define GPT_GIBBS_TANH 3  /* GibbsKernel1dTanh         kernel/gibbs.py:244-505              params [sigma_f, l1, l2, lw, x0] */
#define GPT_GIBBS_AUX 4   /* GibbsKernel1d, any l_func kernel/gibbs.py:244-424, 508-1095      params [sigma_f]; the points carry
                           * three columns (x, l(x), l'(x)) -- the length-scale profile is evaluated by the host l_func,
                           * derivative orders apply to column 0 only */

#define GPT_COMPOSITE 5    /* SumKernel / ProductKernel trees over the kernels above  kernel/core.py:424-670
                           * params: the operands' parameter vectors back to back (BinaryKernel, kernel/core.py:452-459);
                           * structure from gpt_define_composite */

#define GPT_ERR_USAGE (-1)
#define GPT_ERR_CUDA (-2)
#define GPT_ERR_UNSUPPORTED (-3)

int gpt_version(void);
int gpt_device_count(void);

int gpt_create(int device, gpt_handle** out);
void gpt_destroy(gpt_handle* h);
const char* gpt_last_error(gpt_handle* h);
/* Run on an existing cudaStream_t (e.g. torch's current stream); NULL is the legacy default stream. */
int gpt_set_stream(gpt_handle* h, void* cuda_stream);
/* Go back to the handle's own (non-blocking) stream. */
int gpt_use_own_stream(gpt_handle* h);
int gpt_synchronize(gpt_handle* h);

/* Training set exactly as GaussianProcess.add_data leaves it (gaussian_process.py:376-503):
 * X (N x D), n (N x D), y (M) with the mean function already subtracted (gaussian_process.py:1455-1461),
 * err_y (M), T (M x N) or NULL (then N == M). */
int gpt_set_data(gpt_handle* h, int N, int M, int D, const double* X, const int32_t* n, const double* y,
                 const double* err_y, const double* T);
/* Replace y only (mean-function parameters changed). */
int gpt_set_y(gpt_handle* h, const double* y);
/* Covariance kernel of the GP and the diag_factor jitter (gaussian_process.py:81-85, 1450). */
int gpt_set_kernel(gpt_handle* h, int kernel_id, int nparams, double diag_factor);

/* Kernel algebra on the device (SumKernel kernel/core.py:549-600, ProductKernel kernel/core.py:601-670).  The host flattens
 * a tree of sums and products into  k = sum_t prod_{q in term t} leaf_q : `nleaf` operand kernels (SE, Matern-5/2, Matern,
 * Gibbs-tanh; at most 4, at most 10 parameters in total) and `nterms` products (at most 8), term t multiplying the leaves
 * whose bits are set in term_masks[t].  Products of kernels under derivative observations follow the general Leibniz
 * rule over the derivative orders exactly like the reference's enumeration of derivative subsets (kernel/core.py:632-668).
 * Afterwards kernel id GPT_COMPOSITE (nparams = sum of leaf_nparams) is accepted by gpt_set_kernel, gpt_cov_pairs and
 * gpt_compute_Kij on this handle; hyper_deriv / grad_idx index the concatenated parameter vector (the reference has
 * hyper-derivatives for sums only).  gpt_ll_batched runs composite kernels in its persistent many-theta kernel (the leaves
 * of a theta live in the CTA's workspace); gradient requests beyond parameter index 6 run theta after theta instead. */
int gpt_define_composite(gpt_handle* h, int nleaf, const int32_t* leaf_kernel_ids, const int32_t* leaf_nparams,
                         int nterms, const int32_t* term_masks);

/* Kernel.__call__ on flattened pair lists (kernel/core.py:220-257): out[p] = k(Xi[p], Xj[p]; ni[p], nj[p]).
 * hyper_deriv = -1 for the value, else the index into params (every kernel; d/dnu of
 * the generic Matern kernel is GPT_ERR_UNSUPPORTED). Independent of set_data. */
int gpt_cov_pairs(gpt_handle* h, int kernel_id, int D, int nparams, const double* params, int hyper_deriv,
                  int64_t npairs, const double* Xi, const double* Xj, const int32_t* ni, const int32_t* nj,
                  double* out);

/* GaussianProcess.compute_Kij (gaussian_process.py:1535-1605): K_out (Mi x Mj) = k(Xi_i, Xj_j).
 * Xj == NULL means the symmetric case (Xj = Xi). */
int gpt_compute_Kij(gpt_handle* h, int kernel_id, int D, int nparams, const double* params, int hyper_deriv,
                    int Mi, const double* Xi, const int32_t* ni, int Mj, const double* Xj, const int32_t* nj,
                    double* K_out);

/* compute_K_L_alpha_ll (gaussian_process.py:1418-1522) for one hyper-parameter vector.
 *   params       kernel parameters (nparams) in the reference's order
 *   noise_sigma  DiagonalNoiseKernel sigma_n (0 for ZeroKernel), gaussian_process.py:1434-1439
 *   ll           out: -1/2 y'alpha - sum log L_ii - M/2 log 2pi   (hyperprior is added by the host)
 *   grad         out (P) or NULL: 1/2 (alpha' dK alpha - tr(K^-1 dK)) for each entry of grad_idx
 *   grad_idx     P indices into params; the value `nparams` selects sigma_n
 *   status       out: 0, or potrf info > 0 (matrix not positive definite)
 * Leaves L, alpha and the block inverses resident for gpt_predict / gpt_get_*. */
int gpt_ll(gpt_handle* h, const double* params, double noise_sigma, double* ll, double* grad,
           const int32_t* grad_idx, int P, int* status);
/* Same with the latent covariance K + K_noise (N x N) supplied by the host (user-defined Python kernels). */
int gpt_ll_from_K(gpt_handle* h, const double* K_latent, double* ll, int* status);
/* 1/2 (alpha' T dK T' alpha - tr(K_tot^-1 T dK T')) for a host-supplied dK (N x N); after gpt_ll / gpt_ll_from_K. */
int gpt_grad_from_dK(gpt_handle* h, const double* dK_latent, double* g);

/* d ll / d sigma_n of a DiagonalNoiseKernel for host-evaluated kernels: 1/2 tr((alpha alpha^T - K_tot^{-1}) 2 sigma_n I_M)
 * = sigma_n (alpha^T alpha - tr K_tot^{-1}), the identity taken over the M observations also when T is present
 * (gaussian_process.py:1484-1488).  K_tot^{-1} comes from the resident factor (DMMA GEMMs), trace and alpha^T alpha are
 * reduced on the device. */
int gpt_noise_grad(gpt_handle* h, double noise_sigma, double* g);

int gpt_get_alpha(gpt_handle* h, double* alpha /* M */);
int gpt_get_L(gpt_handle* h, double* L /* M x M, lower, zeros above */);
int gpt_get_K(gpt_handle* h, double* K /* N x N latent covariance of the last gpt_ll, without noise */);

/* Many hyper-parameter vectors at once (emcee walkers, optimizer restarts, ll grids):
 * the unit of work of update_hyperparameters (gaussian_process.py:1332-1416), batched.
 *   thetas   B x (nparams + 1): kernel params then sigma_n
 *   y_batch  B x M or NULL (per-theta mean-subtracted targets)
 *   ll (B), grad (B x P) or NULL, status (B), alpha_out (B x M) or NULL
 * Up to 2048 observations without a transformation matrix: ONE launch of the persistent many-theta kernel (one CTA
 * per theta, four per SM); gradients of kernels other than SE are limited to parameter
 * indices < 7 there.  With a transformation matrix T, or beyond 2048 observations, the same call runs the thetas back
 * to back through the single-matrix path (assembly, T K T^T, blocked Cholesky, solves, inverse + trace reduction) on
 * the handle's stream: every scalar stays on the device, nothing synchronises between thetas, one copy at the end.
 * GPT_GIBBS_AUX (per-point columns that depend on theta) is not batched. */
int gpt_ll_batched(gpt_handle* h, int B, const double* thetas, const double* y_batch, double* ll, double* grad,
                   const int32_t* grad_idx, int P, int* status, double* alpha_out);
/* Same, with thetas / outputs already resident in device memory (raw device pointers). */
int gpt_ll_batched_dev(gpt_handle* h, int B, const double* d_thetas, const double* d_y_batch, double* d_ll,
                       double* d_grad, const int32_t* grad_idx, int P, int* d_status, double* d_alpha_out);

/* Predictions at MANY hyper-parameter vectors: the loop of compute_from_MCMC / predict_MCMC over the retained samples
 * (gaussian_process.py:1944-1969: update_hyperparameters + predict per sample, farmed to a process pool) as ONE launch of
 * the persistent many-theta kernel.  The Ms test points ride along as extra tile rows of the factorisation: the panel
 * tiles of a test row block are the rows of (L^-1 K*)^T, so mean = (L^-1 K*)^T z and var = diag(K**) - |L^-1 K*|^2
 * accumulate in the epilogues with no second pass (predict core, gaussian_process.py:965-1006).
 *   thetas (B x (nparams + 1)), y_batch (B x M or NULL) as in gpt_ll_batched; Xs (Ms x D), ns (Ms x D)
 *   mean, var (B x Ms); ll (B) or NULL; status (B): potrf info per theta (rows with status != 0 are undefined)
 * Needs the persistent kernel: no transformation matrix, at most 2048 observations (else GPT_ERR_UNSUPPORTED and the
 * caller predicts theta by theta). */
int gpt_predict_batched(gpt_handle* h, int B, const double* thetas, const double* y_batch, int Ms, const double* Xs,
                        const int32_t* ns, double* mean, double* var, double* ll, int* status);

/* predict numeric core (gaussian_process.py:965-1006) with the state of the last gpt_ll:
 *   mean (Ms); var (Ms) or NULL: diag(K**) - |L^-1 K*|^2 ; cov (Ms x Ms) or NULL: K** - v'v */
int gpt_predict(gpt_handle* h, int Ms, const double* Xs, const int32_t* ns, double* mean, double* var, double* cov);

/* gpt_predict with DEVICE output buffers (mean, and variance when d_var != NULL; Xs / ns are host arrays): the results
 * are copied device to device on the handle's stream and the call does not synchronise -- the multi-GPU host layer
 * points them at the send buffer of its all-gather (gptools_b200/parallel.py: predict_sharded). */
int gpt_predict_dev(gpt_handle* h, int Ms, const double* Xs, const int32_t* ns, double* d_mean, double* d_var);

/* The same numeric core for kernels evaluated on the host (user-defined Python kernels, kernel sums): the caller
 * supplies the cross-covariance exactly as gaussian_process.py:966 builds it, transposed --
 * KstarT[s][i] = k(X_i, Xstar_s; n_i, nstar_s), Ms x N row-major over the LATENT points -- and the prior
 * (co)variance of the test points: kss_diag (Ms) for var, Kss (Ms x Ms) for cov (gaussian_process.py:984).
 * Uses the factorisation left by gpt_ll or gpt_ll_from_K; T is applied on the device. */
int gpt_predict_from_Kstar(gpt_handle* h, int Ms, const double* KstarT, const double* kss_diag, const double* Kss,
                           double* mean, double* var, double* cov);

/* draw_sample, method='cholesky' with explicit rand_vars (gaussian_process.py:1295-1300, 1330):
 * out (Ms x S) = mean + chol(cov + jitter I) * rand_vars (Ms x S). */
int gpt_draw_sample(gpt_handle* h, int Ms, int S, const double* mean, const double* cov, const double* rand_vars,
                    double jitter, double* out, int* status);

/* Diagnostics for bench.py: achieved time of the last batched launch is measured by the caller with
 * CUDA events on the handle's stream (gpt_set_stream). Number of kernel launches issued so far. */
int64_t gpt_launch_count(gpt_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* GPTB200_H */
