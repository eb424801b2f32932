// Internal (non-ABI) declarations shared by the .cu translation units of libgptb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "covfn.cuh"

#define GPT_NB 128  // block size of the blocked single-matrix algorithms (= GEMM tile edge)

// ---- gemm.cu : FP64 DMMA GEMM, C = beta*C + alpha * A * B^T -------------------------------
// A: (tiles_m*128) x K, B: (tiles_n*128) x K, C: (tiles_m*128) x (tiles_n*128); all row-major,
// K a multiple of 16, leading dimensions even, base pointers 16-byte aligned.
struct GemmParams {
    double* C;
    long ldc;
    const double* A;
    long lda;
    const double* B;
    long ldb;
    int tiles_m, tiles_n;
    int K;
    double alpha, beta;
    int lower_only;  // only tiles with tile_n <= tile_m (symmetric rank-k updates, lauum)
    int kbegin_row;  // contraction starts at k = tile_m*128 (A upper block-triangular)
    int kend_row = 0;   // contraction ends at k = (tile_m+1)*128 (A lower block-triangular)
    int batch = 1;      // independent products in one launch: operand b lives at base + b * stride
    long strideA = 0, strideB = 0, strideC = 0;
};
void launch_gemm_nt(const GemmParams& p, cudaStream_t s);
long gemm_rows_per_wave_n128(int num_sms);

// ---- assemble.cu : covariance tile generation ------------------------------------------------
struct AssembleParams {
    CovParams cp;
    const double* Xr;   // row points  (Mr x D)
    const int32_t* nr;  // row derivative orders
    int Mr;
    const double* Xc;   // column points (Mc x D)
    const int32_t* nc;
    int Mc;
    double* out;        // (rows_pad x cols_pad) row-major
    long ldo;
    int rows_pad, cols_pad;
    int hyper_deriv;    // -1: value
    int swap_roles;     // 0: out[r][c] = k(row_r, col_c); 1: out[r][c] = k(col_c, row_r)
    int symmetric;      // rows and cols are the same set: add diag terms, identity padding
    const double* diag_add;  // optional per-row additive diagonal (err_y^2 + jitter [+ noise]) or NULL
    double diag_const;       // added to every diagonal element r < Mr (sigma_n^2 for the latent K)
    int pad_identity;        // symmetric: out[r][r] = 1 for r >= Mr
    int lower_tiles_only = 0;  // symmetric output consumed by the Cholesky only: skip tiles above the diagonal
    int low_order = 0;         // every derivative order <= 1 (both point sets): branch-free SE closed forms
};
void launch_assemble(const AssembleParams& p, cudaStream_t s);
void launch_cov_pairs(const CovParams& cp, int hyper_deriv, long npairs, const double* Xi, const double* Xj,
                      const int32_t* ni, const int32_t* nj, double* out, cudaStream_t s);

// ---- factor.cu : blocked Cholesky pieces, solves, reductions -----------------------------------
// Factor one 128x128 diagonal block in place (lower), write its inverse (lower, zero above) to
// inv, optionally z_k = inv * y_k (in place on y), accumulate sum(log diag) and the LAPACK-style info.
// Pprev (128 x pcols, leading dimension ldp, pcols = 128 or 256) or NULL: first apply the pending update
// Ablk -= Pprev Pprev^T.
void launch_potrf_diag(double* Ablk, long lda, double* inv, double* yk, double* logdet_part, int* info,
                       int row0, const double* Pprev, long ldp, int pcols, cudaStream_t s);
// panel solve in place: A21 (rows x 128, ld lda) <- A21 L11^{-T}; y -= P zk when y != NULL.  The copy the rank updates
// read goes to out0 (rows < split, leading dimension ld0) and out1 (rows >= split, re-based to row 0, ld1).
// inv_k supplies the 8x8 diagonal-block inverses of L11 (its own diagonal blocks).
void launch_panel_trsm(double* A21, long lda, const double* L11, long ldl, const double* inv_k, double* out0, long ld0,
                       int split, double* out1, long ld1, int rows, const double* zk, double* y, cudaStream_t s);
// alpha = L^{-T} z in one launch (chain of CTAs, flags: nblk ints of scratch)
void launch_backsolve_chain(const double* L, long ld, int nblk, const double* inv, const double* z, double* alpha,
                            int* flags, cudaStream_t s);
void launch_transpose(double* out, long ldo, const double* in, long ldi, int rows, int cols, cudaStream_t s);
// `batch` independent transposes, operand b at base + b * stride
void launch_transpose_batched(double* out, long ldo, long stride_out, const double* in, long ldi, long stride_in,
                              int rows, int cols, int batch, cudaStream_t s);
void launch_copy2d(double* out, long ldo, const double* in, long ldi, int rows, int cols, cudaStream_t s);
void launch_add_diag(double* A, long lda, const double* d, int n, cudaStream_t s);
void launch_fill(double* p, long n, double v, cudaStream_t s);
// set the strictly-upper block triangle + padding of a lower-block matrix to a clean state
void launch_set_identity_pad(double* A, long lda, int n_valid, int n_pad, cudaStream_t s);

// gradient reduction: for every requested hyper-parameter p,
//   g[p] = sum_{i>j} W_ij dK_ij(p) + 1/2 sum_i W_ii dK_ii(p),  W = w_outer * a a^T - S  (S lower-valid)
struct GradReduceParams {
    CovParams cp;
    const double* X;
    const int32_t* n;
    int N;               // points (matrix dimension that is valid)
    const double* S;     // N_pad x N_pad, lower triangle valid
    long lds;
    const double* a;     // vector for the outer product (alpha) or NULL
    int nidx;
    int idx[GPT_MAX_PARAMS];  // hyper_deriv indices
    double* partials;    // (num_ctas x nidx) scratch
    double* out;         // nidx
};
void launch_grad_reduce(const GradReduceParams& p, cudaStream_t s);
// out[0] = -1/2 z^T z - sum_k logdet[k] - M/2 log(2 pi): the fused log-likelihood reduction (one CTA, fixed order)
void launch_ll_reduce(const double* z, int M, const double* logdet, int nblk, double* out, cudaStream_t s);
// out[0] = (sum_i [a_i (y_i - a_i D_i) + Kinv_ii D_i] - n) / sigma_f, D_i = d[i] + noise2: the sigma_f entry of the gradient
// for kernels of the form sigma_f^2 g, no transformation matrix
void launch_sigma_identity(const double* Kinv, long ld, const double* a, const double* y, const double* d, double noise2,
                           int n, double sigma_f, double* out, cudaStream_t s);
// out[0] = sum_{i<n} A[i][i], out[1] = sum_{i<n} v[i]^2
void launch_trace_and_sumsq(const double* A, long lda, const double* v, int n, double* out, cudaStream_t s);

// ---- predict.cu ------------------------------------------------------------------------------------
// mean[s] = sum_i Kst[s][i] * alpha[i]  (Kst: rows x ld, first n columns)
void launch_rowdot(const double* Kst, long ld, int rows, int n, const double* alpha, double* mean, cudaStream_t s);
// var[s] = kss[s] - sum_{c<n} V[s][c]^2
void launch_row_var(const double* V, long ld, int rows, int n, const double* kss, double* var, cudaStream_t s);
// fused predictive mean (K* never stored): mean[s] = sum_i k(X_i, x*_s) u_i; partial: nsplit x Ms scratch
int predict_mean_nsplit(int N, int Ms);
void launch_predict_mean_fused(const CovParams& cp, const double* X, const int32_t* n, const double* u, int N,
                               const double* Xs, const int32_t* ns, int Ms, int low_order, double* partial,
                               double* mean, cudaStream_t s);
// kss[s] = k(x*_s, x*_s) with orders ns
void launch_prior_diag(const CovParams& cp, const double* Xs, const int32_t* ns, int rows, double* kss, cudaStream_t s);

// ---- batched4.cu : many-theta persistent kernel, 4 CTAs (= 4 thetas) per SM --------------------------
struct BatchedParams {
    int kid, D, nparams;
    int M;                 // observations (= latent points; no T on this path)
    int nT;                // 64-row tiles, ceil(M / 64)
    const double* X;       // M x D
    const int32_t* n;      // M x D
    int low_order;         // every derivative order <= 1: branch-free closed forms in the tile loops
    const double* y;       // M (shared) or B x M when y_stride != 0
    long y_stride;
    const double* diag;    // err_y^2 + diag_factor*eps, length M
    int B;
    const double* thetas;  // B x (nparams + 1): kernel params then sigma_noise
    int nidx;              // number of gradient entries (0: ll only)
    int idx[GPT_MAX_PARAMS];  // hyper_deriv index per entry; nparams = noise sigma
    double* ll;            // B
    double* grad;          // B x nidx
    int* status;           // B
    double* alpha_out;     // B x M or NULL
    double* workspace;     // per-CTA tile storage
    size_t ws_per_cta;     // in doubles
    int* counter;          // dynamic theta scheduler
    int short_forms = 0;     // SE, D <= 2, orders <= 1: short closed forms + cached exponentials (batched4.cu)
    size_t eb_off = 0;       // offset (doubles) inside the CTA workspace of the cached sigma^2 exp(-r^2/2) tiles
    long long* phase_cycles;  // optional (8): per-phase cycle sums, only with -DGPT_PHASE_TIMING
    // batched prediction (gpt_predict_batched): Ms test points appended to the point arrays as extra tile rows.  X / n then
    // hold nT * 64 rows for the training set (zero padded) followed by the Ms test points; the sweep-1 panel tiles of a
    // test row block are the rows L*(t, k) = (L^{-1} K*)^T of the extended factor: mean = sum_k L*(t,k) z_k,
    // var = k** - sum_k |L*(t,k)|^2 row by row.
    int Ms = 0, nTs = 0;
    size_t ts_off = 0;       // workspace offset (doubles) of the nTs x nT test-row tiles
    size_t pv_off = 0;       // workspace offset of the running mean (nTs * 64) and sum of squares (nTs * 64)
    double* pmean = nullptr; // B x Ms
    double* pvar = nullptr;  // B x Ms
    // kid == GPT_KERNEL_COMPOSITE: the structure (leaf kernels, parameter counts, product terms); every CTA keeps the
    // leaves of its current theta in its workspace at comp_off (doubles)
    int comp_nleaf = 0, comp_nterms = 0;
    int32_t comp_kids[GPT_MAX_LEAVES] = {0}, comp_nps[GPT_MAX_LEAVES] = {0}, comp_masks[GPT_MAX_TERMS] = {0};
    size_t comp_off = 0;
};
size_t batched_ws_doubles_per_cta(int nT);
size_t batched_lower_tiles(int nT);
int batched4_ctas_per_sm();
void launch_ll_batched4(const BatchedParams& p, int num_ctas, cudaStream_t s);
