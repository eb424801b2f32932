// C-ABI entry points (include/gptb200.h) and host-side orchestration of the blocked algorithms.
// Everything numerical runs in the CUDA kernels of this library; the host code below only sequences
// launches on the handle's stream.  No CPU fallback exists.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gptb200.h"
#include "common.cuh"
#include "internal.h"

#define NB GPT_NB

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct GradSlots {
    int s[GPT_MAX_PARAMS];
};

}  // namespace

struct gpt_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t side_stream = nullptr;  // lookahead stream of the blocked Cholesky
    cudaEvent_t ev_col = nullptr, ev_panel = nullptr, ev_rest[2] = {nullptr, nullptr};
    std::string err;
    int64_t launches = 0;

    // training data
    int N = 0, M = 0, D = 0, Np = 0, Mp = 0;
    bool hasT = false;
    int max_order = 0;  // largest derivative order in n
    std::vector<double> h_err2;  // err_y^2
    DevBuf X, n, y, diag, T, Tt;

    // kernel
    int kid = -1, nparams = 0;
    double diag_factor = 1e2;
    // kernel algebra (gpt_define_composite): structure, and the per-theta leaf parameters on the device
    int comp_nleaf = 0, comp_nterms = 0, comp_nparams = 0;
    int32_t comp_kids[GPT_MAX_LEAVES] = {0}, comp_nps[GPT_MAX_LEAVES] = {0}, comp_masks[GPT_MAX_TERMS] = {0};
    std::vector<CovComposite> comp_host;  // staging (kept alive: the uploads are asynchronous)
    DevBuf comp_dev;

    // single-theta factor state
    bool factor_valid = false;
    CovParams cp;
    double noise_sigma = 0.0;
    DevBuf A, Klat, W, Inv, P, z, alpha, logdet, info, scal, llred;
    // gradient workspaces
    DevBuf XT, Kinv, S, partials, gout, u, Sg, Yt;
    // predict workspaces
    DevBuf Xs, ns, Kst, Kso, kss, mean, var, cov, Rt, smp;
    // batched
    DevBuf b_thetas, b_y, b_ll, b_grad, b_status, b_alpha, b_ws, b_counter;
    DevBuf b_Xext, b_next, b_pmean, b_pvar;  // batched prediction: extended point arrays, outputs
    DevBuf ds_C, ds_inv, ds_panel, ds_logdet, ds_info, ds_R, ds_Rt, ds_O, ds_mu, ds_jit;  // draw_sample scratch
    DevBuf flags;  // backsolve chain: one int per 128-row block
    DevBuf Vtmp;  // predict: one 128-column block of V^T (out-of-place multiply by the block inverse)
};

namespace {

int fail(gpt_handle* h, int code, const char* fmt, const char* detail = "") {
    char buf[512];
    snprintf(buf, sizeof(buf), fmt, detail);
    if (h) h->err = buf;
    return code;
}

#define CUDA_OK(h, call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) return fail((h), GPT_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e__)); \
    } while (0)

int ensure(gpt_handle* h, DevBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.bytes >= bytes) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(h, GPT_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    b.bytes = bytes;
    e = cudaMemsetAsync(b.p, 0, bytes, h->stream);
    if (e != cudaSuccess) return fail(h, GPT_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    return 0;
}
void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}
template <typename T>
T* ptr(DevBuf& b) { return reinterpret_cast<T*>(b.p); }

int check_launch(gpt_handle* h) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, GPT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

int supported_kernel(int kid, int D, int nparams);

// Composite (sum of products of device kernels): every leaf must be supported in dimension D on its own.
int supported_composite(const gpt_handle* h, int D, int nparams) {
    if (!h || h->comp_nleaf < 1 || nparams != h->comp_nparams) return 0;
    for (int q = 0; q < h->comp_nleaf; q++)
        if (h->comp_kids[q] == GPT_GIBBS_AUX || !supported_kernel(h->comp_kids[q], D, h->comp_nps[q])) return 0;
    return 1;
}

int supported_kernel_h(const gpt_handle* h, int kid, int D, int nparams) {
    if (kid == GPT_COMPOSITE) return (D >= 1 && D <= GPT_MAX_DIM) ? supported_composite(h, D, nparams) : 0;
    return supported_kernel(kid, D, nparams);
}

// d/dnu of the generic Matern kernel is the one hyper-derivative the device does not have
bool hyper_deriv_unavailable(const gpt_handle* h, int kid, int idx) {
    if (kid == GPT_KERNEL_MATERN) return idx == 1;
    if (kid == GPT_COMPOSITE) {
        int off = 0;
        for (int q = 0; q < h->comp_nleaf; q++) {
            if (h->comp_kids[q] == GPT_KERNEL_MATERN && idx == off + 1) return true;
            off += h->comp_nps[q];
        }
    }
    return false;
}

int supported_kernel(int kid, int D, int nparams) {
    if (D < 1 || D > GPT_MAX_DIM) return 0;
    switch (kid) {
        case GPT_SE:
        case GPT_MATERN52: return nparams == D + 1;
        case GPT_MATERN: return nparams == D + 2;
        case GPT_GIBBS_TANH: return D == 1 && nparams == 5;
        case GPT_GIBBS_AUX: return D == 3 && nparams == 1;
        default: return 0;
    }
}

// Blocked right-looking Cholesky of the nblk*128 square matrix A (lower), in place.  Writes the
// inverses of the diagonal blocks to inv (nblk x 128 x 128), per-block log-det shares, info, and
// (optionally) overwrites rhs with L^{-1} rhs.  All heavy work is DMMA GEMM (gemm.cu).
// Blocked right-looking Cholesky of the nblk*128 square matrix A (lower), in place, with one-block LOOKAHEAD:
// the trailing update of step k is split into the strip that finalises block column k+1 and the rest; the
// (serial) diagonal-block factorisation and the panel of step k+1 run on a side stream while the rest of the
// step-k update keeps the machine busy.  panel2 holds TWO panels (double buffered).  Writes the inverses of
// the diagonal blocks to inv (nblk x 128 x 128), per-block log-det shares, info, and (optionally) overwrites rhs
// with L^{-1} rhs.  All heavy work is DMMA GEMM (gemm.cu).
// doubles of panel scratch blocked_potrf needs for an nblk-block matrix: two pair buffers + one 128 x 128 block
size_t potrf_panel_doubles(int nblk) { return (size_t)4 * nblk * NB * NB + (size_t)NB * NB; }

// Unpaired variant (rank-128 updates): for small matrices the per-step critical chain decides, and the pairing below
// doubles the in-kernel diagonal update of every other step (M = 4000: 3.1 ms against 3.6 ms).
int blocked_potrf_k128(gpt_handle* h, double* A, long ld, int nblk, double* inv, double* panel2, double* rhs, double* logdet,
                       int* info) {
    cudaStream_t sm = h->stream;
    if (!h->side_stream) {
        // highest priority: its few CTAs are placed as soon as an SM frees up, ahead of the queued update tiles
        int prio_lo = 0, prio_hi = 0;
        CUDA_OK(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_OK(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, prio_hi));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_col, cudaEventDisableTiming));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_panel, cudaEventDisableTiming));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_rest[0], cudaEventDisableTiming));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_rest[1], cudaEventDisableTiming));
    }
    cudaStream_t ss = h->side_stream;
    CUDA_OK(h, cudaMemsetAsync(info, 0, sizeof(int), sm));
    // Dependencies (block column c is final once strip(c-1) has run; rest(j) touches block columns >= j+2):
    //   potrf_diag(k): rest(k-2) [ev_rest, alternating pair] + its own in-kernel update with panel(k-1)
    //   trsm(k)      : strip(k-1) [ev_col]            strip(k), rest(k): trsm(k) [ev_panel], in main-stream order
    CUDA_OK(h, cudaEventRecord(h->ev_col, sm));  // everything queued so far (assembly, rhs set-up)
    for (int k = 0; k < nblk; k++) {
        double* Akk = A + (long)k * NB * ld + (long)k * NB;
        double* inv_k = inv + (size_t)k * NB * NB;
        double* panel = panel2 + (size_t)(k & 1) * (size_t)nblk * NB * NB;
        const double* panel_prev = panel2 + (size_t)((k - 1) & 1) * (size_t)nblk * NB * NB;
        const int rest = nblk - k - 1;
        // ---- side stream: diagonal block + panel of step k ----
        if (k == 0) CUDA_OK(h, cudaStreamWaitEvent(ss, h->ev_col, 0));
        if (k >= 2) CUDA_OK(h, cudaStreamWaitEvent(ss, h->ev_rest[k & 1], 0));
        launch_potrf_diag(Akk, ld, inv_k, rhs ? rhs + (long)k * NB : nullptr, logdet + k, info, k * NB,
                          k > 0 ? panel_prev : nullptr, NB, NB, ss);
        h->launches++;
        if (rest > 0) {
            double* A21 = A + (long)(k + 1) * NB * ld + (long)k * NB;
            if (k > 0) CUDA_OK(h, cudaStreamWaitEvent(ss, h->ev_col, 0));
            // panel P = A21 L11^{-T}: one-pass blocked substitution (factor.cu: panel_trsm_kernel), in place, plus the
            // contiguous copy the rank-128 update reads and the right-hand-side update
            launch_panel_trsm(A21, ld, Akk, ld, inv_k, panel, NB, 0, panel, NB, rest * NB,
                              rhs ? rhs + (long)k * NB : nullptr, rhs ? rhs + (long)(k + 1) * NB : nullptr, ss);
            h->launches++;
        }
        CUDA_OK(h, cudaEventRecord(h->ev_panel, ss));
        // ---- main stream: trailing update of step k, next block column first ----
        CUDA_OK(h, cudaStreamWaitEvent(sm, h->ev_panel, 0));
        if (rest > 0) {
            if (rest > 1) {
                GemmParams c;  // strip: block column k+1, rows >= k+2 (the diagonal block is potrf_diag(k+1)'s)
                c.C = A + (long)(k + 2) * NB * ld + (long)(k + 1) * NB; c.ldc = ld;
                c.A = panel + (size_t)NB * NB; c.lda = NB;
                c.B = panel; c.ldb = NB;
                c.tiles_m = rest - 1; c.tiles_n = 1; c.K = NB;
                c.alpha = -1.0; c.beta = 1.0; c.lower_only = 0; c.kbegin_row = 0;
                launch_gemm_nt(c, sm);
                h->launches++;
            }
            CUDA_OK(h, cudaEventRecord(h->ev_col, sm));
            if (rest > 1) {
                GemmParams u;  // rest: block columns >= k+2 (lower tiles only)
                u.C = A + (long)(k + 2) * NB * ld + (long)(k + 2) * NB; u.ldc = ld;
                u.A = panel + (size_t)NB * NB; u.lda = NB;
                u.B = panel + (size_t)NB * NB; u.ldb = NB;
                u.tiles_m = rest - 1; u.tiles_n = rest - 1; u.K = NB;
                u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1; u.kbegin_row = 0;
                launch_gemm_nt(u, sm);
                h->launches++;
            }
            CUDA_OK(h, cudaEventRecord(h->ev_rest[k & 1], sm));
        }
    }
    // join: nothing is left running on the side stream that the main stream has not waited for (ev_panel of the
    // last step), so work queued on the main stream after this call sees the complete factor
    return check_launch(h);
}

constexpr int PAIR_MIN_BLOCKS = 80;  // M >= ~10k: the trailing update dominates and rank-256 pays (tools/pair_tune.py)

int blocked_potrf(gpt_handle* h, double* A, long ld, int nblk, double* inv, double* pbuf, double* rhs, double* logdet,
                  int* info) {
    int pair_min = PAIR_MIN_BLOCKS;
    if (const char* e = getenv("GPT_POTRF_PAIR_MIN")) pair_min = atoi(e);  // tests: force either variant
    if (nblk < pair_min) return blocked_potrf_k128(h, A, ld, nblk, inv, pbuf, rhs, logdet, info);
    cudaStream_t sm = h->stream;
    if (!h->side_stream) {
        // highest priority: its few CTAs are placed as soon as an SM frees up, ahead of the queued update tiles
        int prio_lo = 0, prio_hi = 0;
        CUDA_OK(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_OK(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, prio_hi));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_col, cudaEventDisableTiming));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_panel, cudaEventDisableTiming));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_rest[0], cudaEventDisableTiming));
        CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_rest[1], cudaEventDisableTiming));
    }
    cudaStream_t ss = h->side_stream;
    CUDA_OK(h, cudaMemsetAsync(info, 0, sizeof(int), sm));
    // Panels are taken in PAIRS (2p, 2p+1) so that the big trailing update runs with a contraction of 256 (33.0 vs
    // 31.2 TFLOP/s for the GEMM).  Pair buffer: row 0 = block row 2p+2, columns 0..127 = panel 2p (without its first
    // block, which goes to pfirst), columns 128..255 = panel 2p+1.  Per pair, on the main stream:
    //   strip(2p)    : column 2p+1, rows >= 2p+2,  -= P_2p P_2p[first]^T                         (K = 128)
    //   strip2(2p+1) : columns 2p+2, 2p+3, rows >= 2p+3 (incl. the diagonal block of 2p+3)         (K = 256)
    //   rest(2p+1)   : block columns >= 2p+4, lower tiles                                         (K = 256)
    // and on the side stream potrf(k) / trsm(k); potrf applies the update its diagonal block still lacks itself
    // (even k: K = 256 from the previous pair buffer, odd k: K = 128 from pfirst).  Dependencies: trsm(k) waits for
    // the latest strip (ev_col), the strips wait for trsm (ev_panel), potrf of pair p waits for rest of pair p-2
    // (ev_rest, alternating) -- never for the rest that is running, which is what the lookahead hides it behind.
    const long LP = 2 * NB;
    double* pair[2] = {pbuf, pbuf + (size_t)nblk * NB * LP};
    double* pfirst = pbuf + (size_t)2 * nblk * NB * LP;
    CUDA_OK(h, cudaEventRecord(h->ev_col, sm));  // everything queued so far (assembly, rhs set-up)
    for (int k = 0; k < nblk; k++) {
        const bool even = (k & 1) == 0;
        const int p = k >> 1;
        double* pb = pair[p & 1];
        double* Akk = A + (long)k * NB * ld + (long)k * NB;
        double* inv_k = inv + (size_t)k * NB * NB;
        const int rest = nblk - k - 1;
        // ---- side stream: diagonal block + panel of step k ----
        if (k == 0) CUDA_OK(h, cudaStreamWaitEvent(ss, h->ev_col, 0));
        if (even && p >= 2) CUDA_OK(h, cudaStreamWaitEvent(ss, h->ev_rest[p & 1], 0));
        const double* pprev = nullptr;
        long ldp = NB;
        int pcols = NB;
        if (even && k > 0) {
            pprev = pair[(p - 1) & 1];
            ldp = LP;
            pcols = 2 * NB;
        } else if (!even) {
            pprev = pfirst;
        }
        launch_potrf_diag(Akk, ld, inv_k, rhs ? rhs + (long)k * NB : nullptr, logdet + k, info, k * NB, pprev, ldp, pcols,
                          ss);
        h->launches++;
        if (rest > 0) {
            double* A21 = A + (long)(k + 1) * NB * ld + (long)k * NB;
            // column k is final after strip2 of the previous pair (main stream: ev_col); for odd k also after strip(k-1),
            // which runs on this stream
            if (even && k > 0) CUDA_OK(h, cudaStreamWaitEvent(ss, h->ev_col, 0));
            // panel P = A21 L11^{-T}: one-pass blocked substitution (factor.cu: panel_trsm_kernel), in place, plus the
            // copy the rank updates read and the right-hand-side update
            if (even)
                launch_panel_trsm(A21, ld, Akk, ld, inv_k, pfirst, NB, NB, pb, LP, rest * NB,
                                  rhs ? rhs + (long)k * NB : nullptr, rhs ? rhs + (long)(k + 1) * NB : nullptr, ss);
            else
                launch_panel_trsm(A21, ld, Akk, ld, inv_k, pb + NB, LP, 0, pb + NB, LP, rest * NB,
                                  rhs ? rhs + (long)k * NB : nullptr, rhs ? rhs + (long)(k + 1) * NB : nullptr, ss);
            h->launches++;
            if (even && rest > 1) {
                // strip: block column k+1, rows >= k+2, panel k only.  On the side stream: it touches nothing the
                // running rank-256 update touches, and trsm(k+1) needs it.
                GemmParams c;
                c.C = A + (long)(k + 2) * NB * ld + (long)(k + 1) * NB; c.ldc = ld;
                c.A = pb; c.lda = LP;
                c.B = pfirst; c.ldb = NB;
                c.tiles_m = rest - 1; c.tiles_n = 1; c.K = NB;
                c.alpha = -1.0; c.beta = 1.0; c.lower_only = 0; c.kbegin_row = 0;
                launch_gemm_nt(c, ss);
                h->launches++;
            }
        }
        CUDA_OK(h, cudaEventRecord(h->ev_panel, ss));
        // ---- main stream ----
        CUDA_OK(h, cudaStreamWaitEvent(sm, h->ev_panel, 0));
        if (!even) {
            const int tm = nblk - (k + 2);  // block rows >= k+2
            if (tm >= 1) {
                GemmParams c;  // strip2: block columns k+1 and k+2, rows >= k+2, both panels of the pair
                c.C = A + (long)(k + 2) * NB * ld + (long)(k + 1) * NB; c.ldc = ld;
                c.A = pb + (size_t)NB * LP; c.lda = LP;
                c.B = pb; c.ldb = LP;
                c.tiles_m = tm; c.tiles_n = 2; c.K = 2 * NB;
                c.alpha = -1.0; c.beta = 1.0; c.lower_only = 0; c.kbegin_row = 0;
                launch_gemm_nt(c, sm);
                h->launches++;
            }
            CUDA_OK(h, cudaEventRecord(h->ev_col, sm));
            const int tr = nblk - (k + 3);  // block rows / columns >= k+3
            if (tr >= 1) {
                GemmParams u;  // rest: lower tiles of the trailing matrix, rank-256 update
                u.C = A + (long)(k + 3) * NB * ld + (long)(k + 3) * NB; u.ldc = ld;
                u.A = pb + (size_t)2 * NB * LP; u.lda = LP;
                u.B = pb + (size_t)2 * NB * LP; u.ldb = LP;
                u.tiles_m = tr; u.tiles_n = tr; u.K = 2 * NB;
                u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1; u.kbegin_row = 0;
                launch_gemm_nt(u, sm);
                h->launches++;
            }
            CUDA_OK(h, cudaEventRecord(h->ev_rest[p & 1], sm));
        }
    }
    // join: nothing is left running on the side stream that the main stream has not waited for (ev_panel of the
    // last step), so work queued on the main stream after this call sees the complete factor
    return check_launch(h);
}

int upload(gpt_handle* h, DevBuf& b, const void* src, size_t bytes) {
    int rc = ensure(h, b, bytes);
    if (rc) return rc;
    CUDA_OK(h, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

// CovParams for (kid, params); a composite's leaves go to slot `slot` of the device array `dev` (`nslots` long) through
// the handle's staging vector, and cp.comp points at that slot.
int make_cov_params(gpt_handle* h, int kid, int D, int nparams, const double* params, CovParams& cp, DevBuf& dev,
                    int slot = 0, int nslots = 1) {
    cov_params_init(cp, kid, D, nparams, params);
    if (kid != GPT_COMPOSITE) return 0;
    if ((int)h->comp_host.size() < nslots) h->comp_host.resize(nslots);
    if (comp_init(h->comp_host[slot], D, h->comp_nleaf, h->comp_kids, h->comp_nps, h->comp_nterms, h->comp_masks, params) < 0)
        return fail(h, GPT_ERR_USAGE, "composite kernel: invalid description");
    int rc = ensure(h, dev, sizeof(CovComposite) * (size_t)nslots);
    if (rc) return rc;
    CovComposite* d = reinterpret_cast<CovComposite*>(dev.p) + slot;
    CUDA_OK(h, cudaMemcpyAsync(d, &h->comp_host[slot], sizeof(CovComposite), cudaMemcpyHostToDevice, h->stream));
    cp.comp = d;
    return 0;
}

int upload_padded(gpt_handle* h, DevBuf& b, const double* src, int rows, int cols, int rows_pad, int cols_pad) {
    int rc = ensure(h, b, (size_t)rows_pad * cols_pad * sizeof(double));
    if (rc) return rc;
    CUDA_OK(h, cudaMemsetAsync(b.p, 0, (size_t)rows_pad * cols_pad * sizeof(double), h->stream));
    CUDA_OK(h, cudaMemcpy2DAsync(b.p, (size_t)cols_pad * sizeof(double), src, (size_t)cols * sizeof(double),
                                 (size_t)cols * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int refresh_diag(gpt_handle* h) {
    // err_y^2 + diag_factor * eps  (gaussian_process.py:1449-1450)
    std::vector<double> d(h->M);
    const double jit = h->diag_factor * 2.220446049250313e-16;
    for (int i = 0; i < h->M; i++) d[i] = h->h_err2[i] + jit;
    int rc = upload(h, h->diag, d.data(), sizeof(double) * h->M);
    if (rc) return rc;
    CUDA_OK(h, cudaStreamSynchronize(h->stream));  // d goes out of scope
    return 0;
}

// Factor the matrix currently in h->A (Mp x Mp, K_tot with identity padding), then alpha; ll is left in h->llred and
// the LAPACK-style info word in h->info (device).  No synchronisation.
int factor_and_solve_device(gpt_handle* h);

// ... and bring ll / info back to the host.
int factor_and_solve(gpt_handle* h, double* ll, int* status) {
    int rc = factor_and_solve_device(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    double hll = 0.0;
    int hinfo = 0;
    CUDA_OK(h, cudaMemcpyAsync(&hll, h->llred.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaMemcpyAsync(&hinfo, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    if (hinfo > h->M) hinfo = 0;  // only padding rows (identity) could report beyond M; cannot happen, be safe
    *status = hinfo;
    *ll = hll;
    h->factor_valid = (hinfo == 0);
    return 0;
}

int factor_and_solve_device(gpt_handle* h) {
    cudaStream_t s = h->stream;
    const int Mp = h->Mp, M = h->M, nblk = Mp / NB;
    int rc;
    if ((rc = ensure(h, h->Inv, (size_t)nblk * NB * NB * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->P, potrf_panel_doubles(nblk) * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->z, (size_t)Mp * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->alpha, (size_t)Mp * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->logdet, (size_t)nblk * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->info, sizeof(int)))) return rc;
    CUDA_OK(h, cudaMemsetAsync(h->z.p, 0, (size_t)Mp * sizeof(double), s));
    CUDA_OK(h, cudaMemcpyAsync(h->z.p, h->y.p, (size_t)M * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if ((rc = blocked_potrf(h, ptr<double>(h->A), Mp, nblk, ptr<double>(h->Inv), ptr<double>(h->P),
                            ptr<double>(h->z), ptr<double>(h->logdet), ptr<int>(h->info))))
        return rc;
    if ((rc = ensure(h, h->flags, sizeof(int) * nblk))) return rc;
    launch_backsolve_chain(ptr<double>(h->A), Mp, nblk, ptr<double>(h->Inv), ptr<double>(h->z), ptr<double>(h->alpha),
                           ptr<int>(h->flags), s);
    h->launches++;
    if ((rc = check_launch(h))) return rc;
    // fused reduction on the device: -1/2 z^T z - sum log diag(L) - M/2 log 2 pi; one scalar and the info word come back
    if ((rc = ensure(h, h->llred, sizeof(double)))) return rc;
    launch_ll_reduce(ptr<double>(h->z), M, ptr<double>(h->logdet), nblk, ptr<double>(h->llred), s);
    h->launches++;
    return check_launch(h);
}

// K_tot from a latent covariance already in h->Klat (Np x Np, symmetric, noise included): T K T' + diag
int transform_latent(gpt_handle* h) {
    cudaStream_t s = h->stream;
    const int Mp = h->Mp, Np = h->Np;
    int rc;
    if ((rc = ensure(h, h->W, (size_t)Mp * Np * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->A, (size_t)Mp * Mp * sizeof(double)))) return rc;
    GemmParams g;
    g.C = ptr<double>(h->W); g.ldc = Np;
    g.A = ptr<double>(h->T); g.lda = Np;
    g.B = ptr<double>(h->Klat); g.ldb = Np;  // symmetric, so B[n][k] = K[k][n]
    g.tiles_m = Mp / NB; g.tiles_n = Np / NB; g.K = Np;
    g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0; g.kbegin_row = 0;
    launch_gemm_nt(g, s);
    GemmParams g2 = g;
    g2.C = ptr<double>(h->A); g2.ldc = Mp;
    g2.A = ptr<double>(h->W); g2.lda = Np;
    g2.B = ptr<double>(h->T); g2.ldb = Np;
    g2.tiles_m = Mp / NB; g2.tiles_n = Mp / NB; g2.K = Np;
    launch_gemm_nt(g2, s);
    launch_add_diag(ptr<double>(h->A), Mp, ptr<double>(h->diag), h->M, s);
    launch_set_identity_pad(ptr<double>(h->A), Mp, h->M, Mp, s);
    h->launches += 4;
    return check_launch(h);
}

int assemble_train(gpt_handle* h, const CovParams& cp, double noise_sigma, int hyper_deriv, double* out, int pad,
                   bool add_diag, bool lower_tiles_only = false) {
    AssembleParams a;
    a.cp = cp;
    a.Xr = ptr<double>(h->X); a.nr = ptr<int32_t>(h->n); a.Mr = h->N;
    a.Xc = a.Xr; a.nc = a.nr; a.Mc = h->N;
    a.out = out; a.ldo = pad; a.rows_pad = pad; a.cols_pad = pad;
    a.hyper_deriv = hyper_deriv; a.swap_roles = 0; a.symmetric = 1;
    a.diag_add = add_diag ? ptr<double>(h->diag) : nullptr;
    a.diag_const = noise_sigma * noise_sigma;
    a.pad_identity = add_diag ? 1 : 0;
    a.lower_tiles_only = lower_tiles_only ? 1 : 0;
    a.low_order = (h->max_order <= 1) ? 1 : 0;
    launch_assemble(a, h->stream);
    h->launches++;
    return check_launch(h);
}

// K^{-1} (lower triangle valid) into h->Kinv from the resident factor: XT = L^{-T} by block
// substitution (all GEMM), then XT XT^T restricted to lower tiles.
int compute_Kinv(gpt_handle* h) {
    cudaStream_t s = h->stream;
    const int Mp = h->Mp, nblk = Mp / NB;
    int rc;
    if ((rc = ensure(h, h->XT, (size_t)Mp * Mp * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->Kinv, (size_t)Mp * Mp * sizeof(double)))) return rc;
    double* XT = ptr<double>(h->XT);
    double* L = ptr<double>(h->A);
    double* Inv = ptr<double>(h->Inv);
    // ---- XT = L^{-T} by recursive doubling -------------------------------------------------------------------------
    // X = L^{-1} = [[X11, 0], [-X22 L21 X11, X22]]: all merges of one level are independent, so a level is three
    // batched launches (the per-block-column substitution it replaces was 2 launches per block with at most 2 I CTAs
    // and a contraction as long as the matrix: 12 TFLOP/s at M = 12288, 6 ms of an 11 ms ll+grad at M = 4000).
    // X (lower) lives in the Kinv buffer until lauum overwrites it; XT (upper) is the transposed copy every NT
    // product needs as its second operand.  Per merge (left block range 1, right block range 2):
    //   W^T = XT11 L21^T  -> XT12 region   (A = XT11 upper block-triangular: kbegin_row)
    //   X21 = -X22 W      -> X21 region    (A = X22 lower block-triangular: kend_row; B = W^T)
    //   XT12 = X21^T
    double* X = ptr<double>(h->Kinv);
    const long dstride = (long)NB * Mp + NB;  // one block down the diagonal
    launch_transpose_batched(XT, Mp, dstride, Inv, NB, (long)NB * NB, NB, NB, nblk, s);  // XT_kk = Inv_k^T
    launch_transpose_batched(X, Mp, dstride, XT, Mp, dstride, NB, NB, nblk, s);           // X_kk  = Inv_k
    h->launches += 2;
    for (int b = 1; b < nblk; b *= 2) {
        const int nfull = nblk / (2 * b);                  // merges with a full right half
        const int rem = nblk % (2 * b);
        const int b2r = (rem > b) ? rem - b : 0;           // ragged merge: right half of b2r blocks
        for (int pass = 0; pass < 2; pass++) {
            const int batch = pass == 0 ? nfull : (b2r > 0 ? 1 : 0);
            const int b2 = pass == 0 ? b : b2r;
            if (batch == 0) continue;
            const long p0 = pass == 0 ? 0 : (long)nfull * 2 * b;  // first block of the (first) left range
            const long pstride = (long)2 * b * dstride;           // from one merge to the next
            double* XT11 = XT + p0 * dstride;
            double* XT12 = XT11 + (long)b * NB;                   // rows range 1, columns range 2
            const double* L21 = L + (p0 + b) * NB * Mp + p0 * NB; // rows range 2, columns range 1
            double* X22 = X + (p0 + b) * dstride;
            double* X21 = X + (p0 + b) * NB * Mp + p0 * NB;
            GemmParams g1;
            g1.C = XT12; g1.ldc = Mp;
            g1.A = XT11; g1.lda = Mp;
            g1.B = L21; g1.ldb = Mp;
            g1.tiles_m = b; g1.tiles_n = b2; g1.K = b * NB;
            g1.alpha = 1.0; g1.beta = 0.0; g1.lower_only = 0; g1.kbegin_row = 1;
            g1.batch = batch; g1.strideA = pstride; g1.strideB = pstride; g1.strideC = pstride;
            launch_gemm_nt(g1, s);
            GemmParams g2;
            g2.C = X21; g2.ldc = Mp;
            g2.A = X22; g2.lda = Mp;
            g2.B = XT12; g2.ldb = Mp;
            g2.tiles_m = b2; g2.tiles_n = b; g2.K = b2 * NB;
            g2.alpha = -1.0; g2.beta = 0.0; g2.lower_only = 0; g2.kbegin_row = 0; g2.kend_row = 1;
            g2.batch = batch; g2.strideA = pstride; g2.strideB = pstride; g2.strideC = pstride;
            launch_gemm_nt(g2, s);
            launch_transpose_batched(XT12, Mp, pstride, X21, Mp, pstride, b2 * NB, b * NB, batch, s);
            h->launches += 3;
        }
    }
    GemmParams g;
    g.C = ptr<double>(h->Kinv); g.ldc = Mp;
    g.A = XT; g.lda = Mp;
    g.B = XT; g.ldb = Mp;
    g.tiles_m = nblk; g.tiles_n = nblk; g.K = Mp;
    g.alpha = 1.0; g.beta = 0.0; g.lower_only = 1; g.kbegin_row = 1;
    launch_gemm_nt(g, s);
    h->launches++;
    return check_launch(h);
}

__global__ void symmetrize_from_lower_kernel(double* __restrict__ A, long lda, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < n && r < n && c > r) A[(long)r * lda + c] = A[(long)c * lda + r];
}

__global__ void tril_kernel(double* __restrict__ A, long lda, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < n && r < n && c > r) A[(long)r * lda + c] = 0.0;
}

__global__ void add_mean_kernel(double* __restrict__ out, long ldo, const double* __restrict__ mean, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < cols && r < rows) out[(long)r * ldo + c] += mean[r];
}

__global__ void elem_dot_lower_kernel(const double* __restrict__ W, long ldw, const double* __restrict__ a,
                                      const double* __restrict__ dK, long ldk, int n, double* __restrict__ partial) {
    // partial[block] = sum over this block's rows of sum_c (a_r a_c - W_rc) dK_rc   (full square, W symmetric-full)
    __shared__ double sh[256];
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += 256) s += (a[r] * a[c] - W[(long)r * ldw + c]) * dK[(long)r * ldk + c];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[r] = sh[0];
}

}  // namespace

extern "C" {

int gpt_version(void) { return 100; }

int gpt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int gpt_create(int device, gpt_handle** out) {
    if (!out) return GPT_ERR_USAGE;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return GPT_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return GPT_ERR_CUDA;
    gpt_handle* h = new gpt_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return GPT_ERR_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return 0;
}

void gpt_destroy(gpt_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf* all[] = {&h->X, &h->n, &h->y, &h->diag, &h->T, &h->Tt, &h->A, &h->Klat, &h->W, &h->Inv, &h->P, &h->z,
                     &h->alpha, &h->logdet, &h->info, &h->scal, &h->XT, &h->Kinv, &h->S, &h->partials,
                     &h->gout, &h->u, &h->Sg, &h->Yt, &h->Xs, &h->ns, &h->Kst, &h->Kso, &h->kss, &h->mean, &h->var,
                     &h->cov, &h->Rt, &h->smp, &h->b_thetas, &h->b_y, &h->b_ll, &h->b_grad, &h->b_status,
                     &h->llred, &h->b_alpha, &h->b_ws, &h->b_counter, &h->Vtmp, &h->flags, &h->ds_C, &h->ds_inv, &h->ds_panel, &h->ds_logdet,
                     &h->ds_info, &h->ds_R, &h->ds_Rt, &h->ds_O, &h->ds_mu, &h->ds_jit, &h->comp_dev,
                     &h->b_Xext, &h->b_next, &h->b_pmean, &h->b_pvar};
    for (DevBuf* b : all) release(*b);
    if (h->side_stream) {
        cudaStreamSynchronize(h->side_stream);
        cudaStreamDestroy(h->side_stream);
        cudaEventDestroy(h->ev_col);
        cudaEventDestroy(h->ev_panel);
        cudaEventDestroy(h->ev_rest[0]);
        cudaEventDestroy(h->ev_rest[1]);
    }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

const char* gpt_last_error(gpt_handle* h) { return h ? h->err.c_str() : "null handle"; }

int gpt_set_stream(gpt_handle* h, void* cuda_stream) {
    if (!h) return GPT_ERR_USAGE;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int gpt_use_own_stream(gpt_handle* h) {
    if (!h) return GPT_ERR_USAGE;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->stream = h->own_stream;
    return 0;
}

int gpt_synchronize(gpt_handle* h) {
    if (!h) return GPT_ERR_USAGE;
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int64_t gpt_launch_count(gpt_handle* h) { return h ? h->launches : 0; }

int gpt_set_data(gpt_handle* h, int N, int M, int D, const double* X, const int32_t* n, const double* y,
                 const double* err_y, const double* T) {
    if (!h || N < 1 || M < 1 || D < 1 || D > GPT_MAX_DIM || !X || !n || !y || !err_y)
        return fail(h, GPT_ERR_USAGE, "gpt_set_data: bad arguments");
    if (!T && N != M) return fail(h, GPT_ERR_USAGE, "gpt_set_data: N must equal M when T is NULL");
    CUDA_OK(h, cudaSetDevice(h->device));
    h->N = N; h->M = M; h->D = D;
    h->Np = round_up(N, NB); h->Mp = round_up(M, NB);
    h->hasT = (T != nullptr);
    h->factor_valid = false;
    h->max_order = 0;
    for (long i = 0; i < (long)N * D; i++) {
        if (n[i] < 0) return fail(h, GPT_ERR_USAGE, "gpt_set_data: negative derivative order");
        if (n[i] > h->max_order) h->max_order = n[i];
    }
    int rc;
    if ((rc = upload(h, h->X, X, sizeof(double) * N * D))) return rc;
    if ((rc = upload(h, h->n, n, sizeof(int32_t) * N * D))) return rc;
    if ((rc = upload(h, h->y, y, sizeof(double) * M))) return rc;
    h->h_err2.resize(M);
    for (int i = 0; i < M; i++) h->h_err2[i] = err_y[i] * err_y[i];
    if (T) {
        if ((rc = upload_padded(h, h->T, T, M, N, h->Mp, h->Np))) return rc;
        if ((rc = ensure(h, h->Tt, (size_t)h->Np * h->Mp * sizeof(double)))) return rc;
        CUDA_OK(h, cudaMemsetAsync(h->Tt.p, 0, (size_t)h->Np * h->Mp * sizeof(double), h->stream));
        launch_transpose(ptr<double>(h->Tt), h->Mp, ptr<double>(h->T), h->Np, M, N, h->stream);
        h->launches++;
    }
    if ((rc = refresh_diag(h))) return rc;
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return check_launch(h);
}

int gpt_set_y(gpt_handle* h, const double* y) {
    if (!h || !y || h->M < 1) return fail(h, GPT_ERR_USAGE, "gpt_set_y: no data set");
    CUDA_OK(h, cudaSetDevice(h->device));
    h->factor_valid = false;
    int rc = upload(h, h->y, y, sizeof(double) * h->M);
    if (rc) return rc;
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int gpt_set_kernel(gpt_handle* h, int kernel_id, int nparams, double diag_factor) {
    if (!h) return GPT_ERR_USAGE;
    if (kernel_id < 0 || kernel_id > GPT_COMPOSITE || nparams < 1 || nparams > GPT_MAX_PARAMS)
        return fail(h, GPT_ERR_UNSUPPORTED, "gpt_set_kernel: unsupported kernel / parameter count");
    if (kernel_id == GPT_COMPOSITE && (h->comp_nleaf < 1 || nparams != h->comp_nparams))
        return fail(h, GPT_ERR_USAGE, "gpt_set_kernel: GPT_COMPOSITE needs gpt_define_composite first (parameter counts must agree)");
    CUDA_OK(h, cudaSetDevice(h->device));
    h->kid = kernel_id;
    h->nparams = nparams;
    h->factor_valid = false;
    if (diag_factor != h->diag_factor || h->diag.p == nullptr) {
        h->diag_factor = diag_factor;
        if (h->M > 0) {
            int rc = refresh_diag(h);
            if (rc) return rc;
        }
    }
    return 0;
}

int gpt_define_composite(gpt_handle* h, int nleaf, const int32_t* leaf_kernel_ids, const int32_t* leaf_nparams,
                         int nterms, const int32_t* term_masks) {
    if (!h || !leaf_kernel_ids || !leaf_nparams || !term_masks) return fail(h, GPT_ERR_USAGE, "gpt_define_composite: bad arguments");
    if (nleaf < 1 || nleaf > GPT_MAX_LEAVES || nterms < 1 || nterms > GPT_MAX_TERMS)
        return fail(h, GPT_ERR_UNSUPPORTED, "gpt_define_composite: at most 4 operand kernels and 8 product terms");
    int total = 0;
    for (int q = 0; q < nleaf; q++) {
        if (leaf_kernel_ids[q] < 0 || leaf_kernel_ids[q] >= GPT_GIBBS_AUX || leaf_nparams[q] < 1)
            return fail(h, GPT_ERR_UNSUPPORTED, "gpt_define_composite: operand kernels are SE, Matern-5/2, Matern, Gibbs-tanh");
        total += leaf_nparams[q];
    }
    if (total > GPT_MAX_PARAMS) return fail(h, GPT_ERR_UNSUPPORTED, "gpt_define_composite: more than 10 parameters in total");
    for (int t = 0; t < nterms; t++)
        if (term_masks[t] <= 0 || term_masks[t] >= (1 << nleaf)) return fail(h, GPT_ERR_USAGE, "gpt_define_composite: bad term mask");
    bool same = (h->comp_nleaf == nleaf && h->comp_nterms == nterms);
    for (int q = 0; same && q < nleaf; q++) same = (h->comp_kids[q] == leaf_kernel_ids[q] && h->comp_nps[q] == leaf_nparams[q]);
    for (int t = 0; same && t < nterms; t++) same = (h->comp_masks[t] == term_masks[t]);
    if (same) return 0;
    h->comp_nleaf = nleaf;
    h->comp_nterms = nterms;
    h->comp_nparams = total;
    for (int q = 0; q < nleaf; q++) { h->comp_kids[q] = leaf_kernel_ids[q]; h->comp_nps[q] = leaf_nparams[q]; }
    for (int t = 0; t < nterms; t++) h->comp_masks[t] = term_masks[t];
    if (h->kid == GPT_COMPOSITE) {  // the kernel set by gpt_set_kernel changed under the handle
        h->factor_valid = false;
        h->nparams = total;
    }
    return 0;
}

int gpt_cov_pairs(gpt_handle* h, int kernel_id, int D, int nparams, const double* params, int hyper_deriv,
                  int64_t npairs, const double* Xi, const double* Xj, const int32_t* ni, const int32_t* nj,
                  double* out) {
    if (!h || !params || npairs < 0) return fail(h, GPT_ERR_USAGE, "gpt_cov_pairs: bad arguments");
    if (!supported_kernel_h(h, kernel_id, D, nparams)) return fail(h, GPT_ERR_UNSUPPORTED, "gpt_cov_pairs: unsupported kernel");
    if (hyper_deriv >= nparams || hyper_deriv_unavailable(h, kernel_id, hyper_deriv))
        return fail(h, GPT_ERR_UNSUPPORTED, "hyper_deriv: index out of range, or d/dnu of the Matern kernel");
    if (npairs == 0) return 0;
    CUDA_OK(h, cudaSetDevice(h->device));
    CovParams cp;
    const size_t xb = sizeof(double) * npairs * D, nb_ = sizeof(int32_t) * npairs * D;
    DevBuf dXi, dXj, dni, dnj, dout, dcomp;
    int rc = 0;
    if ((rc = make_cov_params(h, kernel_id, D, nparams, params, cp, dcomp))) {
        release(dcomp);
        return rc;
    }
    if ((rc = upload(h, dXi, Xi, xb)) || (rc = upload(h, dXj, Xj, xb)) || (rc = upload(h, dni, ni, nb_)) ||
        (rc = upload(h, dnj, nj, nb_)) || (rc = ensure(h, dout, sizeof(double) * npairs))) {
        release(dXi); release(dXj); release(dni); release(dnj); release(dout); release(dcomp);
        return rc;
    }
    launch_cov_pairs(cp, hyper_deriv, npairs, ptr<double>(dXi), ptr<double>(dXj), ptr<int32_t>(dni), ptr<int32_t>(dnj),
                     ptr<double>(dout), h->stream);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(out, dout.p, sizeof(double) * npairs, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    release(dXi); release(dXj); release(dni); release(dnj); release(dout); release(dcomp);
    if (e != cudaSuccess) return fail(h, GPT_ERR_CUDA, "gpt_cov_pairs: %s", cudaGetErrorString(e));
    return 0;
}

int gpt_compute_Kij(gpt_handle* h, int kernel_id, int D, int nparams, const double* params, int hyper_deriv,
                    int Mi, const double* Xi, const int32_t* ni, int Mj, const double* Xj, const int32_t* nj,
                    double* K_out) {
    if (!h || !params || Mi < 1 || !Xi || !ni || !K_out) return fail(h, GPT_ERR_USAGE, "gpt_compute_Kij: bad arguments");
    if (!supported_kernel_h(h, kernel_id, D, nparams)) return fail(h, GPT_ERR_UNSUPPORTED, "gpt_compute_Kij: unsupported kernel");
    if (hyper_deriv >= nparams || hyper_deriv_unavailable(h, kernel_id, hyper_deriv))
        return fail(h, GPT_ERR_UNSUPPORTED, "hyper_deriv: index out of range, or d/dnu of the Matern kernel");
    CUDA_OK(h, cudaSetDevice(h->device));
    const bool sym = (Xj == nullptr);
    if (sym) Mj = Mi;
    CovParams cp;
    DevBuf dXi, dXj, dni, dnj, dout, dcomp;
    int rc = 0;
    auto cleanup = [&]() { release(dXi); release(dXj); release(dni); release(dnj); release(dout); release(dcomp); };
    if ((rc = make_cov_params(h, kernel_id, D, nparams, params, cp, dcomp))) {
        cleanup();
        return rc;
    }
    if ((rc = upload(h, dXi, Xi, sizeof(double) * Mi * D)) || (rc = upload(h, dni, ni, sizeof(int32_t) * Mi * D)) ||
        (rc = ensure(h, dout, sizeof(double) * (size_t)Mi * Mj))) {
        cleanup();
        return rc;
    }
    if (!sym) {
        if ((rc = upload(h, dXj, Xj, sizeof(double) * Mj * D)) || (rc = upload(h, dnj, nj, sizeof(int32_t) * Mj * D))) {
            cleanup();
            return rc;
        }
    }
    AssembleParams a;
    a.cp = cp;
    a.Xr = ptr<double>(dXi); a.nr = ptr<int32_t>(dni); a.Mr = Mi;
    a.Xc = sym ? a.Xr : ptr<double>(dXj); a.nc = sym ? a.nr : ptr<int32_t>(dnj); a.Mc = Mj;
    a.out = ptr<double>(dout); a.ldo = Mj; a.rows_pad = Mi; a.cols_pad = Mj;
    a.hyper_deriv = hyper_deriv; a.swap_roles = 0; a.symmetric = 0;
    a.diag_add = nullptr; a.diag_const = 0.0; a.pad_identity = 0;
    {
        int mo = 0;  // every derivative order <= 1: the short SE / Matern-5/2 tile generators apply
        for (size_t i = 0; i < (size_t)Mi * D; i++) mo = ni[i] > mo ? ni[i] : mo;
        if (!sym) for (size_t i = 0; i < (size_t)Mj * D; i++) mo = nj[i] > mo ? nj[i] : mo;
        a.low_order = (mo <= 1) ? 1 : 0;
    }
    launch_assemble(a, h->stream);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(K_out, dout.p, sizeof(double) * (size_t)Mi * Mj, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (e != cudaSuccess) return fail(h, GPT_ERR_CUDA, "gpt_compute_Kij: %s", cudaGetErrorString(e));
    return 0;
}

// Gradient of the resident factorisation, device part: h->gout[0 .. nkern) = the kernel-parameter entries (slot[i] is the
// position of entry i in the caller's grad_idx), h->gout[GPT_MAX_PARAMS] = tr K_tot^{-1}, [GPT_MAX_PARAMS + 1] = alpha^T alpha.
static int gradient_device(gpt_handle* h, const int32_t* grad_idx, int P, int* slot, int* nkern, int* noise_slot_out) {
    cudaStream_t s = h->stream;
    int rc;
    if ((rc = compute_Kinv(h))) return rc;
    GradReduceParams gp;
    gp.cp = h->cp;
    gp.X = ptr<double>(h->X); gp.n = ptr<int32_t>(h->n); gp.N = h->N;
    gp.nidx = 0;
    int noise_slot = -1;
    // sigma_f of a single non-SE kernel (k = sigma_f^2 g) without T: from the identity of launch_sigma_identity instead
    // of a pass of dual-number closed forms over M^2 / 2 entries; its value goes behind the entries grad_reduce writes
    const bool sig_id_ok = !h->hasT && (h->kid == GPT_KERNEL_MATERN52 || h->kid == GPT_KERNEL_MATERN || h->kid == GPT_KERNEL_GIBBS_TANH);
    int sig_slot = -1;
    for (int q = 0; q < P; q++) {
        if (grad_idx[q] == h->nparams) noise_slot = q;
        else if (sig_id_ok && grad_idx[q] == 0 && sig_slot < 0) sig_slot = q;
        else { slot[gp.nidx] = q; gp.idx[gp.nidx++] = grad_idx[q]; }
    }
    const int nt = (h->N + 63) / 64;
    if ((rc = ensure(h, h->partials, sizeof(double) * (size_t)nt * nt * GPT_MAX_PARAMS))) return rc;
    if ((rc = ensure(h, h->gout, sizeof(double) * (GPT_MAX_PARAMS + 2)))) return rc;
    gp.partials = ptr<double>(h->partials);
    gp.out = ptr<double>(h->gout);
    if (gp.nidx > 0) {
        if (h->hasT) {
            // W_latent = T' (alpha alpha' - K^-1) T = u u' - T' K^-1 T, u = T' alpha
            const int Np = h->Np, Mp = h->Mp;
            if ((rc = ensure(h, h->u, sizeof(double) * Np))) return rc;
            if ((rc = ensure(h, h->Yt, sizeof(double) * (size_t)Np * Mp))) return rc;
            if ((rc = ensure(h, h->Sg, sizeof(double) * (size_t)Np * Np))) return rc;
            launch_rowdot(ptr<double>(h->Tt), Mp, h->N, h->M, ptr<double>(h->alpha), ptr<double>(h->u), s);
            dim3 sg((Mp + 255) / 256, Mp);
            symmetrize_from_lower_kernel<<<sg, 256, 0, s>>>(ptr<double>(h->Kinv), Mp, Mp);
            GemmParams g;
            g.C = ptr<double>(h->Yt); g.ldc = Mp;
            g.A = ptr<double>(h->Tt); g.lda = Mp;
            g.B = ptr<double>(h->Kinv); g.ldb = Mp;
            g.tiles_m = Np / NB; g.tiles_n = Mp / NB; g.K = Mp;
            g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0; g.kbegin_row = 0;
            launch_gemm_nt(g, s);
            GemmParams g2 = g;
            g2.C = ptr<double>(h->Sg); g2.ldc = Np;
            g2.A = ptr<double>(h->Yt); g2.lda = Mp;
            g2.B = ptr<double>(h->Tt); g2.ldb = Mp;
            g2.tiles_m = Np / NB; g2.tiles_n = Np / NB; g2.K = Mp;
            launch_gemm_nt(g2, s);
            h->launches += 4;
            gp.S = ptr<double>(h->Sg); gp.lds = Np; gp.a = ptr<double>(h->u);
        } else {
            gp.S = ptr<double>(h->Kinv); gp.lds = h->Mp; gp.a = ptr<double>(h->alpha);
        }
        launch_grad_reduce(gp, s);
        h->launches += 2;
    }
    launch_trace_and_sumsq(ptr<double>(h->Kinv), h->Mp, ptr<double>(h->alpha), h->M, ptr<double>(h->gout) + GPT_MAX_PARAMS, s);
    h->launches++;
    if (sig_slot >= 0) {
        launch_sigma_identity(ptr<double>(h->Kinv), h->Mp, ptr<double>(h->alpha), ptr<double>(h->y), ptr<double>(h->diag),
                              h->noise_sigma * h->noise_sigma, h->M, h->cp.p[0], ptr<double>(h->gout) + gp.nidx, s);
        h->launches++;
        slot[gp.nidx++] = sig_slot;
    }
    *nkern = gp.nidx;
    *noise_slot_out = noise_slot;
    return check_launch(h);
}

int gpt_ll(gpt_handle* h, const double* params, double noise_sigma, double* ll, double* grad,
           const int32_t* grad_idx, int P, int* status) {
    if (!h || !params || !ll || !status) return fail(h, GPT_ERR_USAGE, "gpt_ll: bad arguments");
    if (h->M < 1 || h->kid < 0) return fail(h, GPT_ERR_USAGE, "gpt_ll: set_data / set_kernel first");
    if (!supported_kernel_h(h, h->kid, h->D, h->nparams)) return fail(h, GPT_ERR_UNSUPPORTED, "gpt_ll: kernel/dimension unsupported");
    if (grad && P > 0) {
        if (P > GPT_MAX_PARAMS || !grad_idx) return fail(h, GPT_ERR_USAGE, "gpt_ll: bad gradient request");
        for (int q = 0; q < P; q++) {
            if (grad_idx[q] < 0 || grad_idx[q] > h->nparams) return fail(h, GPT_ERR_USAGE, "gpt_ll: bad grad_idx");
            if (hyper_deriv_unavailable(h, h->kid, grad_idx[q]))
                return fail(h, GPT_ERR_UNSUPPORTED, "hyper-parameter derivatives: d/dnu of the Matern kernel is not available");
        }
    }
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    h->factor_valid = false;
    {
        int rcp = make_cov_params(h, h->kid, h->D, h->nparams, params, h->cp, h->comp_dev);
        if (rcp) return rcp;
    }
    h->noise_sigma = noise_sigma;
    int rc;
    if (h->hasT) {
        if ((rc = ensure(h, h->Klat, (size_t)h->Np * h->Np * sizeof(double)))) return rc;
        if ((rc = assemble_train(h, h->cp, noise_sigma, -1, ptr<double>(h->Klat), h->Np, false))) return rc;
        if ((rc = transform_latent(h))) return rc;
    } else {
        if ((rc = ensure(h, h->A, (size_t)h->Mp * h->Mp * sizeof(double)))) return rc;
        if ((rc = assemble_train(h, h->cp, noise_sigma, -1, ptr<double>(h->A), h->Mp, true, true))) return rc;
    }
    if ((rc = factor_and_solve(h, ll, status))) return rc;
    if (!(grad && P > 0)) return 0;
    for (int q = 0; q < P; q++) grad[q] = 0.0;
    if (*status != 0) return 0;

    // ---- gradient: 1/2 tr((alpha alpha^T - K^-1) dK_p), dK tiles regenerated on the fly ----
    int slot[GPT_MAX_PARAMS], nkern = 0, noise_slot = -1;
    if ((rc = gradient_device(h, grad_idx, P, slot, &nkern, &noise_slot))) return rc;
    double hg[GPT_MAX_PARAMS + 2];
    CUDA_OK(h, cudaMemcpyAsync(hg, h->gout.p, sizeof(hg), cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    for (int i = 0; i < nkern; i++) grad[slot[i]] = hg[i];
    if (noise_slot >= 0) {
        // gaussian_process.py:1484-1488: dK = 2 sigma_n I_M  (identity over the observations, also with T)
        grad[noise_slot] = noise_sigma * (hg[GPT_MAX_PARAMS + 1] - hg[GPT_MAX_PARAMS]);
    }
    return 0;
}

int gpt_ll_from_K(gpt_handle* h, const double* K_latent, double* ll, int* status) {
    if (!h || !K_latent || !ll || !status) return fail(h, GPT_ERR_USAGE, "gpt_ll_from_K: bad arguments");
    if (h->M < 1) return fail(h, GPT_ERR_USAGE, "gpt_ll_from_K: set_data first");
    CUDA_OK(h, cudaSetDevice(h->device));
    h->factor_valid = false;
    int rc;
    if (h->hasT) {
        if ((rc = upload_padded(h, h->Klat, K_latent, h->N, h->N, h->Np, h->Np))) return rc;
        if ((rc = transform_latent(h))) return rc;
    } else {
        if ((rc = upload_padded(h, h->A, K_latent, h->N, h->N, h->Mp, h->Mp))) return rc;
        launch_add_diag(ptr<double>(h->A), h->Mp, ptr<double>(h->diag), h->M, h->stream);
        launch_set_identity_pad(ptr<double>(h->A), h->Mp, h->M, h->Mp, h->stream);
        h->launches += 2;
    }
    h->cp.kid = -1;  // predict must be driven with host-supplied K* for plugin kernels
    return factor_and_solve(h, ll, status);
}

int gpt_grad_from_dK(gpt_handle* h, const double* dK_latent, double* g) {
    if (!h || !dK_latent || !g) return fail(h, GPT_ERR_USAGE, "gpt_grad_from_dK: bad arguments");
    if (!h->factor_valid) return fail(h, GPT_ERR_USAGE, "gpt_grad_from_dK: no valid factorisation");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    int rc;
    const int Mp = h->Mp, Np = h->Np;
    if ((rc = compute_Kinv(h))) return rc;
    dim3 sg((Mp + 255) / 256, Mp);
    symmetrize_from_lower_kernel<<<sg, 256, 0, s>>>(ptr<double>(h->Kinv), Mp, Mp);
    const double *W, *a;
    long ldw;
    int n;
    if (h->hasT) {
        // dK_obs = T dK T'  ->  W via two GEMMs into Sg (Mp x Mp reuse of A-shaped buffers)
        if ((rc = upload_padded(h, h->Klat, dK_latent, h->N, h->N, Np, Np))) return rc;
        if ((rc = ensure(h, h->W, (size_t)Mp * Np * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->Sg, (size_t)Mp * Mp * sizeof(double)))) return rc;
        GemmParams g1;
        g1.C = ptr<double>(h->W); g1.ldc = Np;
        g1.A = ptr<double>(h->T); g1.lda = Np;
        g1.B = ptr<double>(h->Klat); g1.ldb = Np;
        g1.tiles_m = Mp / NB; g1.tiles_n = Np / NB; g1.K = Np;
        g1.alpha = 1.0; g1.beta = 0.0; g1.lower_only = 0; g1.kbegin_row = 0;
        launch_gemm_nt(g1, s);
        GemmParams g2 = g1;
        g2.C = ptr<double>(h->Sg); g2.ldc = Mp;
        g2.A = ptr<double>(h->W); g2.lda = Np;
        g2.B = ptr<double>(h->T); g2.ldb = Np;
        g2.tiles_m = Mp / NB; g2.tiles_n = Mp / NB; g2.K = Np;
        launch_gemm_nt(g2, s);
        h->launches += 2;
        W = ptr<double>(h->Sg); ldw = Mp;
    } else {
        if ((rc = upload_padded(h, h->Sg, dK_latent, h->N, h->N, Mp, Mp))) return rc;
        W = ptr<double>(h->Sg); ldw = Mp;
    }
    a = ptr<double>(h->alpha);
    n = h->M;
    if ((rc = ensure(h, h->partials, sizeof(double) * (size_t)Mp))) return rc;
    elem_dot_lower_kernel<<<n, 256, 0, s>>>(ptr<double>(h->Kinv), Mp, a, W, ldw, n, ptr<double>(h->partials));
    h->launches += 2;
    if ((rc = check_launch(h))) return rc;
    std::vector<double> part(n);
    CUDA_OK(h, cudaMemcpyAsync(part.data(), h->partials.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += part[i];
    *g = 0.5 * sum;
    return 0;
}

int gpt_noise_grad(gpt_handle* h, double noise_sigma, double* g) {
    if (!h || !g) return fail(h, GPT_ERR_USAGE, "gpt_noise_grad: bad arguments");
    if (!h->factor_valid) return fail(h, GPT_ERR_USAGE, "gpt_noise_grad: no valid factorisation");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    int rc;
    if ((rc = compute_Kinv(h))) return rc;
    if ((rc = ensure(h, h->gout, sizeof(double) * (GPT_MAX_PARAMS + 2)))) return rc;
    launch_trace_and_sumsq(ptr<double>(h->Kinv), h->Mp, ptr<double>(h->alpha), h->M, ptr<double>(h->gout) + GPT_MAX_PARAMS, s);
    h->launches++;
    if ((rc = check_launch(h))) return rc;
    double hg[2];
    CUDA_OK(h, cudaMemcpyAsync(hg, ptr<double>(h->gout) + GPT_MAX_PARAMS, sizeof(hg), cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    *g = noise_sigma * (hg[1] - hg[0]);  // 1/2 tr((alpha alpha^T - K_tot^{-1}) 2 sigma_n I_M)
    return 0;
}

int gpt_get_alpha(gpt_handle* h, double* alpha) {
    if (!h || !alpha || !h->alpha.p) return fail(h, GPT_ERR_USAGE, "gpt_get_alpha: nothing computed");
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaMemcpyAsync(alpha, h->alpha.p, sizeof(double) * h->M, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int gpt_get_L(gpt_handle* h, double* L) {
    if (!h || !L || !h->A.p) return fail(h, GPT_ERR_USAGE, "gpt_get_L: nothing computed");
    CUDA_OK(h, cudaSetDevice(h->device));
    const int M = h->M;
    CUDA_OK(h, cudaMemcpy2DAsync(L, sizeof(double) * M, h->A.p, sizeof(double) * h->Mp, sizeof(double) * M, M,
                                 cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    for (int r = 0; r < M; r++)
        for (int c = r + 1; c < M; c++) L[(size_t)r * M + c] = 0.0;
    return 0;
}

int gpt_get_K(gpt_handle* h, double* K) {
    if (!h || !K || h->kid < 0 || h->N < 1 || h->cp.kid < 0) return fail(h, GPT_ERR_USAGE, "gpt_get_K: nothing computed");
    CUDA_OK(h, cudaSetDevice(h->device));
    DevBuf tmp;
    int rc = ensure(h, tmp, sizeof(double) * (size_t)h->N * h->N);
    if (rc) return rc;
    rc = assemble_train(h, h->cp, 0.0, -1, ptr<double>(tmp), h->N, false);
    cudaError_t e = cudaSuccess;
    if (!rc) {
        e = cudaMemcpyAsync(K, tmp.p, sizeof(double) * (size_t)h->N * h->N, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    release(tmp);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(h, GPT_ERR_CUDA, "gpt_get_K: %s", cudaGetErrorString(e));
    return 0;
}

// Shared by gpt_predict and gpt_predict_from_Kstar.  On entry h->Kst holds K*^T for `rows` test points (test points
// as rows, ld Np, zero padded to rows_pad).  Applies T (when present), writes the predictive mean for these points
// to h->mean + s0 and, when need_second, overwrites the chunk with V^T = K*^T L^{-T} (block forward substitution,
// every step a DMMA GEMM; gaussian_process.py:983).  *Ko / *ldk: where the (transformed) chunk lives.
static int predict_solve_chunk(gpt_handle* h, int s0, int rows, int rows_pad, bool need_second, double** Ko_out, long* ldk_out) {
    cudaStream_t s = h->stream;
    const int N = h->N, M = h->M, Np = h->Np, Mp = h->Mp, nblk = Mp / NB;
    (void)N;
    double* Ko = ptr<double>(h->Kst);
    long ldk = Np;
    if (h->hasT) {
        GemmParams g;  // (K*^T) T^T
        g.C = ptr<double>(h->Kso); g.ldc = Mp;
        g.A = ptr<double>(h->Kst); g.lda = Np;
        g.B = ptr<double>(h->T); g.ldb = Np;
        g.tiles_m = rows_pad / NB; g.tiles_n = Mp / NB; g.K = Np;
        g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0; g.kbegin_row = 0;
        launch_gemm_nt(g, s);
        h->launches++;
        Ko = ptr<double>(h->Kso);
        ldk = Mp;
    }
    launch_rowdot(Ko, ldk, rows, M, ptr<double>(h->alpha), ptr<double>(h->mean) + s0, s);
    h->launches++;
    if (need_second) {
        for (int I = 0; I < nblk; I++) {
            if (I > 0) {
                GemmParams g;
                g.C = Ko + (long)I * NB; g.ldc = ldk;
                g.A = Ko; g.lda = ldk;
                g.B = ptr<double>(h->A) + (long)I * NB * Mp; g.ldb = Mp;
                g.tiles_m = rows_pad / NB; g.tiles_n = 1; g.K = I * NB;
                g.alpha = -1.0; g.beta = 1.0; g.lower_only = 0; g.kbegin_row = 0;
                launch_gemm_nt(g, s);
                h->launches++;
            }
            // V_I = W_I Inv_I^T, out of place (two CTAs share the rows of W_I: in place would race), then back
            GemmParams g2;
            g2.C = ptr<double>(h->Vtmp); g2.ldc = NB;
            g2.A = Ko + (long)I * NB; g2.lda = ldk;
            g2.B = ptr<double>(h->Inv) + (size_t)I * NB * NB; g2.ldb = NB;
            g2.tiles_m = rows_pad / NB; g2.tiles_n = 1; g2.K = NB;
            g2.alpha = 1.0; g2.beta = 0.0; g2.lower_only = 0; g2.kbegin_row = 0;
            launch_gemm_nt(g2, s);
            launch_copy2d(Ko + (long)I * NB, ldk, ptr<double>(h->Vtmp), NB, rows_pad, NB, s);
            h->launches += 2;
        }
    }
    *Ko_out = Ko;
    *ldk_out = ldk;
    return 0;
}

// Test-point chunk size: the full covariance needs every test point in one chunk; otherwise whole waves of the
// solve GEMMs (an 8192-row chunk filled 43% of the machine), as many as fit a third of the free device memory
// (capped at 48 GB).
static int predict_chunk_rows(gpt_handle* h, int Ms, bool full_cov) {
    int CH = round_up(Ms, NB);
    if (full_cov) return CH;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t budget = free_b / 3;
    if (budget > ((size_t)48 << 30)) budget = (size_t)48 << 30;
    const size_t per_row = (size_t)(h->Np + (h->hasT ? h->Mp : 0) + NB) * sizeof(double);
    const long wave = gemm_rows_per_wave_n128(sms);
    long fit = (long)(budget / per_row);
    if (fit >= wave) fit = fit / wave * wave;
    else fit = fit / NB * NB;
    if (fit < NB) fit = NB;
    if (CH > fit) CH = (int)fit;
    return CH;
}

// Results stay in h->mean / h->var / h->cov (device); the entry points below copy them out.
static int predict_core(gpt_handle* h, int Ms, const double* Xs, const int32_t* ns, bool var, bool cov) {
    if (!h->factor_valid || h->cp.kid < 0) return fail(h, GPT_ERR_USAGE, "gpt_predict: no valid factorisation (call gpt_ll)");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int D = h->D, N = h->N, M = h->M, Np = h->Np, Mp = h->Mp;
    int rc;
    const int CH = predict_chunk_rows(h, Ms, cov);
    if ((rc = upload(h, h->Xs, Xs, sizeof(double) * (size_t)Ms * D))) return rc;
    if ((rc = upload(h, h->ns, ns, sizeof(int32_t) * (size_t)Ms * D))) return rc;
    int max_ns = 0;
    for (size_t i = 0; i < (size_t)Ms * D; i++) max_ns = ns[i] > max_ns ? ns[i] : max_ns;
    const int low_order = (h->max_order <= 1 && max_ns <= 1) ? 1 : 0;  // the short SE / Matern-5/2 tile generators apply
    if (!var && !cov) {
        // mean only: K(X*, X) alpha fused with tile generation, nothing materialised (u = alpha, or T^T alpha)
        const double* u = ptr<double>(h->alpha);
        if (h->hasT) {
            if ((rc = ensure(h, h->u, sizeof(double) * Np))) return rc;
            launch_rowdot(ptr<double>(h->Tt), Mp, N, M, ptr<double>(h->alpha), ptr<double>(h->u), s);
            h->launches++;
            u = ptr<double>(h->u);
        }
        const int nsplit = predict_mean_nsplit(N, Ms);
        if ((rc = ensure(h, h->Kst, sizeof(double) * (size_t)nsplit * Ms))) return rc;
        if ((rc = ensure(h, h->mean, sizeof(double) * (size_t)round_up(Ms, NB)))) return rc;
        launch_predict_mean_fused(h->cp, ptr<double>(h->X), ptr<int32_t>(h->n), u, N, ptr<double>(h->Xs),
                                  ptr<int32_t>(h->ns), Ms, low_order,
                                  ptr<double>(h->Kst), ptr<double>(h->mean), s);
        h->launches += 2;
        return check_launch(h);
    }
    if ((rc = ensure(h, h->Kst, sizeof(double) * (size_t)CH * Np))) return rc;
    if (h->hasT && (rc = ensure(h, h->Kso, sizeof(double) * (size_t)CH * Mp))) return rc;
    if ((rc = ensure(h, h->mean, sizeof(double) * (size_t)round_up(Ms, NB)))) return rc;
    if (var || cov) {
        if ((rc = ensure(h, h->Vtmp, sizeof(double) * (size_t)CH * NB))) return rc;
        if ((rc = ensure(h, h->kss, sizeof(double) * (size_t)round_up(Ms, NB)))) return rc;
        if ((rc = ensure(h, h->var, sizeof(double) * (size_t)round_up(Ms, NB)))) return rc;
    }
    for (int s0 = 0; s0 < Ms; s0 += CH) {
        const int rows = (Ms - s0 < CH) ? (Ms - s0) : CH;
        const int rows_pad = round_up(rows, NB);
        // K*^T (test points as rows): out[s][i] = k(X_i, X*_s; n_i, n*_s)  (gaussian_process.py:966)
        AssembleParams a;
        a.cp = h->cp;
        a.Xr = ptr<double>(h->Xs) + (size_t)s0 * D; a.nr = ptr<int32_t>(h->ns) + (size_t)s0 * D; a.Mr = rows;
        a.Xc = ptr<double>(h->X); a.nc = ptr<int32_t>(h->n); a.Mc = N;
        a.out = ptr<double>(h->Kst); a.ldo = Np; a.rows_pad = rows_pad; a.cols_pad = Np;
        a.hyper_deriv = -1; a.swap_roles = 1; a.symmetric = 0;
        a.diag_add = nullptr; a.diag_const = 0.0; a.pad_identity = 0;
        a.low_order = low_order;
        launch_assemble(a, s);
        h->launches++;
        double* Ko = nullptr;
        long ldk = 0;
        if ((rc = predict_solve_chunk(h, s0, rows, rows_pad, var || cov, &Ko, &ldk))) return rc;
        if (var || cov) {
            launch_prior_diag(h->cp, ptr<double>(h->Xs) + (size_t)s0 * D, ptr<int32_t>(h->ns) + (size_t)s0 * D, rows,
                              ptr<double>(h->kss) + s0, s);
            launch_row_var(Ko, ldk, rows, M, ptr<double>(h->kss) + s0, ptr<double>(h->var) + s0, s);
            h->launches += 2;
            if (cov) {
                const int Sp = rows_pad;
                if ((rc = ensure(h, h->cov, sizeof(double) * (size_t)Sp * Sp))) return rc;
                AssembleParams c;
                c.cp = h->cp;
                c.Xr = ptr<double>(h->Xs); c.nr = ptr<int32_t>(h->ns); c.Mr = Ms;
                c.Xc = c.Xr; c.nc = c.nr; c.Mc = Ms;
                c.out = ptr<double>(h->cov); c.ldo = Sp; c.rows_pad = Sp; c.cols_pad = Sp;
                c.hyper_deriv = -1; c.swap_roles = 0; c.symmetric = 0;
                c.diag_add = nullptr; c.diag_const = 0.0; c.pad_identity = 0;
                c.low_order = (max_ns <= 1) ? 1 : 0;
                launch_assemble(c, s);
                // padding columns (>= M) of V are zero rows of the identity-padded system: K* is zero there
                GemmParams g;
                g.C = ptr<double>(h->cov); g.ldc = Sp;
                g.A = Ko; g.lda = ldk;
                g.B = Ko; g.ldb = ldk;
                g.tiles_m = Sp / NB; g.tiles_n = Sp / NB; g.K = Mp;
                g.alpha = -1.0; g.beta = 1.0; g.lower_only = 0; g.kbegin_row = 0;
                launch_gemm_nt(g, s);
                h->launches += 2;
            }
        }
        if ((rc = check_launch(h))) return rc;
    }
    return 0;
}

int gpt_predict(gpt_handle* h, int Ms, const double* Xs, const int32_t* ns, double* mean, double* var, double* cov) {
    if (h && Ms == 0) return 0;  // empty test set: nothing to write
    if (!h || Ms < 1 || !Xs || !ns || !mean) return fail(h, GPT_ERR_USAGE, "gpt_predict: bad arguments");
    int rc = predict_core(h, Ms, Xs, ns, var != nullptr, cov != nullptr);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    CUDA_OK(h, cudaMemcpyAsync(mean, h->mean.p, sizeof(double) * Ms, cudaMemcpyDeviceToHost, s));
    if (var) CUDA_OK(h, cudaMemcpyAsync(var, h->var.p, sizeof(double) * Ms, cudaMemcpyDeviceToHost, s));
    if (cov) {
        const int Sp = round_up(Ms, NB);
        CUDA_OK(h, cudaMemcpy2DAsync(cov, sizeof(double) * Ms, h->cov.p, sizeof(double) * Sp, sizeof(double) * Ms, Ms,
                                     cudaMemcpyDeviceToHost, s));
    }
    CUDA_OK(h, cudaStreamSynchronize(s));
    return 0;
}

int gpt_predict_dev(gpt_handle* h, int Ms, const double* Xs, const int32_t* ns, double* d_mean, double* d_var) {
    if (h && Ms == 0) return 0;
    if (!h || Ms < 1 || !Xs || !ns || !d_mean) return fail(h, GPT_ERR_USAGE, "gpt_predict_dev: bad arguments");
    int rc = predict_core(h, Ms, Xs, ns, d_var != nullptr, false);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    CUDA_OK(h, cudaMemcpyAsync(d_mean, h->mean.p, sizeof(double) * Ms, cudaMemcpyDeviceToDevice, s));
    if (d_var) CUDA_OK(h, cudaMemcpyAsync(d_var, h->var.p, sizeof(double) * Ms, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int gpt_predict_from_Kstar(gpt_handle* h, int Ms, const double* KstarT, const double* kss_diag, const double* Kss,
                           double* mean, double* var, double* cov) {
    if (h && Ms == 0) return 0;
    if (!h || Ms < 1 || !KstarT || !mean || (var && !kss_diag && !Kss) || (cov && !Kss))
        return fail(h, GPT_ERR_USAGE, "gpt_predict_from_Kstar: bad arguments");
    if (!h->factor_valid) return fail(h, GPT_ERR_USAGE, "gpt_predict_from_Kstar: no valid factorisation (call gpt_ll / gpt_ll_from_K)");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int N = h->N, M = h->M, Np = h->Np, Mp = h->Mp;
    int rc;
    const int CH = predict_chunk_rows(h, Ms, cov != nullptr);
    const int Msp = round_up(Ms, NB);
    if (h->hasT && (rc = ensure(h, h->Kso, sizeof(double) * (size_t)CH * Mp))) return rc;
    if ((rc = ensure(h, h->mean, sizeof(double) * (size_t)Msp))) return rc;
    const bool second = var || cov;
    if (second) {
        if ((rc = ensure(h, h->Vtmp, sizeof(double) * (size_t)CH * NB))) return rc;
        if ((rc = ensure(h, h->var, sizeof(double) * (size_t)Msp))) return rc;
        if ((rc = ensure(h, h->kss, sizeof(double) * (size_t)Msp))) return rc;
        if (kss_diag) {
            CUDA_OK(h, cudaMemcpyAsync(h->kss.p, kss_diag, sizeof(double) * Ms, cudaMemcpyHostToDevice, s));
        } else {  // diagonal of the full prior covariance
            CUDA_OK(h, cudaMemcpy2DAsync(h->kss.p, sizeof(double), Kss, sizeof(double) * ((size_t)Ms + 1), sizeof(double),
                                         Ms, cudaMemcpyHostToDevice, s));
        }
    }
    for (int s0 = 0; s0 < Ms; s0 += CH) {
        const int rows = (Ms - s0 < CH) ? (Ms - s0) : CH;
        const int rows_pad = round_up(rows, NB);
        // the caller's K*^T rows (test points as rows, N latent columns), zero padded into the chunk buffer
        if ((rc = upload_padded(h, h->Kst, KstarT + (size_t)s0 * N, rows, N, rows_pad, Np))) return rc;
        double* Ko = nullptr;
        long ldk = 0;
        if ((rc = predict_solve_chunk(h, s0, rows, rows_pad, second, &Ko, &ldk))) return rc;
        if (second) {
            launch_row_var(Ko, ldk, rows, M, ptr<double>(h->kss) + s0, ptr<double>(h->var) + s0, s);
            h->launches++;
            if (cov) {  // single chunk: cov = K** - V^T V
                const int Sp = rows_pad;
                if ((rc = upload_padded(h, h->cov, Kss, Ms, Ms, Sp, Sp))) return rc;
                GemmParams g;
                g.C = ptr<double>(h->cov); g.ldc = Sp;
                g.A = Ko; g.lda = ldk;
                g.B = Ko; g.ldb = ldk;
                g.tiles_m = Sp / NB; g.tiles_n = Sp / NB; g.K = Mp;
                g.alpha = -1.0; g.beta = 1.0; g.lower_only = 0; g.kbegin_row = 0;
                launch_gemm_nt(g, s);
                h->launches++;
            }
        }
        if ((rc = check_launch(h))) return rc;
        CUDA_OK(h, cudaStreamSynchronize(s));  // the next chunk overwrites the upload buffer
    }
    CUDA_OK(h, cudaMemcpyAsync(mean, h->mean.p, sizeof(double) * Ms, cudaMemcpyDeviceToHost, s));
    if (var) CUDA_OK(h, cudaMemcpyAsync(var, h->var.p, sizeof(double) * Ms, cudaMemcpyDeviceToHost, s));
    if (cov) {
        const int Sp = round_up(Ms, NB);
        CUDA_OK(h, cudaMemcpy2DAsync(cov, sizeof(double) * Ms, h->cov.p, sizeof(double) * Sp, sizeof(double) * Ms, Ms,
                                     cudaMemcpyDeviceToHost, s));
    }
    CUDA_OK(h, cudaStreamSynchronize(s));
    return 0;
}

int gpt_draw_sample(gpt_handle* h, int Ms, int S, const double* mean, const double* cov, const double* rand_vars,
                    double jitter, double* out, int* status) {
    if (!h || Ms < 1 || S < 1 || !mean || !cov || !rand_vars || !out || !status)
        return fail(h, GPT_ERR_USAGE, "gpt_draw_sample: bad arguments");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int Sp = round_up(Ms, NB), nblk = Sp / NB, Rp = round_up(S, NB);
    int rc;
    // scratch lives in the handle (grown on demand, freed by gpt_destroy): cudaMalloc / cudaFree per call cost more than
    // the factorisation itself at config-5 sizes
    DevBuf &C = h->ds_C, &inv = h->ds_inv, &panel = h->ds_panel, &logdet = h->ds_logdet, &info = h->ds_info, &R = h->ds_R,
           &Rt = h->ds_Rt, &O = h->ds_O, &mu = h->ds_mu, &jit = h->ds_jit;
    auto cleanup = [&]() {};
    std::vector<double> hj(Ms, jitter);
    if ((rc = upload_padded(h, C, cov, Ms, Ms, Sp, Sp)) || (rc = upload(h, jit, hj.data(), sizeof(double) * Ms)) ||
        (rc = ensure(h, inv, sizeof(double) * (size_t)nblk * NB * NB)) ||
        (rc = ensure(h, panel, sizeof(double) * potrf_panel_doubles(nblk))) ||
        (rc = ensure(h, logdet, sizeof(double) * nblk)) ||
        (rc = ensure(h, info, sizeof(int))) || (rc = upload_padded(h, R, rand_vars, Ms, S, Sp, Rp)) ||
        (rc = ensure(h, Rt, sizeof(double) * (size_t)Rp * Sp)) || (rc = ensure(h, O, sizeof(double) * (size_t)Sp * Rp)) ||
        (rc = upload(h, mu, mean, sizeof(double) * Ms))) {
        cleanup();
        return rc;
    }
    launch_add_diag(ptr<double>(C), Sp, ptr<double>(jit), Ms, s);
    launch_set_identity_pad(ptr<double>(C), Sp, Ms, Sp, s);
    rc = blocked_potrf(h, ptr<double>(C), Sp, nblk, ptr<double>(inv), ptr<double>(panel), nullptr, ptr<double>(logdet),
                       ptr<int>(info));
    if (!rc) {
        dim3 tg((Sp + 255) / 256, Sp);
        tril_kernel<<<tg, 256, 0, s>>>(ptr<double>(C), Sp, Sp);
        cudaMemsetAsync(Rt.p, 0, sizeof(double) * (size_t)Rp * Sp, s);
        launch_transpose(ptr<double>(Rt), Sp, ptr<double>(R), Rp, Ms, S, s);
        GemmParams g;
        g.C = ptr<double>(O); g.ldc = Rp;
        g.A = ptr<double>(C); g.lda = Sp;
        g.B = ptr<double>(Rt); g.ldb = Sp;
        g.tiles_m = Sp / NB; g.tiles_n = Rp / NB; g.K = Sp;
        g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0; g.kbegin_row = 0;
        launch_gemm_nt(g, s);
        dim3 mg((S + 255) / 256, Ms);
        add_mean_kernel<<<mg, 256, 0, s>>>(ptr<double>(O), Rp, ptr<double>(mu), Ms, S);
        h->launches += 6;
        rc = check_launch(h);
    }
    cudaError_t e = cudaSuccess;
    int hinfo = 0;
    if (!rc) {
        e = cudaMemcpy2DAsync(out, sizeof(double) * S, O.p, sizeof(double) * Rp, sizeof(double) * S, Ms,
                              cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    cleanup();
    if (rc) return rc;
    if (e != cudaSuccess) return fail(h, GPT_ERR_CUDA, "gpt_draw_sample: %s", cudaGetErrorString(e));
    *status = hinfo;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched many-theta path
// ------------------------------------------------------------------------------------------------
// ll / gradient scalars of one theta from the handle's scratch into row b of the batch outputs
__global__ void batch_row_scatter_kernel(const double* __restrict__ llred, const int* __restrict__ info, int M,
                                         const double* __restrict__ gout, int P, int nkern, int noise_slot,
                                         double noise_sigma, double* __restrict__ ll, int* __restrict__ status,
                                         double* __restrict__ grad, int b, GradSlots slots) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int inf = info[0];
    if (inf > M) inf = 0;
    ll[b] = llred[0];
    status[b] = inf;
    if (grad != nullptr) {
        for (int q = 0; q < P; q++) grad[(size_t)b * P + q] = 0.0;
        if (inf == 0) {
            for (int i = 0; i < nkern; i++) grad[(size_t)b * P + slots.s[i]] = gout[i];
            if (noise_slot >= 0) grad[(size_t)b * P + noise_slot] = noise_sigma * (gout[GPT_MAX_PARAMS + 1] - gout[GPT_MAX_PARAMS]);
        }
    }
}

// Batches the persistent kernel does not take -- a transformation matrix T (the T K T^T products are GEMMs that fill the
// device on their own) or more than 2048 observations: the thetas run one after the other through the single-matrix
// path (assembly, [T K T^T,] blocked Cholesky, solves, inverse + trace reduction) on the handle's stream, every scalar
// stays on the device and lands in row b of the outputs, and nothing synchronises between thetas.
static int batched_serial(gpt_handle* h, int B, const double* d_thetas, const double* d_y, double* d_ll, double* d_grad,
                          const int32_t* grad_idx, int P, int* d_status, double* d_alpha) {
    cudaStream_t s = h->stream;
    const int np1 = h->nparams + 1, M = h->M;
    const bool want_grad = (d_grad && P > 0);
    if (want_grad) {
        if (P > GPT_MAX_PARAMS || !grad_idx) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched: bad gradient request");
        for (int q = 0; q < P; q++) {
            if (grad_idx[q] < 0 || grad_idx[q] > h->nparams) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched: bad grad_idx");
            if (hyper_deriv_unavailable(h, h->kid, grad_idx[q]))
                return fail(h, GPT_ERR_UNSUPPORTED, "hyper-parameter derivatives: d/dnu of the Matern kernel is not available");
        }
    }
    std::vector<double> th((size_t)B * np1);   // CovParams travel by value in the launches
    CUDA_OK(h, cudaMemcpyAsync(th.data(), d_thetas, sizeof(double) * th.size(), cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    int rc;
    DevBuf ykeep;
    if (d_y) {  // per-theta right-hand sides: h->y is restored afterwards
        if ((rc = ensure(h, ykeep, sizeof(double) * M))) return rc;
        CUDA_OK(h, cudaMemcpyAsync(ykeep.p, h->y.p, sizeof(double) * M, cudaMemcpyDeviceToDevice, s));
    }
    h->factor_valid = false;
    for (int b = 0; b < B && !rc; b++) {
        const double* tb = th.data() + (size_t)b * np1;
        if ((rc = make_cov_params(h, h->kid, h->D, h->nparams, tb, h->cp, h->comp_dev, b, B))) break;  // one slot per theta
        h->noise_sigma = tb[h->nparams];
        if (d_y) CUDA_OK(h, cudaMemcpyAsync(h->y.p, d_y + (size_t)b * M, sizeof(double) * M, cudaMemcpyDeviceToDevice, s));
        if (h->hasT) {
            if ((rc = ensure(h, h->Klat, (size_t)h->Np * h->Np * sizeof(double)))) break;
            if ((rc = assemble_train(h, h->cp, h->noise_sigma, -1, ptr<double>(h->Klat), h->Np, false))) break;
            if ((rc = transform_latent(h))) break;
        } else {
            if ((rc = ensure(h, h->A, (size_t)h->Mp * h->Mp * sizeof(double)))) break;
            if ((rc = assemble_train(h, h->cp, h->noise_sigma, -1, ptr<double>(h->A), h->Mp, true, true))) break;
        }
        if ((rc = factor_and_solve_device(h))) break;
        if (d_alpha) CUDA_OK(h, cudaMemcpyAsync(d_alpha + (size_t)b * M, h->alpha.p, sizeof(double) * M, cudaMemcpyDeviceToDevice, s));
        GradSlots slots;
        int nkern = 0, noise_slot = -1;
        if (want_grad) {
            if ((rc = gradient_device(h, grad_idx, P, slots.s, &nkern, &noise_slot))) break;
        } else if ((rc = ensure(h, h->gout, sizeof(double) * (GPT_MAX_PARAMS + 2)))) {
            break;
        }
        batch_row_scatter_kernel<<<1, 32, 0, s>>>(ptr<double>(h->llred), ptr<int>(h->info), M, ptr<double>(h->gout),
                                                  want_grad ? P : 0, nkern, noise_slot, h->noise_sigma, d_ll, d_status,
                                                  want_grad ? d_grad : nullptr, b, slots);
        h->launches++;
    }
    if (d_y) {
        cudaMemcpyAsync(h->y.p, ykeep.p, sizeof(double) * M, cudaMemcpyDeviceToDevice, s);
        cudaStreamSynchronize(s);
        release(ykeep);
    }
    if (rc) return rc;
    return check_launch(h);
}

// Test points of a prediction batch (already appended to the extended point arrays h->b_Xext / h->b_next)
struct PredictBatch {
    int Ms = 0;
    int max_order = 0;
    double* d_mean = nullptr;
    double* d_var = nullptr;
};

static int batched_common(gpt_handle* h, int B, const double* d_thetas, const double* d_y, double* d_ll,
                          double* d_grad, const int32_t* grad_idx, int P, int* d_status, double* d_alpha,
                          const PredictBatch* pb = nullptr) {
    if (pb && (h->hasT || (h->M + 63) / 64 > 32 || h->kid == GPT_KERNEL_GIBBS_AUX || (d_grad && P > 0)))
        return fail(h, GPT_ERR_UNSUPPORTED, "gpt_predict_batched: needs the persistent many-theta kernel (no T, at most 2048 observations, no gradient request)");
    // composite kernels whose requested gradient entries fit the persistent kernel's slots run there as well
    bool comp_serial = false;
    if (h->kid == GPT_COMPOSITE && d_grad && P > 0 && grad_idx)
        for (int q = 0; q < P; q++) comp_serial |= (grad_idx[q] < h->nparams && grad_idx[q] >= 1 + GPT_MAX_DIM);
    if (h->hasT || (h->M + 63) / 64 > 32 || comp_serial) {
        if (h->kid == GPT_KERNEL_GIBBS_AUX)
            return fail(h, GPT_ERR_UNSUPPORTED, "gpt_ll_batched: the per-point length scales of GPT_GIBBS_AUX depend on theta; use gpt_ll");
        if (!supported_kernel_h(h, h->kid, h->D, h->nparams)) return fail(h, GPT_ERR_UNSUPPORTED, "gpt_ll_batched: kernel unsupported");
        return batched_serial(h, B, d_thetas, d_y, d_ll, d_grad, grad_idx, P, d_status, d_alpha);
    }
    if (h->kid == GPT_KERNEL_GIBBS_AUX)
        return fail(h, GPT_ERR_UNSUPPORTED, "gpt_ll_batched: the per-point length scales of GPT_GIBBS_AUX depend on theta; use gpt_ll");
    if (!supported_kernel_h(h, h->kid, h->D, h->nparams)) return fail(h, GPT_ERR_UNSUPPORTED, "gpt_ll_batched: kernel unsupported");
    if (P < 0 || P > GPT_MAX_PARAMS) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched: bad P");
    BatchedParams bp;
    bp.kid = h->kid; bp.D = h->D; bp.nparams = h->nparams;
    bp.M = h->M; bp.nT = (h->M + 63) / 64;
    bp.X = ptr<double>(h->X); bp.n = ptr<int32_t>(h->n);
    bp.low_order = (h->max_order <= 1 && (!pb || pb->max_order <= 1)) ? 1 : 0;
    if (pb) {
        bp.X = ptr<double>(h->b_Xext); bp.n = ptr<int32_t>(h->b_next);
        bp.Ms = pb->Ms; bp.nTs = (pb->Ms + 63) / 64;
        bp.pmean = pb->d_mean; bp.pvar = pb->d_var;
    }
    bp.y = d_y ? d_y : ptr<double>(h->y); bp.y_stride = d_y ? h->M : 0;
    bp.diag = ptr<double>(h->diag);
    bp.B = B; bp.thetas = d_thetas;
    bp.nidx = (d_grad && P > 0) ? P : 0;
    for (int q = 0; q < bp.nidx; q++) {
        if (grad_idx[q] < 0 || grad_idx[q] > h->nparams) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched: bad grad_idx");
        if (hyper_deriv_unavailable(h, h->kid, grad_idx[q]))
            return fail(h, GPT_ERR_UNSUPPORTED, "hyper-parameter derivatives: d/dnu of the Matern kernel is not available");
        if (h->kid != GPT_KERNEL_SE && grad_idx[q] < h->nparams && grad_idx[q] >= 1 + GPT_MAX_DIM)
            return fail(h, GPT_ERR_UNSUPPORTED, "gpt_ll_batched: hyper-parameter index beyond the batched gradient slots");
        bp.idx[q] = grad_idx[q];
    }
    bp.ll = d_ll; bp.grad = d_grad; bp.status = d_status; bp.alpha_out = d_alpha;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    int ctas = sms * batched4_ctas_per_sm();
    if (ctas > B) ctas = B;
    bp.ws_per_cta = batched_ws_doubles_per_cta(bp.nT);
    int rc;
    // SE in one or two dimensions with derivative orders <= 1: short closed forms; phase 1 leaves sigma^2 exp(-r^2/2)
    // of every lower tile behind in the CTA workspace for the gradient pass
    if (h->kid == GPT_KERNEL_SE && h->D <= 2 && bp.low_order && !getenv("GPT_B4_LONG_FORMS")) {
        bp.short_forms = 1;
        if (bp.nidx > 0) {
            bp.eb_off = bp.ws_per_cta;
            bp.ws_per_cta += batched_lower_tiles(bp.nT) * 4096;
        }
    }
    if (pb) {
        bp.ts_off = bp.ws_per_cta;
        bp.ws_per_cta += (size_t)bp.nTs * bp.nT * 4096;
        bp.pv_off = bp.ws_per_cta;
        bp.ws_per_cta += (size_t)2 * bp.nTs * 64;
        if (sizeof(double) * bp.ws_per_cta * (size_t)ctas > ((size_t)24 << 30))
            return fail(h, GPT_ERR_UNSUPPORTED, "gpt_predict_batched: too many test points for one call (workspace > 24 GB); split them");
    }
    if (h->kid == GPT_COMPOSITE) {
        bp.comp_nleaf = h->comp_nleaf; bp.comp_nterms = h->comp_nterms;
        for (int q = 0; q < GPT_MAX_LEAVES; q++) { bp.comp_kids[q] = h->comp_kids[q]; bp.comp_nps[q] = h->comp_nps[q]; }
        for (int t = 0; t < GPT_MAX_TERMS; t++) bp.comp_masks[t] = h->comp_masks[t];
        bp.comp_off = bp.ws_per_cta;
        bp.ws_per_cta += (sizeof(CovComposite) + sizeof(double) - 1) / sizeof(double);
        bp.ws_per_cta = (bp.ws_per_cta + 15) / 16 * 16;  // keep the next CTA's tiles 128-byte aligned
    }
    if ((rc = ensure(h, h->b_ws, sizeof(double) * bp.ws_per_cta * (size_t)ctas))) return rc;
    if ((rc = ensure(h, h->b_counter, sizeof(int)))) return rc;
    bp.workspace = ptr<double>(h->b_ws);
    bp.counter = ptr<int>(h->b_counter);
    bp.phase_cycles = nullptr;
#ifdef GPT_PHASE_TIMING
    if ((rc = ensure(h, h->scal, 8 * sizeof(long long)))) return rc;
    CUDA_OK(h, cudaMemsetAsync(h->scal.p, 0, 8 * sizeof(long long), h->stream));
    bp.phase_cycles = ptr<long long>(h->scal);
#endif
    CUDA_OK(h, cudaMemsetAsync(bp.counter, 0, sizeof(int), h->stream));
    launch_ll_batched4(bp, ctas, h->stream);
    h->launches++;
#ifdef GPT_PHASE_TIMING
    {
        long long pc[8];
        cudaMemcpyAsync(pc, h->scal.p, sizeof(pc), cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        static const char* names[8] = {"gemm jobs", "K generation", "potrf+inv", "panel products", "residual/z", "backsolve", "gradient", "other"};
        long long tot = 0;
        for (int q = 0; q < 8; q++) tot += pc[q];
        fprintf(stderr, "[phase timing] B=%d, CTA-cycles per theta (thread 0 of each CTA):\n", B);
        for (int q = 0; q < 8; q++) fprintf(stderr, "   %-16s %10.0f  %5.1f%%\n", names[q], (double)pc[q] / B, 100.0 * pc[q] / (double)tot);
        fprintf(stderr, "   %-16s %10.0f\n", "total", (double)tot / B);
    }
#endif
    return check_launch(h);
}

int gpt_ll_batched_dev(gpt_handle* h, int B, const double* d_thetas, const double* d_y_batch, double* d_ll,
                       double* d_grad, const int32_t* grad_idx, int P, int* d_status, double* d_alpha_out) {
    if (h && B == 0) return 0;  // empty batch
    if (!h || B < 1 || !d_thetas || !d_ll || !d_status) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched_dev: bad arguments");
    if (h->M < 1 || h->kid < 0) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched: set_data / set_kernel first");
    CUDA_OK(h, cudaSetDevice(h->device));
    return batched_common(h, B, d_thetas, d_y_batch, d_ll, d_grad, grad_idx, P, d_status, d_alpha_out);
}

int gpt_predict_batched(gpt_handle* h, int B, const double* thetas, const double* y_batch, int Ms, const double* Xs,
                        const int32_t* ns, double* mean, double* var, double* ll, int* status) {
    if (h && (B == 0 || Ms == 0)) return 0;  // empty batch / no test points
    if (!h || B < 1 || Ms < 1 || !thetas || !Xs || !ns || !mean || !var || !status)
        return fail(h, GPT_ERR_USAGE, "gpt_predict_batched: bad arguments");
    if (h->M < 1 || h->kid < 0) return fail(h, GPT_ERR_USAGE, "gpt_predict_batched: set_data / set_kernel first");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int np1 = h->nparams + 1, M = h->M, D = h->D;
    const int rows_train = (M + 63) / 64 * 64;
    int rc;
    // extended point arrays: the training rows padded to whole tiles, then the test points
    const size_t next = (size_t)rows_train + Ms;
    if ((rc = ensure(h, h->b_Xext, sizeof(double) * next * D))) return rc;
    if ((rc = ensure(h, h->b_next, sizeof(int32_t) * next * D))) return rc;
    CUDA_OK(h, cudaMemsetAsync(h->b_Xext.p, 0, sizeof(double) * next * D, s));
    CUDA_OK(h, cudaMemsetAsync(h->b_next.p, 0, sizeof(int32_t) * next * D, s));
    CUDA_OK(h, cudaMemcpyAsync(h->b_Xext.p, h->X.p, sizeof(double) * (size_t)M * D, cudaMemcpyDeviceToDevice, s));
    CUDA_OK(h, cudaMemcpyAsync(h->b_next.p, h->n.p, sizeof(int32_t) * (size_t)M * D, cudaMemcpyDeviceToDevice, s));
    CUDA_OK(h, cudaMemcpyAsync(ptr<double>(h->b_Xext) + (size_t)rows_train * D, Xs, sizeof(double) * (size_t)Ms * D,
                               cudaMemcpyHostToDevice, s));
    CUDA_OK(h, cudaMemcpyAsync(ptr<int32_t>(h->b_next) + (size_t)rows_train * D, ns, sizeof(int32_t) * (size_t)Ms * D,
                               cudaMemcpyHostToDevice, s));
    PredictBatch pb;
    pb.Ms = Ms;
    for (size_t i = 0; i < (size_t)Ms * D; i++) {
        if (ns[i] < 0) return fail(h, GPT_ERR_USAGE, "gpt_predict_batched: negative derivative order");
        if (ns[i] > pb.max_order) pb.max_order = ns[i];
    }
    if ((rc = upload(h, h->b_thetas, thetas, sizeof(double) * (size_t)B * np1))) return rc;
    if (y_batch && (rc = upload(h, h->b_y, y_batch, sizeof(double) * (size_t)B * M))) return rc;
    if ((rc = ensure(h, h->b_ll, sizeof(double) * B))) return rc;
    if ((rc = ensure(h, h->b_status, sizeof(int) * B))) return rc;
    if ((rc = ensure(h, h->b_pmean, sizeof(double) * (size_t)B * Ms))) return rc;
    if ((rc = ensure(h, h->b_pvar, sizeof(double) * (size_t)B * Ms))) return rc;
    pb.d_mean = ptr<double>(h->b_pmean);
    pb.d_var = ptr<double>(h->b_pvar);
    rc = batched_common(h, B, ptr<double>(h->b_thetas), y_batch ? ptr<double>(h->b_y) : nullptr, ptr<double>(h->b_ll),
                        nullptr, nullptr, 0, ptr<int>(h->b_status), nullptr, &pb);
    if (rc) return rc;
    CUDA_OK(h, cudaMemcpyAsync(mean, h->b_pmean.p, sizeof(double) * (size_t)B * Ms, cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaMemcpyAsync(var, h->b_pvar.p, sizeof(double) * (size_t)B * Ms, cudaMemcpyDeviceToHost, s));
    if (ll) CUDA_OK(h, cudaMemcpyAsync(ll, h->b_ll.p, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaMemcpyAsync(status, h->b_status.p, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    return 0;
}

int gpt_ll_batched(gpt_handle* h, int B, const double* thetas, const double* y_batch, double* ll, double* grad,
                   const int32_t* grad_idx, int P, int* status, double* alpha_out) {
    if (h && B == 0) return 0;  // empty batch
    if (!h || B < 1 || !thetas || !ll || !status) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched: bad arguments");
    if (h->M < 1 || h->kid < 0) return fail(h, GPT_ERR_USAGE, "gpt_ll_batched: set_data / set_kernel first");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int np1 = h->nparams + 1, M = h->M;
    const bool want_grad = (grad && P > 0);
    int rc;
    if ((rc = upload(h, h->b_thetas, thetas, sizeof(double) * (size_t)B * np1))) return rc;
    if (y_batch && (rc = upload(h, h->b_y, y_batch, sizeof(double) * (size_t)B * M))) return rc;
    if ((rc = ensure(h, h->b_ll, sizeof(double) * B))) return rc;
    if ((rc = ensure(h, h->b_status, sizeof(int) * B))) return rc;
    if (want_grad && (rc = ensure(h, h->b_grad, sizeof(double) * (size_t)B * P))) return rc;
    if (alpha_out && (rc = ensure(h, h->b_alpha, sizeof(double) * (size_t)B * M))) return rc;
    rc = batched_common(h, B, ptr<double>(h->b_thetas), y_batch ? ptr<double>(h->b_y) : nullptr, ptr<double>(h->b_ll),
                        want_grad ? ptr<double>(h->b_grad) : nullptr, grad_idx, P, ptr<int>(h->b_status),
                        alpha_out ? ptr<double>(h->b_alpha) : nullptr);
    if (rc) return rc;
    CUDA_OK(h, cudaMemcpyAsync(ll, h->b_ll.p, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaMemcpyAsync(status, h->b_status.p, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    if (want_grad) CUDA_OK(h, cudaMemcpyAsync(grad, h->b_grad.p, sizeof(double) * (size_t)B * P, cudaMemcpyDeviceToHost, s));
    if (alpha_out) CUDA_OK(h, cudaMemcpyAsync(alpha_out, h->b_alpha.p, sizeof(double) * (size_t)B * M, cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"
