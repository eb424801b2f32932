// Register-resident squared-exponential evaluators shared by the batched kernel (batched4.cu) and the fused predictive mean (predict.cu),
// plus the per-phase cycle-accounting macros of the development build (-DGPT_PHASE_TIMING).
#pragma once
#include "common.cuh"
#include "internal.h"

#ifdef GPT_PHASE_TIMING
#define PT_DECL long long pt_t0 = clock64(), pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PT_MARK(slot)                          \
    do {                                       \
        const long long pt_t1 = clock64();     \
        pt_acc[slot] += pt_t1 - pt_t0;         \
        pt_t0 = pt_t1;                         \
    } while (0)
#else
#define PT_DECL
#define PT_MARK(slot)
#endif
// slots: 0 GEMM jobs, 1 K-tile generation, 2 potrf+inverse of the diagonal tile, 3 panel products + stores,
//        4 residual / z, 5 back substitution, 6 gradient contraction, 7 everything else

namespace sefast {

// ---- register-resident squared-exponential evaluation, input dimension known at compile time ----------
// (the generic cov_eval keeps per-dimension arrays in local memory and re-reads the parameters from shared
// memory for every entry; ncu showed ~60% of all warp samples there, profiles/r01a_*)
template <int D>
struct SEHoist {
    double sig2, sig;
    double il[D];
    const double* etab;  // 2^(j/64) table in shared memory for the TAB variants of the low-order evaluators, or null
};

template <int D>
__device__ __forceinline__ SEHoist<D> se_hoist(const CovParams& cp) {
    SEHoist<D> h;
    h.sig2 = cp.sig2;
    h.sig = cp.p[0];
#pragma unroll
    for (int d = 0; d < D; d++) h.il[d] = cp.inv_l[d];
    h.etab = nullptr;
    return h;
}

template <int D>
struct PointReg {
    double x[D];
    int n[D];
};

template <int D>
__device__ __forceinline__ PointReg<D> load_point(const double* __restrict__ X, const int32_t* __restrict__ n, int gi) {
    PointReg<D> q;
#pragma unroll
    for (int d = 0; d < D; d++) {
        q.x[d] = __ldg(X + (size_t)gi * D + d);
        q.n[d] = __ldg(n + (size_t)gi * D + d);
    }
    return q;
}

// value only
template <int D>
__device__ __forceinline__ double se_value(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b) {
    double r2 = 0.0, prod = 1.0;
    int sj = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        double tl = tau * h.il[d];
        if (tau == 0.0) tl = 0.0;
        r2 += tl * tl;
        sj += b.n[d];
        double f, g;
        se_dim_factor(tau, h.il[d], a.n[d] + b.n[d], false, f, g);
        prod *= f;
    }
    double k = h.sig2 * exp_nonpos(-0.5 * r2) * prod;
    return (sj & 1) ? -k : k;
}

// value K and dK/dl_d for every dimension (dK/dsigma = 2K/sigma is formed by the caller)
template <int D>
__device__ __forceinline__ void se_value_grad(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b,
                                              double& K, double (&dl)[D]) {
    double r2 = 0.0;
    int sj = 0;
    double f[D], g[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        double tl = tau * h.il[d];
        if (tau == 0.0) tl = 0.0;
        r2 += tl * tl;
        sj += b.n[d];
        se_dim_factor(tau, h.il[d], a.n[d] + b.n[d], true, f[d], g[d]);
    }
    double base = h.sig2 * exp_nonpos(-0.5 * r2);
    if (sj & 1) base = -base;
    double prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; d++) prod *= f[d];
    K = base * prod;
#pragma unroll
    for (int d = 0; d < D; d++) {
        double pr = g[d];
#pragma unroll
        for (int e = 0; e < D; e++)
            if (e != d) pr *= f[e];
        dl[d] = base * pr;
    }
}

// Branch-free versions for derivative orders <= 1 on both sides (m_d <= 2): selects only, so that the compiler
// can interleave the unrolled evaluations (a conditional branch per entry serialises them -- measured).
template <bool TAB, int D>
__device__ __forceinline__ double exp_low(const SEHoist<D>& h, double x) {
    if constexpr (TAB) return exp_nonpos_tab(x, h.etab);
    else return exp_nonpos_nobranch(x);
}

template <int D, bool TAB = false>
__device__ __forceinline__ double se_value_low(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b) {
    double r2 = 0.0, prod = 1.0;
    int sj = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        const double tl = (tau == 0.0) ? 0.0 : tau * h.il[d];
        r2 = fma(tl, tl, r2);
        sj += b.n[d];
        prod *= se_dim_value_low(tau, h.il[d], a.n[d] + b.n[d]);
    }
    const double k = h.sig2 * exp_low<TAB>(h, -0.5 * r2) * prod;
    return (sj & 1) ? -k : k;
}

template <int D, bool TAB = false>
__device__ __forceinline__ void se_value_grad_low(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b,
                                                  double& K, double (&dl)[D]) {
    double r2 = 0.0;
    int sj = 0;
    double f[D], g[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        const double tl = (tau == 0.0) ? 0.0 : tau * h.il[d];
        r2 = fma(tl, tl, r2);
        sj += b.n[d];
        se_dim_factor_low(tau, h.il[d], a.n[d] + b.n[d], f[d], g[d]);
    }
    double base = h.sig2 * exp_low<TAB>(h, -0.5 * r2);
    base = (sj & 1) ? -base : base;
    double prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; d++) prod *= f[d];
    K = base * prod;
#pragma unroll
    for (int d = 0; d < D; d++) {
        double pr = g[d];
#pragma unroll
        for (int e = 0; e < D; e++)
            if (e != d) pr *= f[e];
        dl[d] = base * pr;
    }
}

// ---- Matern 5/2, at most one first derivative per point (the Matern52Kernel contract, kernel/matern.py:545-546) -----
// k(a, b) with r2 = sum tau_d^2 / l_d^2, s = sqrt(5 r2), tau = x_a - x_b (reference kernel/src/matern.c:61-186):
//   value  sig2 (1 + s + 5/3 r2) e^-s;   d/dx_a,i: -5/3 sig2 (1 + s) e^-s tau_i / l_i^2;   d/dx_b,j: the same with +;
//   both:  5/3 sig2 e^-s ((1 + s) [i == j] / l_i^2 - 5 tau_i tau_j / (l_i^2 l_j^2)).   Branch-free (selects), table exp.
template <int D>
struct M52Hoist {
    double sig2, c53;
    double il2[D];
    const double* etab;
};

template <int D>
__device__ __forceinline__ M52Hoist<D> m52_hoist(const CovParams& cp, const double* etab) {
    M52Hoist<D> h;
    h.sig2 = cp.sig2;
    h.c53 = 1.6666666666666667 * cp.sig2;
#pragma unroll
    for (int d = 0; d < D; d++) h.il2[d] = cp.inv_l[d] * cp.inv_l[d];
    h.etab = etab;
    return h;
}

template <int D>
__device__ __forceinline__ double m52_value_low(const M52Hoist<D>& h, const PointReg<D>& a, const PointReg<D>& b) {
    double r2 = 0.0, ta = 0.0, ia = 0.0, tb = 0.0, ib = 0.0;
    int da = -1, db = -1;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        r2 = fma(tau * tau, h.il2[d], r2);
        const bool oa = a.n[d] != 0, ob = b.n[d] != 0;
        da = oa ? d : da;
        db = ob ? d : db;
        ta = oa ? tau : ta;
        ia = oa ? h.il2[d] : ia;
        tb = ob ? tau : tb;
        ib = ob ? h.il2[d] : ib;
    }
    const double q = 5.0 * r2;
    const double s = (q > 0.0) ? q * fast_rsqrt_pos(q) : 0.0;
    const double e = exp_nonpos_tab(-s, h.etab);
    const double Bf = (1.0 + s) * e;
    const double v00 = h.sig2 * fma(1.6666666666666667, r2, 1.0 + s) * e;
    const double v10 = -h.c53 * Bf * ta * ia;
    const double v01 = h.c53 * Bf * tb * ib;
    const double v11 = h.c53 * (((da == db) ? Bf * ia : 0.0) - 5.0 * e * (ta * ia) * (tb * ib));
    return (da < 0) ? ((db < 0) ? v00 : v01) : ((db < 0) ? v10 : v11);
}

}  // namespace sefast
