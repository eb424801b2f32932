// Pairwise covariance closed forms with derivative orders -- the bodies of the reference's
// Kernel.__call__ plugins, written as scalar __host__ __device__ functions so that the very
// same source is (a) inlined into the tile-generating CUDA kernels and (b) compiled for the
// host by csrc/hostmath.cpp and checked against the golden vectors without a GPU.
//
// Reference semantics followed (paths under /root/reference/gptools):
//   SE          kernel/squared_exponential.py:110-174, kernel/core.py:384-421
//   MATERN52    kernel/matern.py:543-555, kernel/src/matern.c:61-186
//   MATERN      kernel/matern.py:296-459, kernel/core.py:728-750, utils.py:1429-1518
//               (half-integer nu, total derivative order <= 2, incl. the 0 < y <= 5e-4 series zone)
//   GIBBS_TANH  kernel/gibbs.py:324-423, 458-461
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GPT_HD __host__ __device__ __forceinline__
#else
#define GPT_HD inline
#endif

#define GPT_KERNEL_SE 0
#define GPT_KERNEL_MATERN52 1
#define GPT_KERNEL_MATERN 2
#define GPT_KERNEL_GIBBS_TANH 3
#define GPT_KERNEL_GIBBS_AUX 4  // Gibbs kernel, length scale l(x) and l'(x) supplied per point in columns 1, 2

#define GPT_KERNEL_COMPOSITE 5  // sum of products of up to GPT_MAX_LEAVES of the kernels above (kernel/core.py:424-670)

#define GPT_MAX_DIM 6
#define GPT_MAX_PARAMS 10
#define GPT_MAX_LEAVES 4
#define GPT_MAX_TERMS 8

struct CovComposite;

// Kernel hyperparameters in the reference's order plus quantities derived once per theta.
struct CovParams {
    int kid;
    int D;
    int nparams;
    double p[GPT_MAX_PARAMS];  // the reference's params vector
    // derived
    double sig2;               // sigma_f^2
    double inv_l[GPT_MAX_DIM]; // 1 / l_d
    // generic Matern: derivative-series constants, utils.py:1496-1516
    int matern_p;              // floor(nu)
    int mat_kind;              // 0: nu = p + 1/2 (closed form K_nu), 1: any other non-integer nu, 2: integer nu
    double mat_c;              // 2^{1-nu} / Gamma(nu)
    double mat_nu[2];          // the order(s) the series zone is evaluated at: {nu, nu}; integer nu: {nu - 0.001, nu + 0.001}
    double mat_A[2][4];        // [side][n]: c * Gamma(nu') / (2^{1-nu'+2n} (1-nu')_n)       (x 1/2 for integer nu)
    double mat_B[2][4];        // [side][n]: c * Gamma(-nu') (1+nu'-n)_n / 2^{1+nu'}         (x 1/2 for integer nu)
    // kid == GPT_KERNEL_COMPOSITE: the operand kernels with their parameters (device memory inside the CUDA kernels,
    // host memory in the host build); p[] then holds the concatenated vector for reference only
    const CovComposite* comp;
};

// SumKernel / ProductKernel trees of the reference (kernel/core.py:549-670), flattened by the host into a sum of
// products: k = sum_t prod_{q in mask[t]} leaf_q.  The parameter vector is the reference's concatenation
// [leaf_0 params, leaf_1 params, ...] (BinaryKernel, kernel/core.py:452-459); off[q] is leaf q's first index in it.
struct CovComposite {
    int nleaf, nterms;
    int off[GPT_MAX_LEAVES];
    int mask[GPT_MAX_TERMS];
    CovParams leaf[GPT_MAX_LEAVES];
};

// exp(x) for x <= 0 (the squared-exponential / Gibbs exponents): Cody-Waite reduction x = k ln2 + r,
// |r| <= ln2/2, degree-13 Taylor polynomial evaluated with Estrin's scheme (16 FP64 operations at dependency
// depth 5, instead of the ~20-deep Horner chain of the library exp -- on B200 scalar FP64 shares its pipe with
// DMMA, so chain depth is what limits tile generation next to a CTA doing tensor work).  Relative error
// < 4e-16 on [-708, 0]; outside that range the library exp takes over.
GPT_HD double exp_nonpos(double x) {
    if (!(x >= -708.0) || x > 0.0) return exp(x);
    const double kf = rint(x * 1.4426950408889634074);
    double r = fma(-kf, 6.93147180369123816490e-01, x);
    r = fma(-kf, 1.90821492927058770002e-10, r);
    const double r2 = r * r;
    const double a0 = fma(r, 1.0, 1.0);
    const double a1 = fma(r, 1.6666666666666665741e-01, 0.5);
    const double a2 = fma(r, 8.3333333333333332177e-03, 4.1666666666666664354e-02);
    const double a3 = fma(r, 1.9841269841269841253e-04, 1.3888888888888889419e-03);
    const double a4 = fma(r, 2.7557319223985892511e-06, 2.4801587301587301566e-05);
    const double a5 = fma(r, 2.5052108385441720224e-08, 2.7557319223985888276e-07);
    const double a6 = fma(r, 1.6059043836821613341e-10, 2.0876756987868100187e-09);
    const double r4 = r2 * r2;
    const double b0 = fma(a1, r2, a0);
    const double b1 = fma(a3, r2, a2);
    const double b2 = fma(a5, r2, a4);
    const double r8 = r4 * r4;
    const double d0 = fma(b1, r4, b0);
    const double d1 = fma(a6, r4, b2);
    const double pr = fma(d1, r8, d0);
    // 2^k with k in [-1022, 0]
#if defined(__CUDA_ARCH__)
    const double sc = __longlong_as_double((long long)(__double2int_rn(kf) + 1023) << 52);
#else
    const long long bits = (long long)((int)kf + 1023) << 52;
    double sc;
    memcpy(&sc, &bits, sizeof(sc));
#endif
    return pr * sc;
}

// Branch-free variant for the unrolled tile loops: the argument is clamped to [-708, 0] (results below
// 3e-308 are flushed to that value; irrelevant next to sigma_f^2) so that no basic-block boundary separates
// the interleaved evaluations of neighbouring entries.
GPT_HD double exp_nonpos_nobranch(double x) {
    x = fmax(x, -708.0);
    const double kf = rint(x * 1.4426950408889634074);
    double r = fma(-kf, 6.93147180369123816490e-01, x);
    r = fma(-kf, 1.90821492927058770002e-10, r);
    const double r2 = r * r;
    const double a0 = fma(r, 1.0, 1.0);
    const double a1 = fma(r, 1.6666666666666665741e-01, 0.5);
    const double a2 = fma(r, 8.3333333333333332177e-03, 4.1666666666666664354e-02);
    const double a3 = fma(r, 1.9841269841269841253e-04, 1.3888888888888889419e-03);
    const double a4 = fma(r, 2.7557319223985892511e-06, 2.4801587301587301566e-05);
    const double a5 = fma(r, 2.5052108385441720224e-08, 2.7557319223985888276e-07);
    const double a6 = fma(r, 1.6059043836821613341e-10, 2.0876756987868100187e-09);
    const double r4 = r2 * r2;
    const double b0 = fma(a1, r2, a0);
    const double b1 = fma(a3, r2, a2);
    const double b2 = fma(a5, r2, a4);
    const double r8 = r4 * r4;
    const double d0 = fma(b1, r4, b0);
    const double d1 = fma(a6, r4, b2);
    const double pr = fma(d1, r8, d0);
#if defined(__CUDA_ARCH__)
    const double sc = __longlong_as_double((long long)(__double2int_rn(kf) + 1023) << 52);
#else
    const long long bits = (long long)((int)kf + 1023) << 52;
    double sc;
    memcpy(&sc, &bits, sizeof(sc));
#endif
    return pr * sc;
}

#if defined(__CUDACC__)
// exp(x), x <= 0, with a 64-entry table of 2^(j/64) (in shared memory, pointer passed in): x = (64 k + j) ln2/64 + r,
// |r| <= ln2/128, degree-5 polynomial.  13 FP64 operations at depth ~7 instead of 22 + an F2I conversion for the
// table-free version above -- tile generation shares the FP64 pipe with the DMMA stream, so every operation saved
// there is tensor throughput.  rint and the integer k come from the 1.5 * 2^52 add trick; 2^k is applied by an
// integer add on the exponent field.  Relative error < 4e-16 on [-708, 0].
static __device__ const double GPT_EXP2_64[64] = {
    1.00000000000000000e+00, 1.01088928605170048e+00, 1.02189714865411663e+00, 1.03302487902122841e+00,
    1.04427378242741375e+00, 1.05564517836055716e+00, 1.06714040067682370e+00, 1.07876079775711986e+00,
    1.09050773266525769e+00, 1.10238258330784089e+00, 1.11438674259589243e+00, 1.12652161860824185e+00,
    1.13878863475669156e+00, 1.15118922995298267e+00, 1.16372485877757748e+00, 1.17639699165028122e+00,
    1.18920711500272103e+00, 1.20215673145270308e+00, 1.21524735998046896e+00, 1.22848053610687002e+00,
    1.24185781207348400e+00, 1.25538075702469110e+00, 1.26905095719173322e+00, 1.28287001607877826e+00,
    1.29683955465100964e+00, 1.31096121152476441e+00, 1.32523664315974132e+00, 1.33966752405330292e+00,
    1.35425554693689265e+00, 1.36900242297459052e+00, 1.38390988196383202e+00, 1.39897967253831124e+00,
    1.41421356237309515e+00, 1.42961333839197002e+00, 1.44518080697704665e+00, 1.46091779418064704e+00,
    1.47682614593949935e+00, 1.49290772829126484e+00, 1.50916442759342284e+00, 1.52559815074453842e+00,
    1.54221082540794074e+00, 1.55900440023783693e+00, 1.57598084510788650e+00, 1.59314215134226700e+00,
    1.61049033194925428e+00, 1.62802742185734783e+00, 1.64575547815396495e+00, 1.66367658032673638e+00,
    1.68179283050742900e+00, 1.70010635371852348e+00, 1.71861929812247793e+00, 1.73733383527370622e+00,
    1.75625216037329945e+00, 1.77537649252652119e+00, 1.79470907500310717e+00, 1.81425217550039886e+00,
    1.83400808640934243e+00, 1.85397912508338547e+00, 1.87416763411029996e+00, 1.89457598158696561e+00,
    1.91520656139714740e+00, 1.93606179349229435e+00, 1.95714412417540018e+00, 1.97845602638795093e+00,
};
__device__ __forceinline__ double exp_nonpos_tab(double x, const double* __restrict__ tab) {
    x = fmax(x, -708.0);
    const double MAGIC = 6755399441055744.0;
    const double tm = fma(x, 9.23324826168936567683e+01, MAGIC);
    const int ki = __double2loint(tm);
    const double kf = tm - MAGIC;
    double r = fma(-kf, 1.08304246932675596327e-02, x);
    r = fma(-kf, 2.98158582698529328128e-12, r);
    const double T = tab[ki & 63];
    const double r2 = r * r;
    const double a = 1.0 + r;
    const double b = fma(r, 1.6666666666666665741e-01, 0.5);
    const double c = fma(r, 8.3333333333333332177e-03, 4.1666666666666664354e-02);
    const double v = T * fma(fma(c, r2, b), r2, a);
    return __hiloint2double(__double2hiint(v) + ((ki >> 6) << 20), __double2loint(v));
}
#endif

// se_dim_factor restricted to m in {0, 1, 2}, written with selects only (no branches).
GPT_HD void se_dim_factor_low(double tau, double inv_l, int m, double& f, double& g) {
    // written as guarded assignments: the compiler predicates the three short bodies, so only the arithmetic
    // of the order that is actually present runs on the FP64 pipe
    const double il2 = inv_l * inv_l;
    const double u = tau * tau * il2;
    f = 1.0;
    g = u * inv_l;
    if (m == 1) {
        f = -tau * il2;
        g = tau * il2 * inv_l * (2.0 - u);
    }
    if (m == 2) {
        f = il2 * (u - 1.0);
        g = il2 * inv_l * (2.0 + u * (u - 5.0));
    }
}

// value-only variant (K-tile generation does not need the length-scale derivative factor)
GPT_HD double se_dim_value_low(double tau, double inv_l, int m) {
    const double il2 = inv_l * inv_l;
    double f = 1.0;
    if (m == 1) f = -tau * il2;
    if (m == 2) f = il2 * (tau * tau * il2 - 1.0);
    return f;
}

GPT_HD void cov_params_init(CovParams& cp, int kid, int D, int nparams, const double* params) {
    cp.kid = kid;
    cp.D = D;
    cp.nparams = nparams;
    for (int i = 0; i < GPT_MAX_PARAMS; i++) cp.p[i] = (i < nparams) ? params[i] : 0.0;
    cp.comp = nullptr;
    cp.sig2 = cp.p[0] * cp.p[0];
    const int loff = (kid == GPT_KERNEL_MATERN) ? 2 : 1;
    for (int d = 0; d < GPT_MAX_DIM; d++) cp.inv_l[d] = 0.0;
    if (kid != GPT_KERNEL_GIBBS_TANH && kid != GPT_KERNEL_GIBBS_AUX)
        for (int d = 0; d < D; d++) cp.inv_l[d] = 1.0 / cp.p[loff + d];
    cp.matern_p = 0;
    cp.mat_kind = 0;
    cp.mat_c = 0.0;
    for (int sd = 0; sd < 2; sd++) {
        cp.mat_nu[sd] = 0.0;
        for (int i = 0; i < 4; i++) { cp.mat_A[sd][i] = 0.0; cp.mat_B[sd][i] = 0.0; }
    }
    if (kid == GPT_KERNEL_MATERN) {
        const double nu = cp.p[1];
        cp.matern_p = (int)floor(nu);
        const double twice = 2.0 * nu;
        if (floor(nu) == nu) cp.mat_kind = 2;
        else if (floor(twice) == twice) cp.mat_kind = 0;
        else cp.mat_kind = 1;
        cp.mat_c = pow(2.0, 1.0 - nu) / tgamma(nu);
        // integer nu: the reference averages the series at nu -+ 0.001 (utils.py:1480-1484, 1498-1502) and keeps the
        // normalisation 2^{1-nu} / Gamma(nu) of the integer order (matern.py:377)
        const double half = (cp.mat_kind == 2) ? 0.5 : 1.0;
        cp.mat_nu[0] = (cp.mat_kind == 2) ? nu - 0.001 : nu;
        cp.mat_nu[1] = (cp.mat_kind == 2) ? nu + 0.001 : nu;
        for (int sd = 0; sd < 2; sd++) {
            const double v = cp.mat_nu[sd];
            const double g_nu = tgamma(v), g_mnu = tgamma(-v);
            for (int n = 1; n <= 3; n++) {
                double poch1 = 1.0, poch2 = 1.0;  // (1-nu)_n and (1+nu-n)_n
                for (int k = 0; k < n; k++) { poch1 *= (1.0 - v + k); poch2 *= (1.0 + v - n + k); }
                // Gamma(nu) n! / (2^{1-nu+2n} (1-nu)_n n!)   (utils.py:1503-1507 with k = n, nterms = 1)
                cp.mat_A[sd][n] = half * cp.mat_c * g_nu / (pow(2.0, 1.0 - v + 2.0 * n) * poch1);
                // Gamma(-nu) (1+nu-n)_n y^{nu-n} / 2^{1+nu}   (utils.py:1508-1515 with k = 0)
                cp.mat_B[sd][n] = half * cp.mat_c * g_mnu * poch2 / pow(2.0, 1.0 + v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Squared exponential, arbitrary derivative orders.
//   K = s2 e^{-r2/2} (-1)^{sum nj} prod_d c_d^{m_d} H_{m_d}(x_d),  c = -1/(sqrt2 l), x = tau/(sqrt2 l)
// f = c^m H_m(x); g = d f e^{..}/dl / e^{..} in the singular-free form:
//   g = c^m [ H_m(x) (tau^2/l^3 - m/l) - (sqrt2 m tau / l^2) H_{m-1}(x) ]
// ------------------------------------------------------------------------------------------
GPT_HD void se_dim_factor(double tau, double inv_l, int m, bool want_g, double& f, double& g) {
    const double RSQRT2 = 0.70710678118654752440;
    const double SQRT2 = 1.41421356237309504880;
    const double il2 = inv_l * inv_l;
    // orders 0..2 (value / first-derivative observations and predictions) in closed form:
    //   f_0 = 1, f_1 = -tau/l^2, f_2 = (u - 1)/l^2 with u = tau^2/l^2;
    //   g_0 = u/l, g_1 = tau (2 - u)/l^3, g_2 = (2 - 5u + u^2)/l^3
    if (m <= 2) {
        const double u = tau * tau * il2;
        if (m == 0) {
            f = 1.0;
            g = want_g ? u * inv_l : 0.0;
        } else if (m == 1) {
            f = -tau * il2;
            g = want_g ? tau * il2 * inv_l * (2.0 - u) : 0.0;
        } else {
            f = il2 * (u - 1.0);
            g = want_g ? il2 * inv_l * (2.0 + u * (u - 5.0)) : 0.0;
        }
        return;
    }
    const double c = -RSQRT2 * inv_l;
    const double x = tau * RSQRT2 * inv_l;
    // physicists' Hermite recurrence H_{k+1} = 2x H_k - 2k H_{k-1}
    double hm1 = 1.0;       // H_0
    double h = 2.0 * x;     // H_1
    double cm = c;          // c^1
    double twok = 2.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 1; k < m; k++) {
        const double hn = 2.0 * x * h - twok * hm1;
        hm1 = h;
        h = hn;
        cm *= c;
        twok += 2.0;
    }
    f = cm * h;
    if (want_g) {
        g = cm * (h * (tau * tau * il2 * inv_l - m * inv_l) - SQRT2 * m * tau * il2 * hm1);
    } else {
        g = 0.0;
    }
}

// hyper_deriv: -1 none, 0 sigma_f, 1+d length scale d.
GPT_HD double se_cov(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                     const int32_t* nj, int hyper_deriv) {
    double r2 = 0.0, prod = 1.0;
    int ntot_j = 0;
    for (int d = 0; d < cp.D; d++) {
        const double tau = xi[d] - xj[d];
        double tl = tau * cp.inv_l[d];
        if (tau == 0.0) tl = 0.0;  // core.py:416 (0/0 -> 0)
        r2 += tl * tl;
        const int m = ni[d] + nj[d];
        ntot_j += nj[d];
        double f, g;
        se_dim_factor(tau, cp.inv_l[d], m, hyper_deriv == d + 1, f, g);
        prod *= (hyper_deriv == d + 1) ? g : f;
    }
    double k = cp.sig2 * exp_nonpos(-0.5 * r2) * prod;
    if (ntot_j & 1) k = -k;
    if (hyper_deriv == 0) k = (cp.p[0] != 0.0) ? 2.0 * k / cp.p[0] : 0.0;
    return k;
}

// Value and all (1 + D) hyper-derivatives in one pass (one exp). out[0] = K, out[1] = dK/dsigma,
// out[2 + d] = dK/dl_d.
GPT_HD void se_cov_all(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                       const int32_t* nj, double* out) {
    double r2 = 0.0;
    int ntot_j = 0;
    double f[GPT_MAX_DIM], g[GPT_MAX_DIM];
    for (int d = 0; d < cp.D; d++) {
        const double tau = xi[d] - xj[d];
        double tl = tau * cp.inv_l[d];
        if (tau == 0.0) tl = 0.0;
        r2 += tl * tl;
        ntot_j += nj[d];
        se_dim_factor(tau, cp.inv_l[d], ni[d] + nj[d], true, f[d], g[d]);
    }
    double base = cp.sig2 * exp_nonpos(-0.5 * r2);
    if (ntot_j & 1) base = -base;
    double prod = 1.0;
    for (int d = 0; d < cp.D; d++) prod *= f[d];
    out[0] = base * prod;
    out[1] = (cp.p[0] != 0.0) ? 2.0 * out[0] / cp.p[0] : 0.0;
    for (int d = 0; d < cp.D; d++) {
        double pr = g[d];
        for (int e = 0; e < cp.D; e++)
            if (e != d) pr *= f[e];
        out[2 + d] = base * pr;
    }
}

// ------------------------------------------------------------------------------------------
// Matern 5/2, value / first derivative / mixed second derivative (matern.c:61-186).
// ------------------------------------------------------------------------------------------
GPT_HD int first_one(const int32_t* n, int D) {
    for (int d = 0; d < D; d++)
        if (n[d] == 1) return d;
    return -1;
}

GPT_HD double matern52_cov(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                           const int32_t* nj) {
    const double SQRT_5 = 2.2360679774997898;
    const double FIVE_THIRDS = 1.6666666666666667;
    double r2 = 0.0;
    for (int d = 0; d < cp.D; d++) {
        const double disp = xi[d] - xj[d];
        const double var = cp.p[1 + d] * cp.p[1 + d];
        r2 = r2 + disp * disp / var;
    }
    const int a = first_one(ni, cp.D);
    const int b = first_one(nj, cp.D);
    double v;
    if (a < 0 && b < 0) {
        if (r2 == 0.0) {
            v = 1.0;
        } else {
            const double s = SQRT_5 * sqrt(r2);
            v = (1.0 + s + FIVE_THIRDS * r2) * exp(-s);
        }
    } else if (a < 0 || b < 0) {
        if (r2 == 0.0) {
            v = 0.0;
        } else {
            const int n = (a >= 0) ? a : b;
            const double var = cp.p[1 + n] * cp.p[1 + n];
            double disp = xi[n] - xj[n];
            if (a < 0) disp = -disp;  // derivative w.r.t. Xj: arguments swapped (matern.c:182-184)
            const double s = SQRT_5 * sqrt(r2);
            v = -FIVE_THIRDS * (1.0 + s) * exp(-s) * (disp / var);
        }
    } else {
        const double varn = cp.p[1 + a] * cp.p[1 + a];
        const double varm = cp.p[1 + b] * cp.p[1 + b];
        if (r2 == 0.0) {
            v = (a == b) ? FIVE_THIRDS / varn : 0.0;
        } else {
            const double r = sqrt(r2);
            const double dn = xi[a] - xj[a];
            const double dm = xi[b] - xj[b];
            const double dr_dXn = dn / (r * varn);
            const double dr_dYm = -dm / (r * varm);
            double d2r_r3 = dn * dm / (varn * varm);
            if (a == b) d2r_r3 -= r * r / varn;
            const double s = SQRT_5 * r;
            const double e = exp(-s);
            const double dk_over_r = -FIVE_THIRDS * (1.0 + s) * e;
            const double d2k = FIVE_THIRDS * (5.0 * r2 - s - 1.0) * e;
            v = dk_over_r * d2r_r3 / r2 + d2k * dr_dXn * dr_dYm;
        }
    }
    return cp.sig2 * v;
}

// ------------------------------------------------------------------------------------------
// Generic Matern, any nu > 0, total derivative order <= 2.
//   f(y) = c y^{nu/2} K_nu(sqrt y),  c = 2^{1-nu}/Gamma(nu),  y = 2 nu r2l2
//   g_mu(y) := y^{mu/2} K_|mu|(sqrt y);  d/dy g_mu = -1/2 g_{mu-1}   =>   f^{(n)} = c (-1/2)^n g_{nu-n}
//   K_{q+1/2}(r) = sqrt(pi/(2r)) e^{-r} sum_{k<=q} (q+k)!/(k!(q-k)!) (2r)^{-k}
// ------------------------------------------------------------------------------------------
GPT_HD double bessel_k_half(int q, double r) {
    // K_{q+1/2}(r), q >= 0 (K_{-1/2} = K_{1/2} handled by the caller)
    double sum = 1.0, term = 1.0;
    for (int k = 1; k <= q; k++) {
        term *= (double)(q + k) * (double)(q - k + 1) / ((double)k * 2.0 * r);
        sum += term;
    }
    return sqrt(1.5707963267948966 / r) * exp(-r) * sum;
}

// ---- K_nu(x) for real order (generic Matern with nu not a half-integer, kernel/matern.py:296-312 calls scipy's kv) ----
// Temme's method (N. M. Temme, J. Comput. Phys. 19 (1975) 324): for |mu| <= 1/2, K_mu and K_{mu+1} from the series
// in x/2 when x <= 2 and from Steed's continued fraction CF2 when x > 2; upward recurrence in the order (stable for
// K) reaches |nu|.  The two Gamma-function combinations of the series,
//   gam1 = (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu),   gam2 = (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2,
// come from the Taylor series 1/Gamma(1+z) = sum c_k z^k (odd / even parts: no cancellation at small mu).
GPT_HD void temme_gammas(double mu, double& gam1, double& gam2, double& gampl, double& gammi) {
    const double C[27] = {
        1.0, 0.57721566490153286061, -0.65587807152025388108, -0.042002635034095235529, 0.1665386113822914895,
        -0.042197734555544336748, -0.0096219715278769735621, 0.0072189432466630995424, -0.0011651675918590651121,
        -0.00021524167411495097282, 0.00012805028238811618615, -0.000020134854780788238656,
        -1.2504934821426706573e-6, 1.1330272319816958824e-6, -2.0563384169776071035e-7, 6.1160951044814158179e-9,
        5.0020076444692229301e-9, -1.1812745704870201446e-9, 1.0434267116911005105e-10, 7.782263439905071254e-12,
        -3.6968056186422057082e-12, 5.100370287454475979e-13, -2.0583260535665067832e-14, -5.3481225394230179824e-15,
        1.2267786282382607902e-15, -1.1812593016974587695e-16, 1.1866922547516003326e-18};
    const double m2 = mu * mu;
    double odd = 0.0, even = 0.0;
    for (int k = 25; k >= 1; k -= 2) odd = odd * m2 + C[k];
    for (int k = 26; k >= 0; k -= 2) even = even * m2 + C[k];
    gam1 = -odd;
    gam2 = even;
    gampl = gam2 - mu * gam1;  // 1 / Gamma(1 + mu)
    gammi = gam2 + mu * gam1;  // 1 / Gamma(1 - mu)
}

GPT_HD void bessel_k_pair(double mu, double x, double& kmu, double& kmu1) {
    const double PI = 3.14159265358979323846;
    if (x <= 2.0) {
        const double b = 0.5 * x, d0 = -log(b);
        double e = mu * d0;
        const double fact2 = (fabs(e) < 1e-8) ? 1.0 : sinh(e) / e;
        const double pimu = PI * mu;
        const double fact = (fabs(pimu) < 1e-8) ? 1.0 : pimu / sin(pimu);
        double gam1, gam2, gampl, gammi;
        temme_gammas(mu, gam1, gam2, gampl, gammi);
        double ff = fact * (gam1 * cosh(e) + gam2 * fact2 * d0);
        double sum = ff;
        e = exp(e);
        double p = 0.5 * e / gampl, q = 0.5 / (e * gammi), c = 1.0, sum1 = p;
        const double d = b * b, mu2 = mu * mu;
        for (int i = 1; i <= 500; i++) {
            ff = (i * ff + p + q) / (i * (double)i - mu2);
            c *= d / i;
            p /= (i - mu);
            q /= (i + mu);
            const double del = c * ff;
            sum += del;
            sum1 += c * (p - i * ff);
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        kmu = sum;
        kmu1 = sum1 * (2.0 / x);
    } else {
        double b = 2.0 * (1.0 + x), d = 1.0 / b, h = d, delh = d;
        double q1 = 0.0, q2 = 1.0;
        const double a1 = 0.25 - mu * mu;
        double q = a1, c = a1, a = -a1;
        double s = 1.0 + q * delh;
        for (int i = 2; i <= 10000; i++) {
            a -= 2 * (i - 1);
            c = -a * c / i;
            const double qnew = (q1 - b * q2) / a;
            q1 = q2;
            q2 = qnew;
            q += c * qnew;
            b += 2.0;
            d = 1.0 / (b + a * d);
            delh = (b * d - 1.0) * delh;
            h += delh;
            const double dels = q * delh;
            s += dels;
            if (fabs(dels / s) < 1e-17) break;
        }
        h = a1 * h;
        kmu = sqrt(PI / (2.0 * x)) * exp(-x) / s;
        kmu1 = kmu * (mu + x + 0.5 - h) / x;
    }
}

GPT_HD double bessel_k_real(double order, double x) {
    const double a = fabs(order);  // K_{-a} = K_a
    const int nl = (int)(a + 0.5);
    const double mu = a - nl;      // |mu| <= 1/2
    double k0, k1;
    bessel_k_pair(mu, x, k0, k1);
    const double xi2 = 2.0 / x;
    for (int i = 1; i <= nl; i++) {
        const double kt = (mu + i) * xi2 * k1 + k0;
        k0 = k1;
        k1 = kt;
    }
    return k0;
}

GPT_HD double matern_fn(const CovParams& cp, double y, int n) {
    // exact f^{(n)}(y) = c (-1/2)^n y^{(nu-n)/2} K_{|nu-n|}(sqrt y) for y > 0
    const double nu = cp.p[1];
    const double r = sqrt(y);
    const double mu = nu - n;  // may be negative
    const double amu = fabs(mu);
    double s = (n & 1) ? -1.0 : 1.0;
    for (int k = 0; k < n; k++) s *= 0.5;
    const double kv = (cp.mat_kind == 0) ? bessel_k_half((int)floor(amu), r) : bessel_k_real(amu, r);
    return cp.mat_c * s * pow(r, mu) * kv;
}

// one side of utils.py:1477-1492 at y == 0 (without the normalisation c): +-inf when n > nu', the series constant else
GPT_HD double matern_origin_side(double v, int n) {
    if ((double)n > v) {
        double poch = 1.0;
        for (int k = 0; k < n; k++) poch *= (1.0 + v - n + k);
        return (tgamma(-v) * poch) * INFINITY;
    }
    double poch1 = 1.0;
    for (int k = 0; k < n; k++) poch1 *= (1.0 - v + k);
    return tgamma(v) / (pow(2.0, 1.0 - v + 2.0 * n) * poch1);
}

GPT_HD double matern_dk_dy(const CovParams& cp, double y, int n) {
    // utils.py:1429-1518 for n >= 1 (value n == 0 is handled by the caller)
    if (y == 0.0) {
        if (cp.mat_kind == 2)
            return cp.mat_c * 0.5 * (matern_origin_side(cp.mat_nu[0], n) + matern_origin_side(cp.mat_nu[1], n));
        return cp.mat_c * matern_origin_side(cp.p[1], n);
    }
    if (y <= 5e-4) {
        double v = cp.mat_A[0][n] + cp.mat_B[0][n] * pow(y, cp.mat_nu[0] - n);
        if (cp.mat_kind == 2) v += cp.mat_A[1][n] + cp.mat_B[1][n] * pow(y, cp.mat_nu[1] - n);
        return v;
    }
    return matern_fn(cp, y, n);
}

GPT_HD double matern_cov(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                         const int32_t* nj) {
    const double nu = cp.p[1];
    double r2 = 0.0;
    int ntot_j = 0, order = 0;
    int dims[2] = {-1, -1};
    double tau_d[2] = {0.0, 0.0};
    for (int d = 0; d < cp.D; d++) {
        const double tau = xi[d] - xj[d];
        double tl = tau * cp.inv_l[d];
        if (tau == 0.0) tl = 0.0;
        r2 += tl * tl;
        ntot_j += nj[d];
        const int m = ni[d] + nj[d];
        for (int k = 0; k < m; k++) {
            if (order < 2) { dims[order] = d; tau_d[order] = tau; }
            order++;
        }
    }
    const double y = 2.0 * nu * r2;
    double v;
    if (order == 0) {
        v = (r2 == 0.0) ? 1.0 : matern_fn(cp, y, 0);
    } else if (order == 1) {
        // single partition {a}: f'(y) * dy/dtau_a, dy/dtau = 4 nu tau / l^2 (matern.py:404-405)
        const double il = cp.inv_l[dims[0]];
        double dk = matern_dk_dy(cp, y, 1);
        if (y == 0.0) {
            const double tau_pow = 2.0 * (nu - 1.0) + 1.0;  // matern.py:448-455
            if (tau_pow == 0.0) dk = NAN;
            else if (tau_pow > 0.0) dk = 0.0;
        }
        v = dk * (4.0 * nu * tau_d[0] * il * il);
    } else if (order == 2) {
        // partitions {a,b} (one block) and {a},{b} (two blocks)
        const double ila = cp.inv_l[dims[0]], ilb = cp.inv_l[dims[1]];
        double t1 = 0.0;
        if (dims[0] == dims[1]) t1 = matern_dk_dy(cp, y, 1) * (4.0 * nu * ila * ila);
        double dk2 = matern_dk_dy(cp, y, 2);
        if (y == 0.0) {
            const double tau_pow = 2.0 * (nu - 2.0) + 2.0;
            if (tau_pow == 0.0) dk2 = NAN;
            else if (tau_pow > 0.0) dk2 = 0.0;
        }
        const double t2 = dk2 * (4.0 * nu * tau_d[0] * ila * ila) * (4.0 * nu * tau_d[1] * ilb * ilb);
        v = t1 + t2;
    } else {
        v = NAN;  // rejected on the host before any device call
    }
    if (ntot_j & 1) v = -v;
    return cp.sig2 * v;
}

// ------------------------------------------------------------------------------------------
// Gibbs kernel with tanh length-scale warp, 1-D, (ni, nj) in {0,1}^2.
//   k00 = sqrt(2 lx ly / S) exp(-d^2/S), S = lx^2 + ly^2, d = x - y
//   k10 = k00 A_x, k01 = k00 A_y, k11 = k00 (A_x A_y + d/dy A_x)       
// ------------------------------------------------------------------------------------------
GPT_HD void gibbs_tanh_l(const CovParams& cp, double x, double& l, double& l1) {
    const double la = cp.p[1], lb = cp.p[2], lw = cp.p[3], x0 = cp.p[4];
    const double t = tanh((x - x0) / lw);
    l = 0.5 * (la + lb) - 0.5 * (la - lb) * t;
    l1 = -(la - lb) / (2.0 * lw) * (1.0 - t * t);  // cosh^-2 = 1 - tanh^2
}

GPT_HD double gibbs_cov_l(const CovParams& cp, double x, double lx, double lx1, int a, double y, double ly,
                          double ly1, int b) {
    const double d = x - y;
    const double S = lx * lx + ly * ly;
    const double iS = 1.0 / S;
    const double k00 = sqrt(2.0 * lx * ly * iS) * exp(-d * d * iS);
    double v = k00;
    if (a | b) {
        const double Ax = lx1 / (2.0 * lx) - lx * lx1 * iS - 2.0 * d * iS + 2.0 * d * d * lx * lx1 * iS * iS;
        const double Ay = ly1 / (2.0 * ly) - ly * ly1 * iS + 2.0 * d * iS + 2.0 * d * d * ly * ly1 * iS * iS;
        if (a && b) {
            const double dAx = 2.0 * lx * lx1 * ly * ly1 * iS * iS + 2.0 * iS + 4.0 * d * ly * ly1 * iS * iS -
                               4.0 * d * lx * lx1 * iS * iS - 8.0 * d * d * lx * lx1 * ly * ly1 * iS * iS * iS;
            v = k00 * (Ax * Ay + dAx);
        } else if (a) {
            v = k00 * Ax;
        } else {
            v = k00 * Ay;
        }
    }
    return cp.sig2 * v;
}

GPT_HD double gibbs_cov(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                        const int32_t* nj) {
    double lx, lx1, ly, ly1;
    gibbs_tanh_l(cp, xi[0], lx, lx1);
    gibbs_tanh_l(cp, xj[0], ly, ly1);
    return gibbs_cov_l(cp, xi[0], lx, lx1, ni[0], xj[0], ly, ly1, nj[0]);
}

#include "covfn_hyper.cuh"

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
GPT_HD double cov_eval_leaf(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                            const int32_t* nj, int hyper_deriv) {
    if (hyper_deriv >= 0 && cp.kid == GPT_KERNEL_GIBBS_AUX) {
        // only sigma_f is a device-side parameter: dk/dsigma_f = 2 k / sigma_f
        const double k = gibbs_cov_l(cp, xi[0], xi[1], xi[2], ni[0], xj[0], xj[1], xj[2], nj[0]);
        return (hyper_deriv == 0 && cp.p[0] != 0.0) ? 2.0 * k / cp.p[0] : NAN;
    }
    if (hyper_deriv >= 0 && cp.kid != GPT_KERNEL_SE) return cov_hyper_eval(cp, xi, ni, xj, nj, hyper_deriv);
    switch (cp.kid) {
        case GPT_KERNEL_SE: return se_cov(cp, xi, ni, xj, nj, hyper_deriv);
        case GPT_KERNEL_MATERN52: return matern52_cov(cp, xi, ni, xj, nj);
        case GPT_KERNEL_MATERN: return matern_cov(cp, xi, ni, xj, nj);
        case GPT_KERNEL_GIBBS_AUX: return gibbs_cov_l(cp, xi[0], xi[1], xi[2], ni[0], xj[0], xj[1], xj[2], nj[0]);
        default: return gibbs_cov(cp, xi, ni, xj, nj);
    }
}

// ------------------------------------------------------------------------------------------
// kernel algebra: sums and products of the kernels above, with derivative orders
// ------------------------------------------------------------------------------------------
// A product of L kernels under the derivative multi-index m = (ni, nj) (2 D slots) follows the general Leibniz rule:
// the reference enumerates every subset of the multiset of unit derivatives (kernel/core.py:632-668); collecting
// equal terms gives  sum_{a <= m} prod_s C(m_s, a_s) k1^(a) k2^(m - a),  applied recursively for three factors.
// hyper_deriv (index into the concatenated parameter vector) differentiates the one leaf that owns the parameter:
// terms without that leaf drop out (SumKernel: kernel/core.py:576-582; for products the reference raises
// NotImplementedError, here the same Leibniz sum runs with the owning leaf's own hyper-derivative).
GPT_HD double binom_small(int n, int k) {
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (double)(n - k + i) / (double)i;
    return r;
}

// One out-of-line copy of the leaf closed forms serves every factor of every term (inlining them at each call site
// multiplies the code of all kernels that evaluate covariances by the number of sites).
#if defined(__CUDACC__)
#define GPT_NOINLINE static __host__ __device__ __noinline__
#else
#define GPT_NOINLINE static inline
#endif
GPT_NOINLINE double cov_leaf_call(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                                  const int32_t* nj, int hyper_deriv) {
    return cov_eval_leaf(cp, xi, ni, xj, nj, hyper_deriv);
}

// sum over splits of (mi, mj) between leaf A and leaf B
GPT_NOINLINE double comp_prod2(const CovParams& A, int hdA, const CovParams& B, int hdB, int D, const double* xi,
                         const int32_t* mi, const double* xj, const int32_t* mj) {
    int32_t ai[GPT_MAX_DIM], aj[GPT_MAX_DIM], bi[GPT_MAX_DIM], bj[GPT_MAX_DIM];
    int total = 0;
    for (int d = 0; d < D; d++) {
        ai[d] = aj[d] = 0;
        total += mi[d] + mj[d];
    }
    if (total == 0) return cov_leaf_call(A, xi, mi, xj, mj, hdA) * cov_leaf_call(B, xi, mi, xj, mj, hdB);
    double sum = 0.0;
    for (;;) {
        double w = 1.0;
        for (int d = 0; d < D; d++) {
            bi[d] = mi[d] - ai[d];
            bj[d] = mj[d] - aj[d];
            w *= binom_small(mi[d], ai[d]) * binom_small(mj[d], aj[d]);
        }
        sum += w * cov_leaf_call(A, xi, ai, xj, aj, hdA) * cov_leaf_call(B, xi, bi, xj, bj, hdB);
        // odometer over the 2 D slots
        int s = 0;
        for (; s < 2 * D; s++) {
            int32_t* a = (s < D) ? &ai[s] : &aj[s - D];
            const int lim = (s < D) ? mi[s] : mj[s - D];
            if (*a < lim) {
                (*a)++;
                break;
            }
            *a = 0;
        }
        if (s == 2 * D) break;
    }
    return sum;
}

GPT_NOINLINE double comp_eval(const CovComposite& c, int D, const double* xi, const int32_t* ni, const double* xj, const int32_t* nj,
                 int hyper_deriv) {
    int owner = -1, hl = -1;
    if (hyper_deriv >= 0) {
        for (int q = 0; q < c.nleaf; q++)
            if (hyper_deriv >= c.off[q] && hyper_deriv < c.off[q] + c.leaf[q].nparams) {
                owner = q;
                hl = hyper_deriv - c.off[q];
            }
        if (owner < 0) return NAN;
    }
    double total = 0.0;
    for (int t = 0; t < c.nterms; t++) {
        const int mask = c.mask[t];
        if (owner >= 0 && !((mask >> owner) & 1)) continue;
        int l[GPT_MAX_LEAVES], L = 0;
        for (int q = 0; q < c.nleaf; q++)
            if ((mask >> q) & 1) l[L++] = q;
        if (L == 0) continue;
        const int h0 = (l[0] == owner) ? hl : -1;
        if (L == 1) {
            total += cov_leaf_call(c.leaf[l[0]], xi, ni, xj, nj, h0);
            continue;
        }
        const int h1 = (l[1] == owner) ? hl : -1;
        if (L == 2) {
            total += comp_prod2(c.leaf[l[0]], h0, c.leaf[l[1]], h1, D, xi, ni, xj, nj);
            continue;
        }
        // three or four factors: split off the leading leaf (and the second for four), the last two go through comp_prod2
        const int h2 = (l[2] == owner) ? hl : -1;
        const int h3 = (L > 3 && l[3] == owner) ? hl : -1;
        int32_t ai[GPT_MAX_DIM], aj[GPT_MAX_DIM], ri[GPT_MAX_DIM], rj[GPT_MAX_DIM];
        for (int d = 0; d < D; d++) ai[d] = aj[d] = 0;
        for (;;) {
            double w = 1.0;
            for (int d = 0; d < D; d++) {
                ri[d] = ni[d] - ai[d];
                rj[d] = nj[d] - aj[d];
                w *= binom_small(ni[d], ai[d]) * binom_small(nj[d], aj[d]);
            }
            const double lead = w * cov_leaf_call(c.leaf[l[0]], xi, ai, xj, aj, h0);
            if (L == 3) {
                total += lead * comp_prod2(c.leaf[l[1]], h1, c.leaf[l[2]], h2, D, xi, ri, xj, rj);
            } else {
                int32_t bi[GPT_MAX_DIM], bj[GPT_MAX_DIM], si[GPT_MAX_DIM], sj[GPT_MAX_DIM];
                for (int d = 0; d < D; d++) bi[d] = bj[d] = 0;
                for (;;) {
                    double w2 = 1.0;
                    for (int d = 0; d < D; d++) {
                        si[d] = ri[d] - bi[d];
                        sj[d] = rj[d] - bj[d];
                        w2 *= binom_small(ri[d], bi[d]) * binom_small(rj[d], bj[d]);
                    }
                    total += lead * w2 * cov_leaf_call(c.leaf[l[1]], xi, bi, xj, bj, h1) *
                             comp_prod2(c.leaf[l[2]], h2, c.leaf[l[3]], h3, D, xi, si, xj, sj);
                    int s = 0;
                    for (; s < 2 * D; s++) {
                        int32_t* b = (s < D) ? &bi[s] : &bj[s - D];
                        const int lim = (s < D) ? ri[s] : rj[s - D];
                        if (*b < lim) {
                            (*b)++;
                            break;
                        }
                        *b = 0;
                    }
                    if (s == 2 * D) break;
                }
            }
            int s = 0;
            for (; s < 2 * D; s++) {
                int32_t* a = (s < D) ? &ai[s] : &aj[s - D];
                const int lim = (s < D) ? ni[s] : nj[s - D];
                if (*a < lim) {
                    (*a)++;
                    break;
                }
                *a = 0;
            }
            if (s == 2 * D) break;
        }
    }
    return total;
}

GPT_HD double cov_eval(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                       const int32_t* nj, int hyper_deriv) {
    if (cp.kid == GPT_KERNEL_COMPOSITE) return comp_eval(*cp.comp, cp.D, xi, ni, xj, nj, hyper_deriv);
    return cov_eval_leaf(cp, xi, ni, xj, nj, hyper_deriv);
}

// Host side of a composite: leaves initialised from the concatenated parameter vector.  `kids` / `nps` describe the
// leaves, `masks` the product terms.  Returns the number of parameters, or -1 when the description is not valid.
GPT_HD int comp_init(CovComposite& c, int D, int nleaf, const int32_t* kids, const int32_t* nps, int nterms,
                     const int32_t* masks, const double* params) {
    if (nleaf < 1 || nleaf > GPT_MAX_LEAVES || nterms < 1 || nterms > GPT_MAX_TERMS) return -1;
    for (int q = 0; q < GPT_MAX_LEAVES; q++) c.off[q] = 0;
    for (int t = 0; t < GPT_MAX_TERMS; t++) c.mask[t] = 0;
    c.nleaf = nleaf;
    c.nterms = nterms;
    int off = 0;
    for (int q = 0; q < nleaf; q++) {
        if (nps[q] < 1 || off + nps[q] > GPT_MAX_PARAMS) return -1;
        c.off[q] = off;
        cov_params_init(c.leaf[q], kids[q], D, nps[q], params + off);
        off += nps[q];
    }
    for (int t = 0; t < nterms; t++) {
        if (masks[t] <= 0 || masks[t] >= (1 << nleaf)) return -1;
        c.mask[t] = masks[t];
    }
    return off;
}
