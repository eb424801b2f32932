// Hyper-parameter derivatives dk/dtheta_p of the Matern-5/2, generic (half-integer) Matern and Gibbs-tanh
// covariances, including all the derivative-observation orders the value paths support.
//
// The reference has no such code: Matern52Kernel / MaternKernel / GibbsKernel1d raise NotImplementedError for
// hyper_deriv (kernel/matern.py:543, kernel/core.py:723, kernel/gibbs.py:319), so with use_hyper_deriv=True
// every evaluation is swallowed into (inf, 0) by update_hyperparameters (gaussian_process.py:1391-1402) and
// config 2 can only be optimised value-only.  SURVEY.md section 8(f) row 2 asks for these; the oracle is the
// central finite difference of the reference's own ll / K (tests/golden/hyperfd_*.npz).
//
// Implementation: the closed forms of covfn.cuh re-evaluated in forward-mode dual numbers with the seed on one
// hyper-parameter.  The value component follows the same operation order as the value path, so a dual
// evaluation reproduces the value path bit for bit in .v (checked on the host in tests/test_covfn_host.py).
#pragma once

struct GptDual {
    double v, d;
};
GPT_HD GptDual dmk(double v, double d = 0.0) {
    GptDual r;
    r.v = v;
    r.d = d;
    return r;
}
GPT_HD GptDual operator+(GptDual a, GptDual b) { return dmk(a.v + b.v, a.d + b.d); }
GPT_HD GptDual operator+(GptDual a, double b) { return dmk(a.v + b, a.d); }
GPT_HD GptDual operator+(double a, GptDual b) { return dmk(a + b.v, b.d); }
GPT_HD GptDual operator-(GptDual a, GptDual b) { return dmk(a.v - b.v, a.d - b.d); }
GPT_HD GptDual operator-(GptDual a, double b) { return dmk(a.v - b, a.d); }
GPT_HD GptDual operator-(double a, GptDual b) { return dmk(a - b.v, -b.d); }
GPT_HD GptDual operator-(GptDual a) { return dmk(-a.v, -a.d); }
GPT_HD GptDual operator*(GptDual a, GptDual b) { return dmk(a.v * b.v, a.d * b.v + a.v * b.d); }
GPT_HD GptDual operator*(GptDual a, double b) { return dmk(a.v * b, a.d * b); }
GPT_HD GptDual operator*(double a, GptDual b) { return dmk(a * b.v, a * b.d); }
GPT_HD GptDual operator/(GptDual a, GptDual b) {
    const double q = a.v / b.v;
    return dmk(q, (a.d - q * b.d) / b.v);
}
GPT_HD GptDual operator/(GptDual a, double b) { return dmk(a.v / b, a.d / b); }
GPT_HD GptDual operator/(double a, GptDual b) {
    const double q = a / b.v;
    return dmk(q, -q * b.d / b.v);
}
GPT_HD GptDual dexp(GptDual a) {
    const double e = exp(a.v);
    return dmk(e, e * a.d);
}
GPT_HD GptDual dsqrt(GptDual a) {
    const double s = sqrt(a.v);
    return dmk(s, a.d / (2.0 * s));
}
GPT_HD GptDual dtanh(GptDual a) {
    const double t = tanh(a.v);
    return dmk(t, (1.0 - t * t) * a.d);
}
GPT_HD GptDual dpow(GptDual a, double q) {
    const double p = pow(a.v, q);
    return dmk(p, (a.v != 0.0) ? q * p / a.v * a.d : 0.0);
}

// parameter p of the kernel as a dual seeded on hyper_deriv
GPT_HD GptDual hparam(const CovParams& cp, int idx, int hyper_deriv) { return dmk(cp.p[idx], idx == hyper_deriv ? 1.0 : 0.0); }

// ---- Matern 5/2 (value path: matern52_cov, reference kernel/src/matern.c:61-186) -----------------------------------
GPT_HD GptDual matern52_cov_dual(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                                 const int32_t* nj, int hd) {
    const double SQRT_5 = 2.2360679774997898;
    const double FIVE_THIRDS = 1.6666666666666667;
    GptDual r2 = dmk(0.0);
    for (int d = 0; d < cp.D; d++) {
        const double disp = xi[d] - xj[d];
        const GptDual l = hparam(cp, 1 + d, hd);
        const GptDual var = l * l;
        r2 = r2 + (disp * disp) / var;
    }
    const int a = first_one(ni, cp.D);
    const int b = first_one(nj, cp.D);
    GptDual v;
    if (a < 0 && b < 0) {
        if (r2.v == 0.0) {
            v = dmk(1.0);
        } else {
            const GptDual s = SQRT_5 * dsqrt(r2);
            v = (1.0 + s + FIVE_THIRDS * r2) * dexp(-s);
        }
    } else if (a < 0 || b < 0) {
        if (r2.v == 0.0) {
            v = dmk(0.0);
        } else {
            const int n = (a >= 0) ? a : b;
            const GptDual l = hparam(cp, 1 + n, hd);
            const GptDual var = l * l;
            double disp = xi[n] - xj[n];
            if (a < 0) disp = -disp;
            const GptDual s = SQRT_5 * dsqrt(r2);
            v = -FIVE_THIRDS * (1.0 + s) * dexp(-s) * (disp / var);
        }
    } else {
        const GptDual la = hparam(cp, 1 + a, hd), lb = hparam(cp, 1 + b, hd);
        const GptDual varn = la * la, varm = lb * lb;
        if (r2.v == 0.0) {
            v = (a == b) ? FIVE_THIRDS / varn : dmk(0.0);
        } else {
            const GptDual r = dsqrt(r2);
            const double dn = xi[a] - xj[a];
            const double dm = xi[b] - xj[b];
            const GptDual dr_dXn = dn / (r * varn);
            const GptDual dr_dYm = -dm / (r * varm);
            GptDual d2r_r3 = (dn * dm) / (varn * varm);
            if (a == b) d2r_r3 = d2r_r3 - r * r / varn;
            const GptDual s = SQRT_5 * r;
            const GptDual e = dexp(-s);
            const GptDual dk_over_r = -FIVE_THIRDS * (1.0 + s) * e;
            const GptDual d2k = FIVE_THIRDS * (5.0 * r2 - s - 1.0) * e;
            v = dk_over_r * d2r_r3 / r2 + d2k * dr_dXn * dr_dYm;
        }
    }
    const GptDual sf = hparam(cp, 0, hd);
    return sf * sf * v;
}

// ---- generic Matern, nu fixed (any nu > 0), total derivative order <= 2 (value path: matern_cov) --------------------
GPT_HD GptDual bessel_k_half_dual(int q, GptDual r) {
    GptDual sum = dmk(1.0), term = dmk(1.0);
    for (int k = 1; k <= q; k++) {
        term = term * ((double)(q + k) * (double)(q - k + 1)) / ((double)k * 2.0 * r);
        sum = sum + term;
    }
    return dsqrt(1.5707963267948966 / r) * dexp(-r) * sum;
}

GPT_HD GptDual matern_fn_dual(const CovParams& cp, GptDual y, int n) {
    const double nu = cp.p[1];
    if (cp.mat_kind != 0) {
        // any other order: d f^{(n)} / dy = f^{(n+1)}, both from the real-order Bessel routine
        return dmk(matern_fn(cp, y.v, n), matern_fn(cp, y.v, n + 1) * y.d);
    }
    const GptDual r = dsqrt(y);
    const double mu = nu - n;
    const int q = (int)floor(fabs(mu));
    double s = (n & 1) ? -1.0 : 1.0;
    for (int k = 0; k < n; k++) s *= 0.5;
    return (cp.mat_c * s) * dpow(r, mu) * bessel_k_half_dual(q, r);
}

GPT_HD GptDual matern_dk_dy_dual(const CovParams& cp, GptDual y, int n) {
    if (y.v == 0.0) return dmk(matern_dk_dy(cp, 0.0, n));  // y == 0 has dy/dl == 0 as well
    if (y.v <= 5e-4) {
        GptDual v = cp.mat_A[0][n] + cp.mat_B[0][n] * dpow(y, cp.mat_nu[0] - n);
        if (cp.mat_kind == 2) v = v + (cp.mat_A[1][n] + cp.mat_B[1][n] * dpow(y, cp.mat_nu[1] - n));
        return v;
    }
    return matern_fn_dual(cp, y, n);
}

GPT_HD GptDual matern_cov_dual(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                               const int32_t* nj, int hd) {
    const double nu = cp.p[1];
    GptDual r2 = dmk(0.0);
    int ntot_j = 0, order = 0;
    int dims[2] = {-1, -1};
    double tau_d[2] = {0.0, 0.0};
    for (int d = 0; d < cp.D; d++) {
        const double tau = xi[d] - xj[d];
        const GptDual il = 1.0 / hparam(cp, 2 + d, hd);
        GptDual tl = tau * il;
        if (tau == 0.0) tl = dmk(0.0);
        r2 = r2 + tl * tl;
        ntot_j += nj[d];
        const int m = ni[d] + nj[d];
        for (int k = 0; k < m; k++) {
            if (order < 2) { dims[order] = d; tau_d[order] = tau; }
            order++;
        }
    }
    const GptDual y = (2.0 * nu) * r2;
    GptDual v;
    if (order == 0) {
        v = (r2.v == 0.0) ? dmk(1.0) : matern_fn_dual(cp, y, 0);
    } else if (order == 1) {
        const GptDual il = 1.0 / hparam(cp, 2 + dims[0], hd);
        GptDual dk = matern_dk_dy_dual(cp, y, 1);
        if (y.v == 0.0) {
            const double tau_pow = 2.0 * (nu - 1.0) + 1.0;
            if (tau_pow == 0.0) dk = dmk(NAN);
            else if (tau_pow > 0.0) dk = dmk(0.0);
        }
        v = dk * ((4.0 * nu * tau_d[0]) * il * il);
    } else if (order == 2) {
        const GptDual ila = 1.0 / hparam(cp, 2 + dims[0], hd), ilb = 1.0 / hparam(cp, 2 + dims[1], hd);
        GptDual t1 = dmk(0.0);
        if (dims[0] == dims[1]) t1 = matern_dk_dy_dual(cp, y, 1) * ((4.0 * nu) * ila * ila);
        GptDual dk2 = matern_dk_dy_dual(cp, y, 2);
        if (y.v == 0.0) {
            const double tau_pow = 2.0 * (nu - 2.0) + 2.0;
            if (tau_pow == 0.0) dk2 = dmk(NAN);
            else if (tau_pow > 0.0) dk2 = dmk(0.0);
        }
        const GptDual t2 = dk2 * ((4.0 * nu * tau_d[0]) * ila * ila) * ((4.0 * nu * tau_d[1]) * ilb * ilb);
        v = t1 + t2;
    } else {
        v = dmk(NAN, NAN);
    }
    if (ntot_j & 1) v = -v;
    const GptDual sf = hparam(cp, 0, hd);
    return sf * sf * v;
}

// ---- Gibbs kernel with tanh length-scale warp (value path: gibbs_cov, reference kernel/gibbs.py:288-465) ----------
GPT_HD void gibbs_tanh_l_dual(const CovParams& cp, double x, int hd, GptDual& l, GptDual& l1) {
    const GptDual la = hparam(cp, 1, hd), lb = hparam(cp, 2, hd), lw = hparam(cp, 3, hd), x0 = hparam(cp, 4, hd);
    const GptDual t = dtanh((x - x0) / lw);
    l = 0.5 * (la + lb) - 0.5 * (la - lb) * t;
    l1 = -(la - lb) / (2.0 * lw) * (1.0 - t * t);
}

GPT_HD GptDual gibbs_cov_dual(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj,
                              const int32_t* nj, int hd) {
    GptDual lx, lx1, ly, ly1;
    gibbs_tanh_l_dual(cp, xi[0], hd, lx, lx1);
    gibbs_tanh_l_dual(cp, xj[0], hd, ly, ly1);
    const int a = ni[0], b = nj[0];
    const double d = xi[0] - xj[0];
    const GptDual S = lx * lx + ly * ly;
    const GptDual iS = 1.0 / S;
    const GptDual k00 = dsqrt(2.0 * lx * ly * iS) * dexp(-(d * d) * iS);
    GptDual v = k00;
    if (a | b) {
        const GptDual Ax = lx1 / (2.0 * lx) - lx * lx1 * iS - (2.0 * d) * iS + (2.0 * d * d) * lx * lx1 * iS * iS;
        const GptDual Ay = ly1 / (2.0 * ly) - ly * ly1 * iS + (2.0 * d) * iS + (2.0 * d * d) * ly * ly1 * iS * iS;
        if (a && b) {
            const GptDual dAx = 2.0 * lx * lx1 * ly * ly1 * iS * iS + 2.0 * iS + (4.0 * d) * ly * ly1 * iS * iS -
                                (4.0 * d) * lx * lx1 * iS * iS - (8.0 * d * d) * lx * lx1 * ly * ly1 * iS * iS * iS;
            v = k00 * (Ax * Ay + dAx);
        } else if (a) {
            v = k00 * Ax;
        } else {
            v = k00 * Ay;
        }
    }
    const GptDual sf = hparam(cp, 0, hd);
    return sf * sf * v;
}

// dk/dtheta_hd for the non-SE kernels.  Not inlined on the device: only the gradient paths call it, and keeping it
// out of line leaves the register allocation of the value / SE paths untouched.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
double cov_hyper_eval(const CovParams& cp, const double* xi, const int32_t* ni, const double* xj, const int32_t* nj,
                      int hyper_deriv) {
    switch (cp.kid) {
        case GPT_KERNEL_MATERN52: return matern52_cov_dual(cp, xi, ni, xj, nj, hyper_deriv).d;
        case GPT_KERNEL_MATERN:
            if (hyper_deriv == 1) return NAN;  // d/dnu: not available (rejected on the host)
            return matern_cov_dual(cp, xi, ni, xj, nj, hyper_deriv).d;
        default: return gibbs_cov_dual(cp, xi, ni, xj, nj, hyper_deriv).d;
    }
}
