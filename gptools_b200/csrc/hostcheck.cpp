// Host build of the device covariance closed forms (covfn.cuh) for CPU-side unit tests.
// NOT part of the product path: only tests/test_covfn_host.py loads libgptb200_hostcheck.so,
// to check the exact source the CUDA kernels inline against the golden vectors without a GPU.
#include "covfn.cuh"

extern "C" {

int gpt_hostcheck_cov_pairs(int kid, int D, int nparams, const double* params, int hyper_deriv,
                            long npairs, const double* Xi, const double* Xj, const int32_t* ni,
                            const int32_t* nj, double* out) {
    if (D > GPT_MAX_DIM || nparams > GPT_MAX_PARAMS) return -1;
    CovParams cp;
    cov_params_init(cp, kid, D, nparams, params);
    for (long p = 0; p < npairs; p++)
        out[p] = cov_eval(cp, Xi + p * D, ni + p * D, Xj + p * D, nj + p * D, hyper_deriv);
    return 0;
}

// dual-number evaluation of the non-SE kernels: value component (must equal the value path bit for bit) and
// derivative component with respect to params[hyper_deriv]
int gpt_hostcheck_cov_dual(int kid, int D, int nparams, const double* params, int hyper_deriv, long npairs,
                           const double* Xi, const double* Xj, const int32_t* ni, const int32_t* nj, double* out_v,
                           double* out_d) {
    if (D > GPT_MAX_DIM || nparams > GPT_MAX_PARAMS || kid == GPT_KERNEL_SE) return -1;
    CovParams cp;
    cov_params_init(cp, kid, D, nparams, params);
    for (long p = 0; p < npairs; p++) {
        const double *xi = Xi + p * D, *xj = Xj + p * D;
        const int32_t *mi = ni + p * D, *mj = nj + p * D;
        GptDual r;
        if (kid == GPT_KERNEL_MATERN52) r = matern52_cov_dual(cp, xi, mi, xj, mj, hyper_deriv);
        else if (kid == GPT_KERNEL_MATERN) r = matern_cov_dual(cp, xi, mi, xj, mj, hyper_deriv);
        else r = gibbs_cov_dual(cp, xi, mi, xj, mj, hyper_deriv);
        out_v[p] = r.v;
        out_d[p] = r.d;
    }
    return 0;
}

// kernel algebra: sum of products of leaves (the device's comp_eval, on the host)
int gpt_hostcheck_composite_pairs(int D, int nleaf, const int32_t* kids, const int32_t* nps, int nterms,
                                  const int32_t* masks, const double* params, int hyper_deriv, long npairs,
                                  const double* Xi, const double* Xj, const int32_t* ni, const int32_t* nj, double* out) {
    if (D > GPT_MAX_DIM) return -1;
    CovComposite c;
    const int total = comp_init(c, D, nleaf, kids, nps, nterms, masks, params);
    if (total < 0) return -1;
    CovParams cp;
    cov_params_init(cp, GPT_KERNEL_COMPOSITE, D, total, params);
    cp.comp = &c;
    for (long p = 0; p < npairs; p++)
        out[p] = cov_eval(cp, Xi + p * D, ni + p * D, Xj + p * D, nj + p * D, hyper_deriv);
    return 0;
}

// out is (npairs, 2 + D): value, d/dsigma, d/dl_1..d/dl_D
int gpt_hostcheck_se_all(int D, const double* params, long npairs, const double* Xi, const double* Xj,
                         const int32_t* ni, const int32_t* nj, double* out) {
    if (D > GPT_MAX_DIM) return -1;
    CovParams cp;
    cov_params_init(cp, GPT_KERNEL_SE, D, D + 1, params);
    for (long p = 0; p < npairs; p++)
        se_cov_all(cp, Xi + p * D, ni + p * D, Xj + p * D, nj + p * D, out + p * (2 + D));
    return 0;
}
}
