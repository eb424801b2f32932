"""The device covariance closed forms (gptools_b200/csrc/covfn.cuh), compiled for the HOST, against
the golden vectors of the unmodified reference.  CPU-only: this checks the exact source the CUDA
tile generators inline, so a formula error is caught before any GPU time is spent."""
import ctypes
import os
import subprocess
import warnings

import numpy as np
import pytest

from helpers import CASE_KERNEL, KERNEL_MATERN, KERNEL_SE, assert_close, load_golden

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gptools_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(CSRC, "libgptb200_hostcheck.so")
    src = os.path.join(CSRC, "hostcheck.cpp")
    hdrs = [os.path.join(CSRC, "covfn.cuh"), os.path.join(CSRC, "covfn_hyper.cuh")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-o", so, src, "-lm"])
    L = ctypes.CDLL(so)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    L.gpt_hostcheck_cov_pairs.argtypes = [ctypes.c_int] * 3 + [dp, ctypes.c_int, ctypes.c_long, dp, dp, ip, ip, dp]
    L.gpt_hostcheck_se_all.argtypes = [ctypes.c_int, dp, ctypes.c_long, dp, dp, ip, ip, dp]
    L.gpt_hostcheck_cov_dual.argtypes = [ctypes.c_int] * 3 + [dp, ctypes.c_int, ctypes.c_long, dp, dp, ip, ip, dp, dp]
    return L


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def pairs(lib, kid, params, Xi, Xj, ni, nj, hyper_deriv=-1):
    Xi = np.ascontiguousarray(Xi, dtype=np.float64)
    Xj = np.ascontiguousarray(Xj, dtype=np.float64)
    ni = np.ascontiguousarray(ni, dtype=np.int32)
    nj = np.ascontiguousarray(nj, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = np.empty(Xi.shape[0])
    rc = lib.gpt_hostcheck_cov_pairs(kid, Xi.shape[1], len(params), _ptr(params, ctypes.c_double), hyper_deriv,
                                     Xi.shape[0], _ptr(Xi, ctypes.c_double), _ptr(Xj, ctypes.c_double),
                                     _ptr(ni, ctypes.c_int32), _ptr(nj, ctypes.c_int32), _ptr(out, ctypes.c_double))
    assert rc == 0
    return out


def Kmat(lib, kid, params, X, n, hyper_deriv=-1):
    M = X.shape[0]
    return pairs(lib, kid, params, np.repeat(X, M, axis=0), np.tile(X, (M, 1)), np.repeat(n, M, axis=0),
                 np.tile(n, (M, 1)), hyper_deriv).reshape(M, M)


@pytest.mark.parametrize("D", [1, 2, 3])
def test_se_pairs_all_orders(lib, D):
    gd = load_golden("se_pairs_D%d" % D)
    a = (gd["params"], gd["Xi"], gd["Xj"], gd["ni"], gd["nj"])
    scale = np.abs(gd["val"]).max()
    assert_close(pairs(lib, KERNEL_SE, *a), gd["val"], rtol=2e-12, atol=1e-14 * scale, what="value")
    for p in range(D + 1):
        ref = gd["hd%d" % p]
        # H5: the reference divides by H_m(x); near its zeros the reference itself loses digits
        assert_close(pairs(lib, KERNEL_SE, *a, hyper_deriv=p), ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max(),
                     what="hyper_deriv %d" % p)
    # fused value + gradient evaluation used by the ll-gradient reduction
    out = np.empty((gd["Xi"].shape[0], 2 + D))
    Xi = np.ascontiguousarray(gd["Xi"])
    Xj = np.ascontiguousarray(gd["Xj"])
    ni = np.ascontiguousarray(gd["ni"], dtype=np.int32)
    nj = np.ascontiguousarray(gd["nj"], dtype=np.int32)
    pr = np.ascontiguousarray(gd["params"])
    lib.gpt_hostcheck_se_all(D, _ptr(pr, ctypes.c_double), Xi.shape[0], _ptr(Xi, ctypes.c_double),
                             _ptr(Xj, ctypes.c_double), _ptr(ni, ctypes.c_int32), _ptr(nj, ctypes.c_int32),
                             _ptr(out, ctypes.c_double))
    assert_close(out[:, 0], gd["val"], rtol=2e-12, atol=1e-14 * scale)
    for p in range(D + 1):
        ref = gd["hd%d" % p]
        assert_close(out[:, 1 + p], ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max(), what="all hd %d" % p)


@pytest.mark.parametrize("case", [c for c in CASE_KERNEL if c not in ("c1_synth200", "c2_small_matern52",
                                                                     "c2_small_matern_generic", "gibbs_c5_small")])
def test_K_entries(lib, case):
    gd = load_golden(case)
    kid = CASE_KERNEL[case]
    K = Kmat(lib, kid, gd["params"], gd["X"], gd["n"])
    if kid == KERNEL_MATERN:
        # outside the series zone the reference's kvp/Bell sum carries ~2e-10 abs round-off
        assert_close(K, gd["K"], rtol=1e-9, atol=1e-9 * np.abs(gd["K"]).max(), what=case)
    else:
        assert_close(K, gd["K"], rtol=1e-12, atol=1e-13 * np.abs(gd["K"]).max(), what=case)
    if case == "se2d_kat1":
        for p in range(3):
            dK = Kmat(lib, kid, gd["params"], gd["X"], gd["n"], hyper_deriv=p)
            assert_close(dK, gd["dK%d" % p], rtol=1e-10, atol=1e-13 * np.abs(gd["dK%d" % p]).max(), what="dK%d" % p)


def test_matern_generic_series_zone_is_emulated(lib):
    """inside 0 < y <= 5e-4 the reference's generic Matern is NOT the exact closed form for
    derivative orders >= 1; the device function must follow the reference there (and Matern52 must not).
    Checked against the pinned numpy restatement on explicit in-zone / origin / far pairs."""
    from oracle import gp_oracle as orc
    for nu in (2.5, 3.5):
        params = np.array([1.3, nu, 0.7])
        tau = np.array([0.0, 1e-4, 3e-3, 6e-3, 9e-3, 0.05, 0.5, 2.0])  # y = 2 nu tau^2 / l^2
        Xi = np.repeat(tau, 4)[:, None] + 0.25
        Xj = np.full_like(Xi, 0.25)
        ni = np.tile([0, 1, 0, 1], len(tau))[:, None]
        nj = np.tile([0, 0, 1, 1], len(tau))[:, None]
        want = orc.matern_pairs(Xi, Xj, ni, nj, params)
        got = pairs(lib, KERNEL_MATERN, params, Xi, Xj, ni, nj)
        assert_close(got, want, rtol=1e-9, atol=1e-9, what="generic Matern nu=%g" % nu)
    params = np.array([1.3, 2.5, 0.7])
    k52 = pairs(lib, 1, np.array([1.3, 0.7]), Xi, Xj, ni, nj)
    got = pairs(lib, KERNEL_MATERN, params, Xi, Xj, ni, nj)
    y = 5.0 * (tau / 0.7) ** 2
    zone = np.repeat((y > 0) & (y <= 5e-4), 4) & ((ni + nj)[:, 0] == 2)
    assert zone.any() and np.abs(got - k52)[zone].max() > 1e-4      # the series zone really differs
    assert np.abs(got - k52)[~np.repeat((y > 0) & (y <= 5e-4), 4)].max() < 1e-9


# ---- hyper-parameter derivatives of the Matern / Gibbs kernels (csrc/covfn_hyper.cuh) ----------
HYPERFD = {"hyperfd_matern52_1d": 1, "hyperfd_matern52_2d": 1, "hyperfd_matern_generic_nu2p5": 2,
           "hyperfd_matern_generic_nu3p5": 2, "hyperfd_matern_generic_nu1p5": 2, "hyperfd_matern_generic_2d": 2,
           "hyperfd_matern_generic_nu2p2": 2, "hyperfd_matern_generic_nu3p0": 2,
           "hyperfd_gibbs_direct": 3, "hyperfd_gibbs_T": 3}


def dual(lib, kid, params, X, n, hyper_deriv):
    M, D = X.shape
    Xi = np.ascontiguousarray(np.repeat(X, M, axis=0), dtype=np.float64)
    Xj = np.ascontiguousarray(np.tile(X, (M, 1)), dtype=np.float64)
    ni = np.ascontiguousarray(np.repeat(n, M, axis=0), dtype=np.int32)
    nj = np.ascontiguousarray(np.tile(n, (M, 1)), dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    v, d = np.empty(M * M), np.empty(M * M)
    rc = lib.gpt_hostcheck_cov_dual(kid, D, len(params), _ptr(params, ctypes.c_double), hyper_deriv, M * M,
                                    _ptr(Xi, ctypes.c_double), _ptr(Xj, ctypes.c_double), _ptr(ni, ctypes.c_int32),
                                    _ptr(nj, ctypes.c_int32), _ptr(v, ctypes.c_double), _ptr(d, ctypes.c_double))
    assert rc == 0
    return v.reshape(M, M), d.reshape(M, M)


@pytest.mark.parametrize("case", sorted(HYPERFD))
def test_hyper_derivative_closed_forms_match_reference_finite_differences(lib, case):
    """dk/dtheta from the dual-number closed forms against the Richardson central difference of the
    reference's own compute_Kij (the reference has no analytic form to compare with).  Tolerance: 1e-6 of
    the largest entry -- the finite difference, not the closed form, is the limiting side."""
    gd = load_golden(case)
    kid, X, n = HYPERFD[case], gd["X"], gd["n"].astype(np.int32)
    for q, p in enumerate(gd["idx"]):
        dK = Kmat(lib, kid, gd["params"], X, n, hyper_deriv=int(p))
        ref = gd["dK_fd"][q]
        assert_close(dK, ref, rtol=0.0, atol=1e-6 * np.abs(ref).max(), what="%s dK/dtheta_%d" % (case, p))


@pytest.mark.parametrize("case", sorted(HYPERFD))
def test_dual_value_component_is_the_value_path(lib, case):
    gd = load_golden(case)
    kid, X, n = HYPERFD[case], gd["X"], gd["n"].astype(np.int32)
    K = Kmat(lib, kid, gd["params"], X, n)
    for p in gd["idx"]:
        v, d = dual(lib, kid, gd["params"], X, n, int(p))
        assert_close(v, K, rtol=1e-14, atol=1e-300, what="%s dual value (seed %d)" % (case, p))
        assert np.array_equal(d, Kmat(lib, kid, gd["params"], X, n, hyper_deriv=int(p)))


# ---- GibbsKernel1d with other length-scale profiles (kernel id 4: l(x), l'(x) as extra point columns) --------------
GIBBS_PROFILES = {"double_tanh": "GibbsKernel1dDoubleTanh", "cubic_bucket": "GibbsKernel1dCubicBucket",
                  "quintic_bucket": "GibbsKernel1dQuinticBucket"}


@pytest.mark.parametrize("name", sorted(GIBBS_PROFILES))
def test_gibbs_profiles_host_functions_and_device_closed_form(lib, name):
    """(1) the host length-scale profile l(x), l'(x) against the reference's l_func (kernel/gibbs.py:508-760);
    (2) the device closed form fed with (x, l, l') against the reference's K incl. derivative observations."""
    import gptools_b200 as g
    gd = load_golden("gibbs_profile_" + name)
    k = getattr(g, GIBBS_PROFILES[name])(initial_params=gd["params"], param_bounds=[(0, 10)] * 8)
    Xa, na = k.device_points(gd["X"], gd["n"])
    assert_close(Xa[:, 1], gd["l"], rtol=1e-13, atol=1e-15, what="l(x)")
    assert_close(Xa[:, 2], gd["l1"], rtol=1e-12, atol=1e-13, what="l'(x)")
    K = Kmat(lib, 4, gd["params"][:1], Xa, na.astype(np.int32))
    assert_close(K, gd["K"], rtol=1e-11, atol=1e-13 * np.abs(gd["K"]).max(), what="K")
    dK = Kmat(lib, 4, gd["params"][:1], Xa, na.astype(np.int32), hyper_deriv=0)
    assert_close(dK, 2.0 * gd["K"] / gd["params"][0], rtol=1e-11, atol=1e-12, what="dK/dsigma_f")


def test_exp_gauss_profile_is_self_consistent():
    """The reference's exp_gauss_warp cannot run under Python 3 (float slice index, kernel/gibbs.py:833), so there is
    no golden: check the formula l = l0 exp(sum b exp(-(x-m)^2 / 2 s^2)) and that n = 1 is its derivative."""
    import gptools_b200 as g
    x = np.linspace(0, 1.2, 41)
    p = [0.4, 0.3, 0.9, 0.1, 0.05, -0.8, -1.5]
    l = g.exp_gauss_warp(x, 0, *p)
    want = 0.4 * np.exp(-0.8 * np.exp(-(x - 0.3) ** 2 / (2 * 0.1 ** 2)) - 1.5 * np.exp(-(x - 0.9) ** 2 / (2 * 0.05 ** 2)))
    assert_close(l, want, rtol=1e-14)
    h = 1e-6
    fd = (g.exp_gauss_warp(x + h, 0, *p) - g.exp_gauss_warp(x - h, 0, *p)) / (2 * h)
    assert_close(g.exp_gauss_warp(x, 1, *p), fd, rtol=1e-6, atol=1e-8)
    k = g.GibbsKernel1dExpGauss(2, initial_params=[1.5] + p, param_bounds=[(-10, 10)] * 8)
    assert k.num_params == 8 and k.device_descriptor()[0] == 4
    with pytest.raises(NotImplementedError):
        g.exp_gauss_warp(x, 2, *p)


def test_bessel_k_real_order_value_path_against_scipy(lib):
    """Round 2: K_nu of real order on the device (Temme series / Steed CF2 + recurrence).  The n = 0 value path is
    2^{1-nu}/Gamma(nu) y^{nu/2} K_nu(sqrt y) (kernel/matern.py:309-312): compare with scipy.special.kv over the
    orders and arguments the closed form sees, including integer orders and both branches (x <= 2, x > 2)."""
    import scipy.special
    r = np.concatenate([np.logspace(-6, np.log10(2.0), 40), np.linspace(2.0001, 40.0, 40)])
    for nu in (0.3, 0.7, 1.0, 1.2, 2.0, 2.2, 3.0, 4.3, 7.9):
        params = np.array([1.0, nu, 1.0])
        Xi = (r / np.sqrt(2.0 * nu))[:, None]           # y = 2 nu tau^2 / l^2 = r^2
        Xj = np.zeros_like(Xi)
        z = np.zeros(Xi.shape, dtype=int)
        got = pairs(lib, KERNEL_MATERN, params, Xi, Xj, z, z)
        want = 2.0 ** (1.0 - nu) / scipy.special.gamma(nu) * r ** nu * scipy.special.kv(nu, r)
        assert_close(got, want, rtol=2e-13, atol=1e-300, what="Matern value, nu = %g" % nu)


def test_matern_real_and_integer_order_match_the_oracle(lib):
    """Series zone, origin limits (+-inf / NaN included) and the regular branch for orders that are not
    half-integers, against the pinned restatement of utils.py:1429-1518."""
    from oracle import gp_oracle as orc
    tau = np.array([0.0, 1e-4, 3e-3, 6e-3, 9e-3, 0.05, 0.3, 0.5, 1.0, 2.0, 5.0, 15.0])
    Xi = np.repeat(tau, 4)[:, None] + 0.25
    Xj = np.full_like(Xi, 0.25)
    ni = np.tile([0, 1, 0, 1], len(tau))[:, None]
    nj = np.tile([0, 0, 1, 1], len(tau))[:, None]
    for nu in (0.7, 1.0, 2.0, 2.2, 3.0, 4.3):
        params = np.array([1.3, nu, 0.7])
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = orc.matern_pairs(Xi, Xj, ni, nj, params)
        got = pairs(lib, KERNEL_MATERN, params, Xi, Xj, ni, nj)
        fin = np.isfinite(want)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])
        assert_close(got[fin], want[fin], rtol=1e-9, atol=1e-9, what="generic Matern nu=%g" % nu)


# ---- hyper-derivatives pinned beyond finite-difference accuracy (round 2) --------------------------------------------
HYPERMP = {"hypermp_matern_nu2p5_1d": [KERNEL_MATERN, 1], "hypermp_matern_nu2p2_1d": [KERNEL_MATERN],
           "hypermp_matern_nu2p5_2d": [KERNEL_MATERN, 1], "hypermp_gibbs_tanh": [3]}


def hypermp_check(evaluate, case, rtol=1e-12):
    """``evaluate(kid, params, Xi, Xj, ni, nj, hyper_deriv)`` against the 40-digit mpmath derivatives of the reference's
    own covariance functions (tests/golden/make_golden.py: case_hyper_mp).  Matern 5/2 is checked twice: through the
    generic kernel and through Matern52Kernel (parameter vector without nu)."""
    gd = load_golden(case)
    for kid in HYPERMP[case]:
        params, idx = gd["params"], [int(i) for i in gd["idx"]]
        if kid == 1:                                      # Matern52: [sigma, l...]
            params = np.delete(params, 1)
            idx = [i if i == 0 else i - 1 for i in idx]
        scale = np.abs(gd["val"]).max()
        got = evaluate(kid, params, gd["Xi"], gd["Xj"], gd["ni"], gd["nj"], None)
        assert_close(got, gd["val"], rtol=rtol, atol=rtol * scale, what="%s value (kernel %d)" % (case, kid))
        for q, p in enumerate(idx):
            got = evaluate(kid, params, gd["Xi"], gd["Xj"], gd["ni"], gd["nj"], p)
            ref = gd["dval"][q]
            assert_close(got, ref, rtol=rtol, atol=rtol * np.abs(ref).max(), what="%s d/dtheta_%d (kernel %d)" % (case, p, kid))


@pytest.mark.parametrize("case", sorted(HYPERMP))
def test_hyper_derivatives_against_mpmath_derivatives_of_the_reference_functions(lib, case):
    hypermp_check(lambda kid, params, Xi, Xj, ni, nj, hd: pairs(lib, kid, params, Xi, Xj, ni, nj,
                                                               hyper_deriv=-1 if hd is None else hd), case)
