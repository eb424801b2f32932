"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the golden vectors of
the unmodified reference and against the pinned numpy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): ll, gradient, predictive mean/std within 1e-9 relative; the
predictive variance is compared with the absolute floor 1e-9 * prior variance."""
import numpy as np
import pytest

from helpers import (CASE_KERNEL, HYPERFD_KERNEL, KERNEL_MATERN, KERNEL_MATERN52, KERNEL_SE, assert_close, load_golden,
                     richardson_fd)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from gptools_b200._lib import Device
    d = Device(0)
    yield d
    d.close()


def _T(gd):
    return gd["T"] if "T" in gd else None


def _setup(dev, case, gd, diag_factor=1e2):
    dev.set_data(gd["X"], gd["n"], gd["y"], gd["err_y"], _T(gd))
    dev.set_kernel(CASE_KERNEL[case], len(gd["params"]), diag_factor)


@pytest.mark.parametrize("D", [1, 2, 3])
def test_cov_pairs_se(dev, D):
    gd = load_golden("se_pairs_D%d" % D)
    a = (gd["params"], gd["Xi"], gd["Xj"], gd["ni"], gd["nj"])
    scale = np.abs(gd["val"]).max()
    assert_close(dev.cov_pairs(KERNEL_SE, *a), gd["val"], rtol=2e-12, atol=1e-14 * scale, what="value")
    for p in range(D + 1):
        ref = gd["hd%d" % p]
        assert_close(dev.cov_pairs(KERNEL_SE, *a, hyper_deriv=p), ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max(),
                     what="hyper_deriv %d" % p)


@pytest.mark.parametrize("case", [c for c in CASE_KERNEL if c not in ("c1_synth200", "c2_small_matern52",
                                                                     "c2_small_matern_generic", "gibbs_c5_small")])
def test_compute_Kij(dev, case):
    gd = load_golden(case)
    kid = CASE_KERNEL[case]
    K = dev.compute_Kij(kid, gd["params"], gd["X"], gd["n"])
    tol = dict(rtol=1e-9, atol=1e-9 * np.abs(gd["K"]).max()) if kid == KERNEL_MATERN else \
        dict(rtol=1e-12, atol=1e-13 * np.abs(gd["K"]).max())
    assert_close(K, gd["K"], what=case, **tol)
    if case == "se2d_kat1":
        for p in range(3):
            dK = dev.compute_Kij(kid, gd["params"], gd["X"], gd["n"], hyper_deriv=p)
            assert_close(dK, gd["dK%d" % p], rtol=1e-10, atol=1e-13 * np.abs(gd["dK%d" % p]).max(), what="dK%d" % p)
        Ks = dev.compute_Kij(kid, gd["params"], gd["X"], gd["n"], gd["Xs"], np.zeros((4, 2), dtype=int))
        assert_close(Ks, gd["Kstar"], rtol=1e-12, atol=1e-14, what="Kstar")


@pytest.mark.parametrize("case", sorted(CASE_KERNEL))
def test_ll_alpha_L(dev, case):
    gd = load_golden(case)
    _setup(dev, case, gd)
    ns = float(gd["noise_sigma"][0]) if "noise_sigma" in gd else 0.0
    ll, _, status = dev.ll(gd["params"], ns)
    assert status == 0
    assert_close(ll + gd["log_prior"], gd["ll"], rtol=1e-9, what=case + " ll")
    alpha = dev.get_alpha()
    # generic Matern: the reference's K entries carry ~2e-10 abs round-off from its kvp / Bell-polynomial
    # sums; alpha = K^-1 y amplifies that by cond(K), so alpha/L can only agree to ~1e-6
    tol = 2e-6 if CASE_KERNEL[case] == KERNEL_MATERN else 1e-9
    assert_close(alpha, gd["alpha"], rtol=tol, atol=tol * np.abs(gd["alpha"]).max(), what=case + " alpha")
    if "L" in gd:
        L = dev.get_L()
        assert_close(L, gd["L"], rtol=tol, atol=tol * np.abs(gd["L"]).max(), what=case + " L")


@pytest.mark.parametrize("case", ["se2d_kat1", "se_diagnoise", "demo_c1_kat4", "c1_synth200"])
def test_ll_gradient(dev, case):
    from oracle import gp_oracle as orc
    gd = load_golden(case)
    _setup(dev, case, gd)
    nk = len(gd["params"])
    ns = float(gd["noise_sigma"][0]) if "noise_sigma" in gd else 0.0
    idx = list(range(nk)) + ([nk] if ns else [])
    ll, grad, status = dev.ll(gd["params"], ns, grad_idx=idx)
    assert status == 0
    if case in ("se2d_kat1", "se_diagnoise"):      # uniform priors: golden gradient == likelihood gradient
        assert_close(grad, gd["ll_deriv"], rtol=1e-9, atol=1e-9 * np.abs(gd["ll_deriv"]).max(), what=case + " grad")
    else:                                          # non-uniform prior in the golden: compare with the oracle
        r = orc.compute_K_L_alpha_ll(KERNEL_SE, gd["params"], gd["X"], gd["n"], gd["y"], gd["err_y"], grad_idx=idx)
        assert_close(grad, r["ll_deriv"], rtol=1e-8, atol=1e-9 * np.abs(r["ll_deriv"]).max(), what=case + " grad")


@pytest.mark.parametrize("case", ["se2d_kat1", "matern52_kat2", "gibbs_kat3", "gibbs_c5_small", "se_diagnoise",
                                  "matern_generic_nu2p5", "matern_generic_nu3p5", "matern_generic_nu2p2",
                                  "matern_generic_nu3p0"])
def test_predict_full(dev, case):
    gd = load_golden(case)
    _setup(dev, case, gd)
    kid = CASE_KERNEL[case]
    ns = float(gd["noise_sigma"][0]) if "noise_sigma" in gd else 0.0
    ll, _, status = dev.ll(gd["params"], ns)
    assert status == 0
    Xs = np.atleast_2d(gd["Xs"])
    if Xs.shape[0] == 1 and gd["X"].shape[1] == 1:
        Xs = Xs.T
    z = np.zeros(Xs.shape, dtype=int)
    mean, var, cov = dev.predict(Xs, z, want_cov=True)
    assert_close(mean, gd["mean"], rtol=1e-9, atol=1e-9 * np.abs(gd["mean"]).max(), what=case + " mean")
    prior = np.diag(dev.compute_Kij(kid, gd["params"], Xs, z))
    assert np.all(np.abs(cov - gd["cov"]) <= 1e-9 * prior.max()), case + " cov"
    assert np.all(np.abs(var - gd["std"] ** 2) <= 1e-9 * prior), case + " var"
    mean2, var2, _ = dev.predict(Xs, z, want_var=True)
    assert_close(mean2, mean, rtol=0, atol=0)
    assert np.all(np.abs(var2 - np.diag(cov)) <= 1e-12 * prior)
    if "mean_d1" in gd:
        o = np.zeros(Xs.shape, dtype=int)
        o[:, 0] = 1                      # d/dx_1 (the goldens use n = [1, 0, ...])
        m1, v1, c1 = dev.predict(Xs, o, want_cov=True)
        assert_close(m1, gd["mean_d1"], rtol=1e-9, atol=1e-9 * np.abs(gd["mean_d1"]).max(), what=case + " mean_d1")
        prior1 = np.diag(dev.compute_Kij(kid, gd["params"], Xs, o))
        assert np.all(np.abs(c1 - gd["cov_d1"]) <= 1e-9 * prior1.max())
    if "draw" in gd:
        jit = 1e3 * 2.220446049250313e-16
        samp, st = dev.draw_sample(gd["mean"], gd["cov"], gd["rand_vars"], jit)
        assert st == 0
        # The sample is mean + chol(cov + jitter) u.  For the 40-point posterior covariance cond(cov + jitter)
        # ~ 1e12, so two correct Cholesky factors differ far above 1e-9 in their trailing columns: compare the
        # samples at the conditioning-limited level and pin the factor by its defining property instead.
        tol = 1e-8 if gd["cov"].shape[0] <= 3 else 1e-5
        assert_close(samp, gd["draw"], rtol=tol, atol=tol * np.abs(gd["draw"]).max(), what=case + " draw")
        Ms = gd["cov"].shape[0]
        Lc, st = dev.draw_sample(np.zeros(Ms), gd["cov"], np.eye(Ms), jit)
        assert st == 0 and np.all(np.triu(Lc, 1) == 0.0)
        target = gd["cov"] + jit * np.eye(Ms)
        assert np.abs(Lc @ Lc.T - target).max() <= 1e-13 * np.abs(target).max()


@pytest.mark.parametrize("case", ["demo_c1_kat4", "c1_synth200", "c2_small_matern52", "c2_small_matern_generic"])
def test_predict_mean_std_many_points(dev, case):
    gd = load_golden(case)
    _setup(dev, case, gd)
    kid = CASE_KERNEL[case]
    ll, _, status = dev.ll(gd["params"], 0.0)
    assert status == 0
    Xs = gd["Xs"][:, None]
    for order, suffix in ((0, ""), (1, "_d1")):
        ns = np.full(Xs.shape, order, dtype=int)
        mean, var, _ = dev.predict(Xs, ns)
        scale = np.abs(gd["mean" + suffix]).max()
        assert_close(mean, gd["mean" + suffix], rtol=1e-9, atol=1e-9 * scale, what=case + " mean" + suffix)
        prior = dev.compute_Kij(kid, gd["params"], Xs[:1], ns[:1])[0, 0]
        ref_var = gd["std" + suffix] ** 2
        ok = np.abs(var - ref_var) <= 1e-9 * prior
        ok |= np.isnan(ref_var) & (var < 1e-9 * prior)   # reference sqrt(negative) -> NaN
        assert ok.all(), case + " var" + suffix


def test_batched_c3_kat5(dev):
    """Headline path: batched ll + gradient on the config-3 problem vs the reference's values."""
    gd = load_golden("c3_kat5")
    dev.set_data(gd["X"], gd["n"], gd["y"], gd["err_y"])
    dev.set_kernel(KERNEL_SE, 3, 1e2)
    thetas = np.hstack([gd["theta"], np.zeros((len(gd["theta"]), 1))])
    ll, grad, status = dev.ll_batched(thetas, grad_idx=[0, 1, 2])
    assert (status == 0).all()
    assert_close(ll + gd["log_prior"], gd["ll"], rtol=1e-9, what="batched ll")
    assert_close(grad, gd["ll_deriv"], rtol=1e-9, atol=1e-9 * np.abs(gd["ll_deriv"]).max(), what="batched grad")
    # ll-only launch and alpha output
    ll2, g2, st2, alpha = dev.ll_batched(thetas[:3], return_alpha=True)
    assert g2 is None
    assert_close(ll2, ll[:3], rtol=1e-13)
    assert_close(alpha[0], gd["alpha0"], rtol=1e-9, atol=1e-9 * np.abs(gd["alpha0"]).max(), what="alpha")
    # single-theta path agrees with the batched path
    for b in (0, 7):
        l1, g1, s1 = dev.ll(gd["theta"][b], 0.0, grad_idx=[0, 1, 2])
        assert s1 == 0
        assert_close(l1, ll[b], rtol=1e-11)
        assert_close(g1, grad[b], rtol=1e-9, atol=1e-9 * np.abs(grad[b]).max())
    # predict at theta[0] (KAT-5)
    dev.ll(gd["theta"][0], 0.0)
    mean, var, _ = dev.predict(gd["Xs"], np.zeros((5, 2), dtype=int))
    assert_close(mean, gd["mean"], rtol=1e-9, atol=1e-9)
    assert np.all(np.abs(var - gd["std"] ** 2) <= 1e-9 * gd["theta"][0][0] ** 2)


@pytest.mark.parametrize("case", ["se2d_kat1", "matern52_kat2", "matern_generic_nu2p5", "c1_synth200",
                                  "c2_small_matern52", "se_diagnoise"])
def test_batched_matches_golden_small_and_ragged(dev, case):
    """M not a multiple of the 64-row tile, several kernels, noise kernel, B = 1 and B > #CTAs."""
    gd = load_golden(case)
    if "T" in gd:
        pytest.skip("batched path has no T")
    _setup(dev, case, gd)
    ns = float(gd["noise_sigma"][0]) if "noise_sigma" in gd else 0.0
    th = np.concatenate([gd["params"], [ns]])[None, :]
    ll, _, status = dev.ll_batched(th)
    assert status[0] == 0
    assert_close(ll[0] + gd["log_prior"], gd["ll"], rtol=1e-9, what=case + " batched ll")
    if CASE_KERNEL[case] == KERNEL_SE and case in ("se2d_kat1", "se_diagnoise"):
        idx = list(range(len(gd["params"]))) + ([len(gd["params"])] if ns else [])
        B = 700
        ths = np.repeat(th, B, axis=0)
        llB, gB, stB = dev.ll_batched(ths, grad_idx=idx)
        assert (stB == 0).all()
        assert np.all(llB == llB[0]) and np.all(gB == gB[0]), "identical thetas must give identical bits"
        assert_close(gB[0], gd["ll_deriv"], rtol=1e-9, atol=1e-9 * np.abs(gd["ll_deriv"]).max(), what=case + " grad")


def test_not_positive_definite_is_reported_per_theta(dev):
    """Duplicate noiseless points => singular K: the reference raises LinAlgError (gaussian_process.py:1452),
    mapped to +inf by update_hyperparameters; here a LAPACK-style status per theta."""
    rs = np.random.RandomState(0)
    X = rs.rand(40, 1)
    X[7] = X[3]
    X[30] = X[3]
    y = np.sin(X[:, 0])
    dev.set_data(X, np.zeros((40, 1), dtype=int), y, np.zeros(40))
    dev.set_kernel(KERNEL_SE, 2, 0.0)
    ll, g, st = dev.ll(np.array([1.0, 0.5]), 0.0)
    assert st > 0
    good = np.array([1.0, 0.5, 0.3])
    bad = np.array([1.0, 0.5, 0.0])
    llb, gb, stb = dev.ll_batched(np.stack([good, bad, good]), grad_idx=[0, 1])
    assert stb[0] == 0 and stb[2] == 0 and stb[1] > 0
    assert llb[0] == llb[2] and np.isfinite(llb[0])
    assert np.all(gb[1] == 0.0)


def test_multiblock_single_path_vs_oracle(dev):
    """M = 700 (6 blocks of 128 with ragged tail): blocked potrf / trtri / lauum / fused gradient vs numpy."""
    from oracle import gp_oracle as orc
    rs = np.random.RandomState(5)
    Xv = rs.rand(400, 2)
    Xd = rs.rand(300, 2)
    f = lambda x: np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    X = np.vstack([Xv, Xd])
    n = np.vstack([np.zeros((400, 2), dtype=int), np.tile([1, 0], (300, 1))])
    y = np.concatenate([f(Xv), 3 * np.cos(3 * Xd[:, 0]) * np.cos(2 * Xd[:, 1])]) + 0.05 * rs.randn(700)
    err = np.full(700, 0.05)
    params = np.array([1.1, 0.35, 0.45])
    r = orc.compute_K_L_alpha_ll(KERNEL_SE, params, X, n, y, err, grad_idx=[0, 1, 2])
    dev.set_data(X, n, y, err)
    dev.set_kernel(KERNEL_SE, 3, 1e2)
    ll, grad, st = dev.ll(params, 0.0, grad_idx=[0, 1, 2])
    assert st == 0
    assert_close(ll, r["ll"], rtol=1e-9)
    assert_close(grad, r["ll_deriv"], rtol=1e-9, atol=1e-9 * np.abs(r["ll_deriv"]).max())
    assert_close(dev.get_alpha(), r["alpha"].ravel(), rtol=1e-8, atol=1e-8 * np.abs(r["alpha"]).max())
    llb, gb, stb = dev.ll_batched(np.concatenate([params, [0.0]])[None, :], grad_idx=[0, 1, 2])
    assert stb[0] == 0
    assert_close(llb[0], r["ll"], rtol=1e-9)
    assert_close(gb[0], r["ll_deriv"], rtol=1e-9, atol=1e-9 * np.abs(r["ll_deriv"]).max())
    Xs = rs.rand(300, 2)
    ns = np.zeros((300, 2), dtype=int)
    mean_o, std_o = orc.predict_blocked(KERNEL_SE, params, X, n, r["L"], r["alpha"], Xs, ns, block=150)
    mean, var, _ = dev.predict(Xs, ns)
    assert_close(mean, mean_o, rtol=1e-9, atol=1e-9)
    assert np.all(np.abs(var - std_o ** 2) <= 1e-9 * params[0] ** 2)


# ---- hyper-parameter derivatives of the Matern / Gibbs kernels ---------------------------------
# The reference raises NotImplementedError for these (kernel/matern.py:543, core.py:723, gibbs.py:319), so the
# goldens are Richardson central differences of the reference's own ll and K (tests/golden/make_golden.py:
# case_hyperfd).  Tolerance 1e-6 relative to the largest gradient entry: the finite difference is the limiting side
# (1e-5 for the generic Matern kernel, whose reference K carries ~2e-10 of kvp round-off).
@pytest.mark.parametrize("case", sorted(HYPERFD_KERNEL))
def test_hyper_gradient_matern_gibbs_vs_reference_fd(dev, case):
    gd = load_golden(case)
    kid = HYPERFD_KERNEL[case]
    idx = [int(i) for i in gd["idx"]]
    for q, p in enumerate(idx):
        dK = dev.compute_Kij(kid, gd["params"], gd["X"], gd["n"], hyper_deriv=p)
        assert_close(dK, gd["dK_fd"][q], rtol=0.0, atol=1e-6 * np.abs(gd["dK_fd"][q]).max(), what="%s dK%d" % (case, p))
    dev.set_data(gd["X"], gd["n"], gd["y"], gd["err_y"], _T(gd))
    dev.set_kernel(kid, len(gd["params"]), 1e2)
    ll, grad, status = dev.ll(gd["params"], 0.0, grad_idx=idx)
    assert status == 0
    tol = 1e-5 if kid == KERNEL_MATERN else 1e-6
    assert_close(grad, gd["ll_grad_fd"], rtol=0.0, atol=tol * np.abs(gd["ll_grad_fd"]).max(), what=case + " ll gradient")
    if "T" not in gd:
        # the batched kernel: same numbers as the single-theta path, identical bits for identical thetas
        th = np.concatenate([gd["params"], [0.0]])[None, :]
        llB, gB, stB = dev.ll_batched(np.repeat(th, 5, axis=0), grad_idx=idx)
        assert (stB == 0).all() and np.all(gB == gB[0])
        assert_close(llB[0], ll, rtol=1e-10, what=case + " batched ll")
        assert_close(gB[0], grad, rtol=1e-8, atol=1e-9 * np.abs(grad).max(), what=case + " batched gradient")
    if kid == KERNEL_MATERN:
        with pytest.raises(Exception):
            dev.ll(gd["params"], 0.0, grad_idx=[1])      # d/dnu is not available


def test_hyper_gradient_c2_shape_vs_oracle_fd(dev):
    """Config-2 shape at a size the oracle handles in seconds (300 locations, value + derivative, Matern 5/2):
    multi-block single path and multi-tile batched path against the finite difference of the pinned oracle."""
    from oracle import gp_oracle as orc
    rs = np.random.RandomState(0)
    X = np.sort(rs.rand(300)) * 10
    Xa = np.concatenate([X, X])[:, None]
    n = np.concatenate([np.zeros(300, int), np.ones(300, int)])[:, None]
    y = np.concatenate([np.sin(X), np.cos(X)]) + 0.05 * rs.randn(600)
    err = np.full(600, 0.05)
    th = np.array([1.0, 0.8])

    def ll_of(t):
        return orc.compute_K_L_alpha_ll(orc.KERNEL_MATERN52, t, Xa, n, y, err, None, 0.0, 1e2)["ll"]

    fd = np.array([richardson_fd(ll_of, th, i, 2e-3 * th[i]) for i in range(2)])
    dev.set_data(Xa, n, y, err)
    dev.set_kernel(KERNEL_MATERN52, 2, 1e2)
    ll, grad, st = dev.ll(th, 0.0, grad_idx=[0, 1])
    assert st == 0
    assert_close(ll, ll_of(th), rtol=1e-9, what="ll")
    assert_close(grad, fd, rtol=0.0, atol=1e-6 * np.abs(fd).max(), what="single-path gradient")
    ths = np.array([[1.0, 0.8, 0.0], [1.3, 0.6, 0.0], [0.7, 1.1, 0.0]])
    llB, gB, stB = dev.ll_batched(ths, grad_idx=[0, 1])
    assert (stB == 0).all()
    assert_close(gB[0], grad, rtol=1e-8, atol=1e-9 * np.abs(grad).max(), what="batched gradient")
    fd1 = np.array([richardson_fd(ll_of, ths[1, :2], i, 2e-3 * ths[1, i]) for i in range(2)])
    assert_close(gB[1], fd1, rtol=0.0, atol=1e-6 * np.abs(fd1).max(), what="batched gradient, theta 1")


@pytest.mark.parametrize("M", [1, 2, 63, 64, 65, 127, 129])
def test_tiny_and_ragged_sizes_all_paths(dev, M):
    """Sizes around the 64-row batched tile and the 128-row block of the single path, down to one observation:
    ll, gradient, alpha, predictive mean and variance against the pinned oracle; empty inputs are no-ops."""
    from oracle import gp_oracle as orc
    rs = np.random.RandomState(M)
    X = rs.rand(M, 1)
    n = np.zeros((M, 1), dtype=int)
    y = np.sin(3 * X[:, 0])
    err = np.full(M, 0.1)
    th = np.array([1.0, 0.4])
    ref = orc.compute_K_L_alpha_ll(orc.KERNEL_SE, th, X, n, y, err, None, 0.0, 1e2, grad_idx=[0, 1])
    dev.set_data(X, n, y, err)
    dev.set_kernel(KERNEL_SE, 2, 1e2)
    ll, g, st = dev.ll(th, 0.0, grad_idx=[0, 1])
    assert st == 0
    assert_close(ll, ref["ll"], rtol=1e-9, what="ll")
    assert_close(g, ref["ll_deriv"], rtol=1e-9, atol=1e-9 * np.abs(ref["ll_deriv"]).max(), what="grad")
    assert_close(dev.get_alpha(), np.ravel(ref["alpha"]), rtol=1e-9, atol=1e-9 * np.abs(ref["alpha"]).max(), what="alpha")
    llb, gb, stb = dev.ll_batched(np.array([[1.0, 0.4, 0.0]] * 3), grad_idx=[0, 1])
    assert (stb == 0).all()
    assert_close(llb, np.full(3, ref["ll"]), rtol=1e-9, what="batched ll")
    assert_close(gb[2], ref["ll_deriv"], rtol=1e-9, atol=1e-9 * np.abs(ref["ll_deriv"]).max(), what="batched grad")
    Xs = rs.rand(5, 1)
    ns = np.zeros((5, 1), dtype=int)
    dev.ll(th, 0.0)
    mean, var, _ = dev.predict(Xs, ns, want_var=True)
    pm, ps, pc = orc.predict(orc.KERNEL_SE, th, X, n, ref["L"], ref["alpha"], Xs, ns)
    assert_close(mean, pm, rtol=1e-9, atol=1e-9, what="mean")
    assert np.all(np.abs(var - np.diag(pc)) <= 1e-9 * th[0] ** 2)
    # empty test set / empty batch: accepted, nothing written
    m0, v0, _ = dev.predict(np.zeros((0, 1)), np.zeros((0, 1), dtype=int), want_var=True)
    assert m0.shape == (0,) and v0.shape == (0,)
    l0, g0, s0 = dev.ll_batched(np.zeros((0, 3)), grad_idx=[0, 1])
    assert l0.shape == (0,) and g0.shape == (0, 2) and s0.shape == (0,)


def test_T_path_gradient_multiblock_vs_oracle(dev):
    """Transformed observations y = T f with N = 300 latent points (three 128-blocks) and M = 40 observations:
    ll, analytic gradient (SE: the reference's own formula through the oracle; Gibbs-tanh: finite differences of the
    oracle ll) and prediction with T, multi-block in both the latent and the observation dimension."""
    from oracle import gp_oracle as orc
    rs = np.random.RandomState(5)
    N, Mo, W = 300, 40, 30
    Xq = np.linspace(0, 1.1, N)[:, None]
    T = np.zeros((Mo, N))
    for i, s in enumerate(rs.randint(0, N - W, size=Mo)):
        T[i, s:s + W] = 1.1 / N
    y = rs.rand(Mo) * 0.3 + 0.1
    err = np.full(Mo, 0.02)
    n = np.zeros((N, 1), dtype=int)
    dev.set_data(Xq, n, y, err, T)
    # SE kernel: analytic gradient
    th = np.array([1.3, 0.25])
    ref = orc.compute_K_L_alpha_ll(orc.KERNEL_SE, th, Xq, n, y, err, T, 0.0, 1e2, grad_idx=[0, 1])
    dev.set_kernel(KERNEL_SE, 2, 1e2)
    ll, grad, st = dev.ll(th, 0.0, grad_idx=[0, 1])
    assert st == 0
    assert_close(ll, ref["ll"], rtol=1e-9, what="ll")
    assert_close(grad, ref["ll_deriv"], rtol=1e-8, atol=1e-9 * np.abs(ref["ll_deriv"]).max(), what="gradient")
    Xs = np.linspace(0, 1.1, 150)[:, None]
    ns = np.zeros((150, 1), dtype=int)
    mean, var, _ = dev.predict(Xs, ns, want_var=True)
    pm, ps, pc = orc.predict(orc.KERNEL_SE, th, Xq, n, ref["L"], ref["alpha"], Xs, ns, T=T)
    assert_close(mean, pm, rtol=1e-9, atol=1e-9, what="mean")
    assert np.all(np.abs(var - np.diag(pc)) <= 1e-9 * th[0] ** 2)
    # Gibbs-tanh: dual-number hyper-derivatives through T
    from helpers import KERNEL_GIBBS_TANH
    tg = np.array([1.5, 0.6, 0.1, 0.05, 0.9])
    dev.set_kernel(KERNEL_GIBBS_TANH, 5, 1e2)
    llg, gg, st = dev.ll(tg, 0.0, grad_idx=[0, 1, 2, 3, 4])
    assert st == 0

    def ll_of(t):
        return orc.compute_K_L_alpha_ll(orc.KERNEL_GIBBS_TANH, t, Xq, n, y, err, T, 0.0, 1e2)["ll"]

    assert_close(llg, ll_of(tg), rtol=1e-9, what="gibbs ll")
    fd = np.array([richardson_fd(ll_of, tg, i, 1e-3 * tg[i]) for i in range(5)])
    assert_close(gg, fd, rtol=0.0, atol=2e-6 * np.abs(fd).max(), what="gibbs gradient through T")


@pytest.mark.parametrize("trial", range(16))
def test_randomised_paths_agree_with_oracle(dev, trial):
    """Seeded sweep over input dimension (1..4: the SE fast paths and the generic one), derivative orders up to 2 (the
    branch-free low-order forms and the general Hermite recurrence) and sizes around the tile edges: single-theta path,
    batched path and prediction against the pinned oracle, all at the north-star tolerance."""
    from oracle import gp_oracle as orc
    rs = np.random.RandomState(1000 + trial)
    D = [1, 2, 3, 4][trial % 4]
    M = int(rs.choice([37, 64, 100, 129, 200, 260]))
    maxord = [1, 2][(trial // 4) % 2]
    X = rs.rand(M, D)
    n = np.zeros((M, D), dtype=int)
    for i in range(M // 3, M):
        n[i, rs.randint(D)] = rs.randint(0, maxord + 1)
    y = rs.randn(M)
    err = np.full(M, 0.2)
    th = np.concatenate([[1.0 + 0.5 * rs.rand()], 0.3 + 0.4 * rs.rand(D)])
    idx = list(range(D + 1))
    ref = orc.compute_K_L_alpha_ll(orc.KERNEL_SE, th, X, n, y, err, None, 0.0, 1e2, grad_idx=idx)
    dev.set_data(X, n, y, err)
    dev.set_kernel(KERNEL_SE, D + 1, 1e2)
    ll, g, st = dev.ll(th, 0.0, grad_idx=idx)
    llb, gb, stb = dev.ll_batched(np.array([list(th) + [0.0]] * 2), grad_idx=idx)
    assert st == 0 and (stb == 0).all()
    gs = np.abs(ref["ll_deriv"]).max()
    assert_close(ll, ref["ll"], rtol=1e-9, what="ll")
    assert_close(g, ref["ll_deriv"], rtol=0.0, atol=1e-9 * gs, what="gradient")
    assert_close(llb[0], ref["ll"], rtol=1e-9, what="batched ll")
    assert_close(gb[0], ref["ll_deriv"], rtol=0.0, atol=1e-9 * gs, what="batched gradient")
    Xs = rs.rand(11, D)
    ns = np.zeros((11, D), dtype=int)
    ns[::3, 0] = 1
    dev.ll(th, 0.0)
    m, v, _ = dev.predict(Xs, ns, want_var=True)
    pm, ps, pc = orc.predict(orc.KERNEL_SE, th, X, n, ref["L"], ref["alpha"], Xs, ns)
    assert_close(m, pm, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(pm).max()), what="mean")
    assert np.all(np.abs(v - np.diag(pc)) <= 1e-9 * np.abs(np.diag(pc)).max())


@pytest.mark.parametrize("trial", range(12))
def test_randomised_matern_gibbs_paths_agree_with_oracle(dev, trial):
    """The same sweep for the other kernels: Matern-5/2 (D = 1..3), generic Matern (nu = 1/2 .. 7/2, D = 1, 2) and
    Gibbs-tanh, value and first-derivative observations, sizes around the tile edges."""
    from oracle import gp_oracle as orc
    from helpers import KERNEL_GIBBS_TANH
    rs = np.random.RandomState(2000 + trial)
    fam = trial % 3
    M = int(rs.choice([37, 64, 100, 129, 200]))
    if fam == 0:
        D = 1 + (trial // 3) % 3
        kid, okid = KERNEL_MATERN52, orc.KERNEL_MATERN52
        th = np.concatenate([[1.0 + 0.5 * rs.rand()], 0.4 + 0.5 * rs.rand(D)])
        tol = 1e-9
    elif fam == 1:
        D = 1 + (trial // 3) % 2
        nu = [2.5, 3.5, 1.5, 0.5][(trial // 3) % 4]
        kid, okid = KERNEL_MATERN, orc.KERNEL_MATERN
        th = np.concatenate([[1.0 + 0.5 * rs.rand(), nu], 0.4 + 0.5 * rs.rand(D)])
        tol = 2e-6       # the oracle mirrors the reference's kvp round-off (see test_ll_alpha_L)
    else:
        D = 1
        kid, okid = KERNEL_GIBBS_TANH, orc.KERNEL_GIBBS_TANH
        th = np.array([1.2 + 0.5 * rs.rand(), 0.5, 0.15, 0.08, 0.6])
        tol = 1e-9
    X = rs.rand(M, D)
    n = np.zeros((M, D), dtype=int)
    if not (fam == 1 and th[1] < 2):            # nu <= 3/2 has no usable derivative observations
        for i in range(2 * M // 3, M):
            n[i, rs.randint(D)] = 1
    y = rs.randn(M)
    err = np.full(M, 0.2)
    ref = orc.compute_K_L_alpha_ll(okid, th, X, n, y, err, None, 0.0, 1e2)
    dev.set_data(X, n, y, err)
    dev.set_kernel(kid, len(th), 1e2)
    ll, _, st = dev.ll(th, 0.0)
    llb, _, stb = dev.ll_batched(np.array([list(th) + [0.0]] * 2))
    assert st == 0 and (stb == 0).all()
    assert_close(ll, ref["ll"], rtol=max(tol, 1e-9), what="ll")
    assert_close(llb[0], ref["ll"], rtol=max(tol, 1e-9), what="batched ll")
    assert_close(dev.get_alpha(), np.ravel(ref["alpha"]), rtol=0.0, atol=max(tol, 1e-8) * np.abs(ref["alpha"]).max(),
                 what="alpha")
    Xs = rs.rand(9, D)
    ns = np.zeros((9, D), dtype=int)
    m, v, _ = dev.predict(Xs, ns, want_var=True)
    pm, ps, pc = orc.predict(okid, th, X, n, ref["L"], ref["alpha"], Xs, ns)
    assert_close(m, pm, rtol=0.0, atol=max(tol, 1e-9) * max(1.0, np.abs(pm).max()), what="mean")
    assert np.all(np.abs(v - np.diag(pc)) <= max(tol, 1e-9) * th[0] ** 2)


@pytest.mark.parametrize("pair_min", ["1", "100000"])
@pytest.mark.parametrize("M", [130, 300, 700, 800, 1100])
def test_blocked_cholesky_variants_vs_oracle(dev, M, pair_min, monkeypatch):
    """Both variants of the blocked Cholesky -- paired rank-256 trailing updates (default from 80 blocks) and the
    unpaired one -- forced onto the same small problems (2 .. 9 blocks, odd and even counts): ll, gradient, alpha,
    L and the predictive variance against the pinned oracle."""
    from oracle import gp_oracle as orc
    monkeypatch.setenv("GPT_POTRF_PAIR_MIN", pair_min)
    rs = np.random.RandomState(M)
    X = rs.rand(M, 2)
    n = np.zeros((M, 2), dtype=int)
    n[M // 2:, 0] = 1
    y = rs.randn(M)
    err = np.full(M, 0.3)
    th = np.array([1.1, 0.35, 0.5])
    ref = orc.compute_K_L_alpha_ll(orc.KERNEL_SE, th, X, n, y, err, None, 0.0, 1e2, grad_idx=[0, 1, 2])
    dev.set_data(X, n, y, err)
    dev.set_kernel(KERNEL_SE, 3, 1e2)
    ll, g, st = dev.ll(th, 0.0, grad_idx=[0, 1, 2])
    assert st == 0
    assert_close(ll, ref["ll"], rtol=1e-9, what="ll")
    assert_close(g, ref["ll_deriv"], rtol=0.0, atol=1e-9 * np.abs(ref["ll_deriv"]).max(), what="gradient")
    assert_close(dev.get_alpha(), np.ravel(ref["alpha"]), rtol=0.0, atol=1e-9 * np.abs(ref["alpha"]).max(), what="alpha")
    assert_close(dev.get_L(), ref["L"], rtol=0.0, atol=1e-10 * np.abs(ref["L"]).max(), what="L")
    Xs = rs.rand(20, 2)
    ns = np.zeros((20, 2), dtype=int)
    dev.ll(th, 0.0)
    m, v, _ = dev.predict(Xs, ns, want_var=True)
    pm, ps, pc = orc.predict(orc.KERNEL_SE, th, X, n, ref["L"], ref["alpha"], Xs, ns)
    assert_close(m, pm, rtol=1e-9, atol=1e-9, what="mean")
    assert np.all(np.abs(v - np.diag(pc)) <= 1e-9 * th[0] ** 2)


@pytest.mark.parametrize("case", ["hypermp_matern_nu2p5_1d", "hypermp_matern_nu2p2_1d", "hypermp_matern_nu2p5_2d",
                                  "hypermp_gibbs_tanh"])
def test_hyper_derivatives_against_mpmath_derivatives_of_the_reference_functions(dev, case):
    """gpt_cov_pairs with hyper_deriv on the device against 40-digit mpmath derivatives of the reference's own
    covariance functions (the north-star tolerance 1e-9 with three orders of margin)."""
    from test_covfn_host import hypermp_check
    hypermp_check(lambda kid, params, Xi, Xj, ni, nj, hd: dev.cov_pairs(kid, params, Xi, Xj, ni, nj, hyper_deriv=hd), case)
