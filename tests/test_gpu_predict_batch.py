"""GPU tests of batched prediction (gpt_predict_batched): predictions at many hyper-parameter vectors in one launch of the
persistent many-theta kernel -- the per-sample unit of compute_from_MCMC / predict_MCMC (reference
gaussian_process.py:1944-1969, 2136-2254) -- against the single-theta path (which is pinned to the reference goldens)
and against the oracle."""
import warnings

import numpy as np
import pytest

import gptools_b200 as g
from helpers import assert_close, load_golden
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def _loop(gp, thetas, Xs, n):
    saved = np.array(gp.free_params[:], dtype=float)
    means, stds = [], []
    for th in thetas:
        gp.update_hyperparameters(th)
        m, s = gp.predict(Xs, n=n)
        means.append(m)
        stds.append(s)
    gp.update_hyperparameters(saved)
    return np.array(means), np.array(stds)


def _check(gp, thetas, Xs, n, prior_var, rtol_mean=1e-9):
    res = gp.predict_batch(thetas, Xs, n=n)
    assert res is not None
    mean, std, good = res
    assert good.all()
    m_l, s_l = _loop(gp, thetas, Xs, n)
    assert_close(mean, m_l, rtol=rtol_mean, atol=rtol_mean * np.abs(m_l).max(), what="mean")
    assert np.all(np.abs(std ** 2 - s_l ** 2) <= 1e-9 * prior_var), np.abs(std ** 2 - s_l ** 2).max()


def test_kat1_golden_through_the_batched_prediction():
    """The reference's own numbers (KAT-1: SE 2-D, value + both gradient observations) through the batch entry."""
    gd = load_golden("se2d_kat1")
    rs = np.random.RandomState(0)
    X = rs.rand(6, 2)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.7, 1.1], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k)
    gp.add_data(X, np.sin(X).sum(1), err_y=0.01)
    gp.add_data(X, np.cos(X[:, 0]), n=np.tile([1, 0], (6, 1)), err_y=0.01)
    gp.add_data(X, np.cos(X[:, 1]), n=np.tile([0, 1], (6, 1)), err_y=0.01)
    th = np.array([[1.3, 0.7, 1.1], [1.2, 0.8, 1.0]])
    mean, std, good = gp.predict_batch(th, gd["Xs"])
    assert good.all()
    assert_close(mean[0], gd["mean"], rtol=1e-9)
    assert_close(std[0], gd["std"], rtol=1e-6)
    m1, s1, _ = gp.predict_batch(th, gd["Xs"], n=np.tile([1, 0], (4, 1)))
    assert_close(m1[0], gd["mean_d1"], rtol=1e-9)
    assert_close(s1[0], gd["std_d1"], rtol=1e-6)


@pytest.mark.parametrize("M,Ms", [(512, 400), (333, 77), (64, 1), (700, 130)])
def test_se2d_with_derivative_observations(M, Ms):
    rs = np.random.RandomState(M)
    nv = M // 2
    nd = (M - nv) // 2
    Xv, Xa, Xb = rs.rand(nv, 2), rs.rand(nd, 2), rs.rand(M - nv - nd, 2)
    f = lambda x: np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k)
    gp.add_data(Xv, f(Xv) + 0.05 * rs.randn(nv), err_y=0.05)
    gp.add_data(Xa, 3 * np.cos(3 * Xa[:, 0]) * np.cos(2 * Xa[:, 1]), err_y=0.05, n=np.tile([1, 0], (len(Xa), 1)))
    gp.add_data(Xb, -2 * np.sin(3 * Xb[:, 0]) * np.sin(2 * Xb[:, 1]), err_y=0.05, n=np.tile([0, 1], (len(Xb), 1)))
    thetas = np.array([1.0, 0.3, 0.4]) * np.exp(0.1 * rs.randn(9, 3))
    Xs = rs.rand(Ms, 2)
    _check(gp, thetas, Xs, 0, prior_var=thetas[:, :1] ** 2)
    ns = np.zeros((Ms, 2), dtype=int)
    ns[::2, 0] = 1                                       # mixed value / derivative predictions
    _check(gp, thetas, Xs, ns, prior_var=(thetas[:, :1] / thetas[:, 1:2]) ** 2 + thetas[:, :1] ** 2)


@pytest.mark.parametrize("kernel", ["matern52", "matern_generic", "gibbs", "composite", "se3d"])
def test_other_kernels_and_noise(kernel):
    rs = np.random.RandomState(11)
    D = 3 if kernel == "se3d" else 1
    X = rs.rand(150, D) * (1.1 if kernel == "gibbs" else 3.0)
    y = np.sin(2 * X[:, 0]) + 0.05 * rs.randn(150)
    if kernel == "matern52":
        k = g.Matern52Kernel(initial_params=[1.0, 0.8], param_bounds=[(0, 10)] * 2)
    elif kernel == "matern_generic":
        k = g.MaternKernel(initial_params=[1.0, 2.2, 0.8], param_bounds=[(0, 10)] * 3)
    elif kernel == "gibbs":
        k = g.GibbsKernel1dTanh(initial_params=[1.5, 0.6, 0.1, 0.05, 0.9],
                                param_bounds=[(0, 10), (0, 5), (0, 5), (0, 1), (0, 2)])
    elif kernel == "composite":
        k = (g.SquaredExponentialKernel(initial_params=[1.0, 0.9], param_bounds=[(0, 10)] * 2) *
             g.Matern52Kernel(initial_params=[0.8, 0.5], param_bounds=[(0, 10)] * 2) +
             g.SquaredExponentialKernel(initial_params=[0.4, 0.15], param_bounds=[(0, 10)] * 2))
    else:
        k = g.SquaredExponentialKernel(num_dim=3, initial_params=[1.0, 0.8, 0.9, 1.1], param_bounds=[(0, 10)] * 4)
    nk = g.DiagonalNoiseKernel(D, initial_noise=0.1, fixed_noise=False, noise_bound=(0, 1))
    gp = g.GaussianProcess(k, noise_k=nk)
    gp.add_data(X, y, err_y=0.05)
    if D == 1:
        gp.add_data(X[::10], 2 * np.cos(2 * X[::10, 0]), err_y=0.2, n=1)
    th0 = np.array(gp.free_params[:], dtype=float)
    thetas = th0 * np.exp(0.03 * rs.randn(5, len(th0)))
    Xs = rs.rand(60, D) * (1.1 if kernel == "gibbs" else 3.0)
    tol = 1e-7 if kernel == "matern_generic" else 1e-9
    _check(gp, thetas, Xs, 0, prior_var=10.0, rtol_mean=tol)
    if D == 1:
        _check(gp, thetas, Xs, 1, prior_var=1e3, rtol_mean=tol)
    # noise=True adds sigma_n^2 on value predictions
    m0, s0, _ = gp.predict_batch(thetas, Xs, n=0, noise=False)
    m1, s1, _ = gp.predict_batch(thetas, Xs, n=0, noise=True)
    assert_close(s1 ** 2 - s0 ** 2, np.repeat(thetas[:, -1:] ** 2, 60, axis=1), rtol=1e-6, atol=1e-12)


def test_against_the_oracle_and_bad_rows():
    """M = 512 config-3 problem: batched prediction against the numpy oracle; a theta outside the prior support and a
    theta whose covariance is not positive definite come back as NaN rows with good == False."""
    import bench
    X, n, y, err = bench.c3_problem()
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n)
    th = bench.theta_batch(6)
    th[1, 0] = 1e170                                     # sigma_f^2 overflows: covariance not positive definite
    th[2, 1] = -th[2, 1]                                 # outside the uniform prior, yet a valid covariance: the reference's
    Xs = np.random.RandomState(2).rand(100, 2)           # per-sample wrapper predicts it (gaussian_process.py:2301-2330)
    mean, std, good = gp.predict_batch(th, Xs)
    assert list(good) == [True, False, True, True, True, True]
    assert np.isnan(mean[1]).all() and np.isnan(std[1]).all()
    t2 = th[2].copy()
    t2[1] = -t2[1]
    m2, s2, _ = gp.predict_batch(t2[None, :], Xs)
    assert_close(mean[2], m2[0], rtol=1e-12, atol=1e-13)  # k depends on l^2 only
    for b in (0, 5):
        r = orc.compute_K_L_alpha_ll(0, th[b], gp.X, gp.n, gp.y, gp.err_y)
        m, s, _ = orc.predict(0, th[b], gp.X, gp.n, r["L"], r["alpha"], Xs, np.zeros((100, 2), int))
        assert_close(mean[b], m, rtol=1e-8, atol=1e-9, what="mean vs oracle")
        assert np.all(np.abs(std[b] ** 2 - s ** 2) <= 1e-9 * th[b, 0] ** 2)
    # T or M > 2048: not served by the batch entry -> None (compute_from_MCMC then loops)
    gpT = g.GaussianProcess(g.SquaredExponentialKernel(initial_params=[1.0, 0.3], param_bounds=[(0, 10)] * 2))
    Xq = np.linspace(0, 1, 12)
    T = np.zeros((3, 12))
    T[0, :6] = T[1, 3:9] = T[2, 6:] = 1 / 6.0
    gpT.add_data(Xq, [2.5, 2.0, 1.0], err_y=0.05, T=T)
    assert gpT.predict_batch(np.array([[1.0, 0.3]]), np.array([0.5])) is None


def test_compute_from_MCMC_and_predict_MCMC_batched_vs_loop():
    rs = np.random.RandomState(3)
    X = np.sort(rs.rand(120)) * 4
    y = np.sin(2 * X) + 0.1 * rs.randn(120)
    k = g.SquaredExponentialKernel(initial_params=[1.0, 0.7], param_bounds=[(0.05, 5), (0.1, 3)])
    gp = g.GaussianProcess(k, X=X, y=y, err_y=0.1)
    trace = np.array([1.0, 0.7]) * np.exp(0.1 * rs.randn(40, 2))
    Xs = np.linspace(0, 4, 90)
    out_b = gp.predict_MCMC(Xs, flat_trace=trace)
    gp._mcmc_predict_by_loop = True
    out_l = gp.predict_MCMC(Xs, flat_trace=trace)
    gp._mcmc_predict_by_loop = False
    assert_close(out_b["mean"], out_l["mean"], rtol=1e-9, atol=1e-10)
    assert_close(out_b["std"], out_l["std"], rtol=1e-7, atol=1e-10)


def test_mean_function_with_free_parameters():
    """Per-theta mean-function parameters: y - mu_theta(X) travels as the per-theta right-hand side, mu_theta(X*) is added
    on the host -- same numbers as the per-sample path."""
    rs = np.random.RandomState(8)
    X = np.sort(rs.rand(100)) * 3
    y = 2.0 + np.sin(2 * X) + 0.05 * rs.randn(100)
    k = g.SquaredExponentialKernel(initial_params=[1.0, 0.7], param_bounds=[(0, 10)] * 2)
    mu = g.ConstantMeanFunction(initial_params=[1.8], param_bounds=[(-5, 5)])
    gp = g.GaussianProcess(k, X=X, y=y, err_y=0.05, mu=mu)
    th0 = np.array(gp.free_params[:], dtype=float)
    assert len(th0) == 3
    thetas = th0 + 0.05 * rs.randn(6, 3)
    Xs = np.linspace(0, 3, 50)
    _check(gp, thetas, Xs, 0, prior_var=thetas[:, :1] ** 2)
    _check(gp, thetas, Xs, 1, prior_var=(thetas[:, :1] / thetas[:, 1:2]) ** 2)


def test_randomised_batched_vs_single_theta_paths():
    """tools/fuzz_batched.py: random sizes (1 ... 700, tile edges), dimensions, derivative orders, kernels (SE, Matern-5/2,
    generic Matern, sums), noise: ll, gradient and batched prediction of the persistent kernel against the single-theta
    path."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_batched.py"), "30", "7"], cwd=root,
                         capture_output=True, text=True)
    assert out.returncode == 0 and "fuzz OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_c3_full_size_batch_prediction():
    """Config-3 size (4096 thetas, M = 512, 2-D SE with derivative observations): one launch predicts at 128 points for
    every theta; rows drawn at random agree with the single-theta path, every row is finite, and the batch entry
    reproduces the reference's KAT-5 values (ll of eight thetas of the batch, predictive mean / std at theta 0)."""
    import bench
    X, n, y, err = bench.c3_problem()
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n)
    th = bench.theta_batch(4096)
    Xs = np.random.RandomState(2).rand(128, 2)
    mean, std, good = gp.predict_batch(th, Xs)
    assert good.all() and np.isfinite(mean).all() and np.isfinite(std).all() and (std > 0).all()
    for b in (0, 1234, 4095):
        gp.update_hyperparameters(th[b])
        m, s = gp.predict(Xs)
        assert_close(mean[b], m, rtol=1e-9, atol=1e-10, what="mean of theta %d" % b)
        assert np.all(np.abs(std[b] ** 2 - s ** 2) <= 1e-9 * th[b, 0] ** 2)
    # the reference's own numbers (KAT-5): ll of eight thetas of the batch, predictions at theta 0
    gd = load_golden("c3_kat5")
    assert_close(th[gd["theta_idx"]], gd["theta"], rtol=1e-8)
    dev, _ = gp._sync_device()
    full = np.hstack([th[gd["theta_idx"]], np.zeros((8, 1))])
    m8, v8, ll, st = dev.predict_batched(full, gd["Xs"], np.zeros(gd["Xs"].shape, dtype=int))
    assert (st == 0).all()
    assert_close(ll + float(gd["log_prior"]), gd["ll"], rtol=1e-9, what="ll next to the predictions")
    assert_close(m8[0], gd["mean"], rtol=1e-8, what="mean at theta 0 vs the reference")
    assert_close(np.sqrt(v8[0]), gd["std"], rtol=1e-5, what="std at theta 0 vs the reference")


def test_predict_MCMC_matches_the_reference_on_the_device():
    """Golden mcmc_predict_se1d from the reference's predict_MCMC / compute_from_MCMC, through gpt_predict_batched."""
    gd = load_golden("mcmc_predict_se1d")
    k = g.SquaredExponentialKernel(initial_params=[1.0, 0.7], param_bounds=[(0.05, 5), (0.1, 3)])
    gp = g.GaussianProcess(k)
    nv = int((gd["n"][:, 0] == 0).sum())
    gp.add_data(gd["X"][:nv, 0], gd["y"][:nv], err_y=gd["err_y"][:nv])
    gp.add_data(gd["X"][nv:, 0], gd["y"][nv:], err_y=gd["err_y"][nv:], n=1)
    res = gp.compute_from_MCMC(gd["Xs"], flat_trace=gd["trace"])
    assert_close(np.array(res["mean"]), gd["means"], rtol=1e-8, atol=1e-10, what="per-sample means")
    assert_close(np.array(res["std"]), gd["stds"], rtol=1e-6, atol=1e-9, what="per-sample stds")
    out = gp.predict_MCMC(gd["Xs"], flat_trace=gd["trace"])
    assert_close(out["mean"], gd["mean"], rtol=1e-8, atol=1e-10)
    assert_close(out["std"], gd["std"], rtol=1e-6, atol=1e-9)
    out1 = gp.predict_MCMC(gd["Xs"], n=1, flat_trace=gd["trace"])
    assert_close(out1["mean"], gd["mean_d1"], rtol=1e-8, atol=1e-9)
    assert_close(out1["std"], gd["std_d1"], rtol=1e-6, atol=1e-9)
    thin = gp.predict_MCMC(gd["Xs"], flat_trace=gd["trace"], burn=2, thin=3)
    assert_close(thin["mean"], gd["mean_thin"], rtol=1e-8, atol=1e-10)
    assert_close(thin["std"], gd["std_thin"], rtol=1e-6, atol=1e-9)
