"""CPU tests of the host side: the C-ABI boundary loads and exports every declared symbol, the product fails
loudly without a GPU, and the GaussianProcess / Kernel / prior bookkeeping reproduces the reference's
(bit-exact for index / derivative-order / row-order work).  The numerics are supplied by a fake device
backed by the pinned oracle (tests/fake_device.py), so what is tested here is the HOST logic only."""
import os
import pickle
import re
import warnings

import numpy as np
import pytest

import gptools_b200 as g
from fake_device import FakeDevice
from helpers import assert_close, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def with_fake(gp):
    gp._dev_obj = FakeDevice()
    return gp


# ------------------------------------------------------------------ boundary
def test_library_exports_every_declared_symbol():
    from gptools_b200 import _lib
    lib = _lib.load_library()
    hdr = open(os.path.join(ROOT, "include", "gptb200.h")).read()
    declared = set(re.findall(r"\b(gpt_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gpt_version() >= 100


def test_product_fails_loudly_without_gpu():
    from gptools_b200 import _lib
    if _lib.load_library().gpt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.GPTLibraryError):
        _lib.Device()
    k = g.SquaredExponentialKernel(num_dim=1, initial_params=[1.0, 1.0], param_bounds=[(0, 10)] * 2)
    gp = g.GaussianProcess(k, X=[0.0, 1.0], y=[0.0, 1.0])
    with pytest.raises(_lib.GPTLibraryError):
        gp.compute_K_L_alpha_ll()
    with pytest.raises(_lib.GPTLibraryError):
        k(np.zeros((1, 1)), np.zeros((1, 1)), np.zeros((1, 1), int), np.zeros((1, 1), int))


def test_product_never_imports_oracle():
    import ast
    pkg = os.path.join(ROOT, "gptools_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(dirpath, f)).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.module:
                        names = [node.module]
                    assert not any(n.split(".")[0] in ("oracle", "fake_device") for n in names), (f, names)


# ------------------------------------------------------------------ add_data bookkeeping (bit-exact)
def _kat1_gp(**kw):
    rs = np.random.RandomState(0)
    X = rs.rand(6, 2)
    y = np.sin(X).sum(1)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.7, 1.1], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, **kw)
    gp.add_data(X, y, err_y=0.01)
    gp.add_data(X, np.cos(X[:, 0]), n=np.tile([1, 0], (6, 1)), err_y=0.01)
    gp.add_data(X, np.cos(X[:, 1]), n=np.tile([0, 1], (6, 1)), err_y=0.01)
    return gp


def test_add_data_matches_reference_rows_bit_exact():
    gd = load_golden("se2d_kat1")
    gp = _kat1_gp()
    for name in ("X", "n", "y", "err_y"):
        assert np.array_equal(getattr(gp, name), gd[name]), name
    assert gp.n.dtype.kind == "i" and gp.T is None


def test_add_data_with_T_and_scalar_point_bit_exact():
    gd = load_golden("gibbs_kat3")
    k = g.GibbsKernel1dTanh(initial_params=[1.5, 0.6, 0.1, 0.05, 0.9],
                            param_bounds=[(0, 10), (0, 5), (0, 5), (0, 1), (0, 2)])
    Xq = np.linspace(0, 1.1, 12)
    T = np.zeros((3, 12))
    T[0, :6] = T[1, 3:9] = T[2, 6:] = 1 / 6.0
    gp = g.GaussianProcess(k)
    gp.add_data(Xq, [2.5, 2.0, 1.0], err_y=0.05, T=T)
    gp.add_data(0, 0, n=1)      # identity back-fill + block_diag (gaussian_process.py:471-491)
    assert gp.T.shape == (4, 13)
    for name in ("X", "n", "y", "err_y", "T"):
        assert np.array_equal(getattr(gp, name), gd[name]), name


def test_add_data_rejects_what_the_reference_rejects():
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1, 1, 1], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k)
    with pytest.raises(ValueError):
        gp.add_data(np.zeros((3, 1)), np.zeros(3))
    with pytest.raises(ValueError):
        gp.add_data(np.zeros((3, 2)), np.zeros(3), err_y=-1.0)
    with pytest.raises(ValueError):
        gp.add_data(np.zeros((3, 2)), np.zeros(3), n=-1)
    with pytest.raises(ValueError):
        gp.add_data(np.zeros((3, 2)), np.zeros(3), err_y=np.zeros(2))
    with pytest.raises(ValueError):
        gp.add_data(np.zeros((3, 2)), np.zeros(2), T=np.zeros((2, 4)))
    with pytest.raises(TypeError):
        g.GaussianProcess("not a kernel")
    with pytest.raises(g.GPArgumentError):
        g.GaussianProcess(k, X=np.zeros((1, 2)))


def test_condense_duplicates_keeps_the_likelihood():
    rs = np.random.RandomState(3)
    Xu = rs.rand(5, 1)
    X = np.vstack([Xu, Xu[:2]])
    y = np.sin(3 * X[:, 0]) + 0.01 * rs.randn(7)
    k = g.SquaredExponentialKernel(num_dim=1, initial_params=[1.0, 0.5], param_bounds=[(0, 10)] * 2)
    gp1 = with_fake(g.GaussianProcess(k, X=X, y=y, err_y=0.1))
    gp1.compute_K_L_alpha_ll()
    gp2 = with_fake(g.GaussianProcess(k, X=X, y=y, err_y=0.1))
    gp2.condense_duplicates()
    assert gp2.X.shape == (5, 1) and gp2.T.shape == (7, 5)
    assert np.array_equal(gp2.T.sum(axis=1), np.ones(7))
    gp2.compute_K_L_alpha_ll()
    assert_close(gp2.ll, gp1.ll, rtol=1e-10)


# ------------------------------------------------------------------ ll / gradient / priors through the host layer
def test_ll_and_gradient_kat1_through_GaussianProcess():
    gd = load_golden("se2d_kat1")
    gp = with_fake(_kat1_gp(use_hyper_deriv=True))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gp.compute_K_L_alpha_ll()
    assert_close(gp.ll, gd["ll"], rtol=1e-12)
    assert_close(gp.ll_deriv, gd["ll_deriv"], rtol=1e-10)
    assert gp.alpha.shape == (18, 1)
    assert_close(gp.alpha.ravel(), gd["alpha"], rtol=1e-9)
    assert_close(gp.K, gd["K"], rtol=1e-12, atol=1e-14)
    assert np.array_equal(gp.noise_K, np.zeros((18, 18)))
    # update_hyperparameters returns (-ll, -grad) and restores use_hyper_deriv (gaussian_process.py:1411-1416)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f, df = gp.update_hyperparameters(np.array([1.3, 0.7, 1.1]))
        assert_close(f, -gd["ll"], rtol=1e-12)
        assert_close(df, -gd["ll_deriv"], rtol=1e-10)
        assert np.isscalar(gp.update_hyperparameters(np.array([1.3, 0.7, 1.1]), hyper_deriv_handling='value'))
        # outside the prior support -> (inf, zeros) without touching the device
        ncalls = len(gp._dev_obj.calls)
        f, df = gp.update_hyperparameters(np.array([11.0, 0.7, 1.1]))
        assert np.isinf(f) and np.all(df == 0) and len(gp._dev_obj.calls) == ncalls


def test_demo_prior_and_fixed_parameter_quirk():
    """ll includes the log-prior of ALL parameters, fixed ones too: the ZeroKernel's fixed sigma_n = 0 with bounds
    (0, 1e16) contributes -log(1e16)."""
    gd = load_golden("demo_c1_kat4")
    hp = g.UniformJointPrior([(0, 20)]) * g.GammaJointPriorAlt([1.0], [0.7])
    k = g.SquaredExponentialKernel(initial_params=gd["params"], hyperprior=hp)
    gp = with_fake(g.GaussianProcess(k, use_hyper_deriv=True))
    gp.add_data(gd["X"][:-1], gd["y"][:-1], err_y=gd["err_y"][:-1])
    gp.add_data(0, 0, n=1)
    assert np.array_equal(gp.X, gd["X"]) and np.array_equal(gp.n, gd["n"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gp.compute_K_L_alpha_ll()
    assert_close(gp.hyperprior(gp.params), gd["log_prior"], rtol=1e-13)
    assert_close(gp.ll, gd["ll"], rtol=1e-12)
    assert_close(gp.ll_deriv, gd["ll_deriv"], rtol=1e-7, atol=1e-9)


def test_diagonal_noise_kernel_layout_and_gradient():
    gd = load_golden("se_diagnoise")
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=gd["params"], param_bounds=[(0, 10)] * 3)
    nk = g.DiagonalNoiseKernel(num_dim=2, initial_noise=float(gd["noise_sigma"][0]), noise_bound=(0, 5))
    gp = with_fake(g.GaussianProcess(k, noise_k=nk, use_hyper_deriv=True))
    gp.add_data(gd["X"][:20], gd["y"][:20], err_y=0.03)
    gp.add_data(gd["X"][20:], gd["y"][20:], n=np.tile([0, 1], (5, 1)), err_y=0.2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gp.compute_K_L_alpha_ll()
    assert_close(gp.ll, gd["ll"], rtol=1e-12)
    assert_close(gp.ll_deriv, gd["ll_deriv"], rtol=1e-9)
    assert len(gp.free_params) == 4 and list(gp.free_param_names[:])[-1] == r'\sigma_n'


def test_mean_function_enters_as_rhs_and_gradient_slot():
    rs = np.random.RandomState(2)
    X = rs.rand(12, 1)
    y = 2.0 + np.sin(4 * X[:, 0])
    k = g.SquaredExponentialKernel(num_dim=1, initial_params=[1.0, 0.3], param_bounds=[(0, 10)] * 2)
    mu = g.ConstantMeanFunction(initial_params=[1.5])
    gp = with_fake(g.GaussianProcess(k, mu=mu, X=X, y=y, err_y=0.05, use_hyper_deriv=True))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f0, g0 = gp.update_hyperparameters(np.array([1.0, 0.3, 1.5]))
        eps = 1e-6
        fp, _ = gp.update_hyperparameters(np.array([1.0, 0.3, 1.5 + eps]))
        fm, _ = gp.update_hyperparameters(np.array([1.0, 0.3, 1.5 - eps]))
    assert_close(g0[2], (fp - fm) / (2 * eps), rtol=1e-6)
    assert "set_y" in gp._dev_obj.calls      # only the right-hand side was re-sent, not the whole data set


# ------------------------------------------------------------------ batched entry (host logic)
def test_update_hyperparameters_batch_matches_scalar_calls():
    gp = with_fake(_kat1_gp(use_hyper_deriv=True))
    rs = np.random.RandomState(1)
    th = np.array([1.3, 0.7, 1.1]) * np.exp(0.1 * rs.randn(6, 3))
    th[3, 0] = 12.0       # outside the prior support
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f, df = gp.update_hyperparameters_batch(th)
        before = np.array(gp.free_params[:])
        for b in range(6):
            fb, dfb = gp.update_hyperparameters(th[b])
            if b == 3:
                assert np.isinf(f[b]) and np.all(df[b] == 0)
            else:
                assert_close(f[b], fb, rtol=1e-12)
                assert_close(df[b], dfb, rtol=1e-10)
    assert np.array_equal(before, np.array([1.3, 0.7, 1.1]))       # batch call leaves the GP's state alone
    assert gp.update_hyperparameters_batch(th, with_deriv=False).shape == (6,)


def test_compute_ll_matrix_shape_and_values():
    gp = with_fake(_kat1_gp())
    ll_vals, pv = gp.compute_ll_matrix([(1.0, 1.5), (0.5, 0.9), (0.9, 1.3)], [2, 3, 2])
    assert ll_vals.shape == (2, 3, 2)
    v = gp.update_hyperparameters(np.array([pv[0][1], pv[1][2], pv[2][0]]))
    assert_close(ll_vals[1, 2, 0], -v, rtol=1e-12)
    # the reference's own grid (gaussian_process.py:1607-1692): same values, same parameter axes
    gd = load_golden("ll_matrix_kat1")
    assert_close(ll_vals, gd["ll_vals"], rtol=1e-9, what="ll grid")
    for a, b in zip(pv, (gd["p0"], gd["p1"], gd["p2"])):
        assert_close(a, b, rtol=0, atol=0)


# ------------------------------------------------------------------ predict / draw_sample host logic
def test_predict_return_conventions_and_output_transform():
    gd = load_golden("se2d_kat1")
    gp = with_fake(_kat1_gp())
    Xs = gd["Xs"]
    mean, std = gp.predict(Xs)
    assert_close(mean, gd["mean"], rtol=1e-9)
    assert_close(std, gd["std"], rtol=1e-6)
    mean2, cov = gp.predict(Xs, return_cov=True)
    assert cov.shape == (4, 4)
    assert gp.predict(Xs, return_std=False).shape == (4,)
    out = gp.predict(Xs, full_output=True)
    assert set(out) == {"mean", "std", "cov"}
    m1, s1 = gp.predict(Xs, n=np.tile([1, 0], (4, 1)))
    assert_close(m1, gd["mean_d1"], rtol=1e-9)
    W = np.array([[0.25, 0.25, 0.25, 0.25], [1.0, -1.0, 0.0, 0.0]])
    mt, ct = gp.predict(Xs, return_cov=True, output_transform=W)
    assert_close(mt, W.dot(gd["mean"]), rtol=1e-9)
    assert_close(ct, W.dot(gd["cov"]).dot(W.T), rtol=1e-6, atol=1e-12)
    with pytest.raises(ValueError):
        gp.predict(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        gp.predict(Xs, n=-1)
    # full_covar is accepted as an alias of return_cov
    assert gp.predict(Xs, full_covar=True)[1].shape == (4, 4)


def test_draw_sample_cholesky_with_rand_vars():
    gd = load_golden("gibbs_kat3")
    k = g.GibbsKernel1dTanh(initial_params=gd["params"], param_bounds=[(0, 10), (0, 5), (0, 5), (0, 1), (0, 2)])
    gp = with_fake(g.GaussianProcess(k))
    gp.add_data(gd["X"][:12, 0], gd["y"][:3], err_y=0.05, T=gd["T"][:3, :12])
    gp.add_data(0, 0, n=1)
    gp.compute_K_L_alpha_ll()
    assert_close(gp.ll, gd["ll"], rtol=1e-12)
    samp = gp.draw_sample(gd["Xs"], rand_vars=gd["rand_vars"], method="cholesky")
    assert_close(samp, gd["draw"], rtol=1e-7, atol=1e-7)
    with pytest.raises(ValueError):
        gp.draw_sample(gd["Xs"], rand_vars=gd["rand_vars"], method="nope")


def test_unsupported_orders_raise_before_any_device_call():
    k = g.Matern52Kernel(num_dim=1, initial_params=[1.0, 0.5], param_bounds=[(0, 10)] * 2)
    gp = with_fake(g.GaussianProcess(k, X=np.linspace(0, 1, 5), y=np.zeros(5), err_y=0.1))
    gp.compute_K_L_alpha_ll()
    n0 = len(gp._dev_obj.calls)
    with pytest.raises(ValueError):
        gp.predict(np.array([0.5]), n=2)
    assert len(gp._dev_obj.calls) == n0
    gk = g.GibbsKernel1dTanh(initial_params=[1, 1, 1, 1, 1], param_bounds=[(0, 10)] * 5)
    with pytest.raises(NotImplementedError):
        gk._check_orders(np.array([[2]]), np.array([[0]]))
    mk = g.MaternKernel(num_dim=1, initial_params=[1.0, 2.0, 0.5], param_bounds=[(0, 10)] * 3)
    assert mk.device_descriptor()[0] == 2    # any nu > 0 runs on the device (round 2: K_nu of real order)
    mk = g.MaternKernel(num_dim=1, initial_params=[1.0, 0.0, 0.5], param_bounds=[(0, 10)] * 3)
    with pytest.raises(ValueError):
        mk.device_descriptor()
    mk = g.MaternKernel(num_dim=1, initial_params=[1.0, 2.5, 0.5], param_bounds=[(0, 10)] * 3)
    # d/dnu has no closed form on the device: finite differences of the device evaluation (GPU suite); with nu free and
    # gradients requested the batched entry is not used
    assert mk.fd_hyper_idxs == (1,) and not mk.batchable(True) and mk.batchable(False)
    mk.check_hyper_deriv([0, 1, 2])
    mkf = g.MaternKernel(num_dim=1, initial_params=[1.0, 2.5, 0.5], fixed_params=[False, True, False],
                         param_bounds=[(0, 10)] * 3)
    assert mkf.batchable(True)


def test_pickle_drops_the_device_handle():
    gp = with_fake(_kat1_gp())
    gp.compute_K_L_alpha_ll()
    gp2 = pickle.loads(pickle.dumps(gp))
    assert gp2._dev_obj is None and not gp2.K_up_to_date
    assert np.array_equal(gp2.X, gp.X) and np.array_equal(gp2.k.params, gp.k.params)


# ------------------------------------------------------------------ priors
def test_priors_logpdf_and_derivatives():
    hp = g.UniformJointPrior([(0, 20)]) * g.GammaJointPriorAlt([1.0], [0.7]) * g.NormalJointPrior([0.5], [2.0]) * \
        g.LogNormalJointPrior([0.1], [0.4])
    th = np.array([1.9, 0.98, 0.3, 1.7])
    import scipy.stats as st
    b = (1.0 + np.sqrt(1.0 + 4 * 0.49)) / (2 * 0.49)
    want = (-np.log(20.0) + st.gamma.logpdf(0.98, 1.0 + b, scale=1.0 / b) + st.norm.logpdf(0.3, 0.5, 2.0) +
            st.lognorm.logpdf(1.7, 0.4, scale=np.exp(0.1)))
    assert_close(hp(th), want, rtol=1e-13)
    for i in range(4):
        e = np.zeros(4)
        e[i] = 1e-6
        fd = (hp(th + e) - hp(th - e)) / 2e-6
        assert_close(hp(th, hyper_deriv=i), fd, rtol=1e-6, atol=1e-8)
    assert np.isinf(hp(np.array([21.0, 0.98, 0.3, 1.7])))
    draw = hp.random_draw(size=5)
    assert draw.shape == (4, 5)
    assert len(hp.bounds) == 4
    u = hp.elementwise_cdf(th)
    assert_close(hp.sample_u(u), th, rtol=1e-9)
    su = g.SortedUniformJointPrior(3, 0.0, 2.0)
    assert_close(su(np.array([0.1, 0.5, 1.5])), np.log(6.0) - 3 * np.log(2.0), rtol=1e-13)


@pytest.mark.refonly
def test_priors_and_bookkeeping_against_the_reference_itself():
    from oracle.ref_shim import load_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = load_reference()
        for mod in (r, g):
            pass
        th = np.array([1.9, 0.98])
        hp_r = r.UniformJointPrior([(0, 20)]) * r.GammaJointPriorAlt([1.0], [0.7])
        hp_g = g.UniformJointPrior([(0, 20)]) * g.GammaJointPriorAlt([1.0], [0.7])
        assert hp_r(th) == hp_g(th)
        assert hp_r(th, hyper_deriv=1) == hp_g(th, hyper_deriv=1)
        # add_data: same calls, same rows, bit for bit
        rs = np.random.RandomState(4)
        X = rs.rand(7, 2)
        outs = []
        for mod in (r, g):
            k = mod.SquaredExponentialKernel(num_dim=2, initial_params=[1, 1, 1], param_bounds=[(0, 10)] * 3)
            gp = mod.GaussianProcess(k)
            gp.add_data(X, X[:, 0], err_y=0.1)
            gp.add_data(X[:3], X[:3, 1], n=[[1, 0]] * 3)
            gp.add_data(X[:2], [1.0], T=[[0.5, 0.5]], err_y=[0.2])
            outs.append((gp.X, gp.n, gp.y, gp.err_y, gp.T))
        for a, b in zip(*outs):
            assert np.array_equal(a, b)


# ------------------------------------------------------------------ sampler
def test_ensemble_sampler_recovers_a_gaussian():
    from gptools_b200.sampler import EnsembleSampler
    mu, sig = np.array([1.0, -2.0]), np.array([0.5, 2.0])
    s = EnsembleSampler(40, 2, lambda th: -0.5 * (((th - mu) / sig) ** 2).sum(axis=1),
                        random_state=np.random.RandomState(0))
    p0 = mu + 0.1 * np.random.RandomState(1).randn(40, 2)
    s.run_mcmc(p0, 600)
    assert s.chain.shape == (40, 600, 2) and s.lnprobability.shape == (40, 600)
    flat = s.chain[:, 200:, :].reshape(-1, 2)
    assert np.all(np.abs(flat.mean(0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(0) / sig - 1) < 0.15)
    assert 0.2 < s.acceptance_fraction.mean() < 0.95
    s.run_mcmc(s.chain[:, -1, :], 10)       # chains can be continued
    assert s.chain.shape[1] == 610


def test_sample_hyperparameter_posterior_uses_batched_calls():
    gp = with_fake(_kat1_gp())
    np.random.seed(0)
    s = gp.sample_hyperparameter_posterior(nwalkers=8, nsamp=3)
    assert s.chain.shape == (8, 3, 3)
    assert gp._dev_obj.calls.count("ll_batched") == 1 + 2 * 3     # initial ensemble + two half-moves per step
    assert np.isfinite(s.lnprobability).any()


class _OracleSE(g.Kernel):
    """numpy stand-in for the device SE kernel (the pinned oracle's pair function), so that the host-side kernel
    algebra can be checked without a GPU."""

    def __init__(self, num_dim, params):
        super(_OracleSE, self).__init__(num_dim=num_dim, num_params=num_dim + 1, initial_params=params,
                                        param_bounds=[(0, 10)] * (num_dim + 1))

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        from oracle import gp_oracle as orc
        return orc.se_pairs(np.atleast_2d(Xi), np.atleast_2d(Xj), np.atleast_2d(ni), np.atleast_2d(nj),
                            np.asarray(self.params, float), hyper_deriv=hyper_deriv)


def test_sum_and_product_kernels_match_reference():
    """k1 + k2 and k1 * k2 (general Leibniz rule over derivative orders up to (2, 1) per dimension) against the
    reference's SumKernel / ProductKernel on seeded pair lists (kernel/core.py:549-670)."""
    gd = load_golden("kernel_algebra_se2d")
    k1, k2 = _OracleSE(2, gd["params1"]), _OracleSE(2, gd["params2"])
    a = (gd["Xi"], gd["Xj"], gd["ni"], gd["nj"])
    assert_close((k1 * k2)(*a), gd["prod"], rtol=1e-10, atol=1e-12 * np.abs(gd["prod"]).max(), what="product")
    assert_close((k1 + k2)(*a), gd["sum"], rtol=1e-12, atol=1e-14 * np.abs(gd["sum"]).max(), what="sum")
    ks = k1 + k2
    assert_close(ks(*a, hyper_deriv=1), gd["sum_hd1"], rtol=1e-9, atol=1e-12 * np.abs(gd["sum_hd1"]).max())
    assert_close(ks(*a, hyper_deriv=4), gd["sum_hd4"], rtol=1e-9, atol=1e-12 * np.abs(gd["sum_hd4"]).max())
    kp = k1 * k2
    assert kp.num_params == 6 and list(kp.params) == list(gd["params1"]) + list(gd["params2"])
    with pytest.raises(NotImplementedError):
        kp(*a, hyper_deriv=0)


def test_warped_kernel_matches_reference():
    """LinearWarpedKernel(BetaWarpedKernel(SE)) -- nested input warping with first-derivative observations -- against
    the reference's K (kernel/warping.py:464-720).  The SE operand is the oracle stand-in, so this runs without a
    GPU; the GPU suite runs the same golden end to end."""
    gd = load_golden("warped_beta_linear_se")
    kse = _OracleSE(1, gd["params"][:2])
    kb = g.BetaWarpedKernel(kse, initial_params=gd["params"][2:4], param_bounds=[(0.01, 10)] * 2)
    k = g.LinearWarpedKernel(kb, [gd["params"][4]], [gd["params"][5]])
    assert_close(np.asarray(k.params, float), gd["params"], rtol=0, atol=0)
    assert list(np.asarray(k.fixed_params, bool)) == list(gd["fixed"])
    assert k.num_free_params == 4
    X, n = gd["X"], gd["n"]
    M = len(X)
    K = k(np.repeat(X, M, axis=0), np.tile(X, (M, 1)), np.repeat(n, M, axis=0), np.tile(n, (M, 1))).reshape(M, M)
    assert_close(K, gd["K"], rtol=1e-10, atol=1e-12 * np.abs(gd["K"]).max(), what="warped K")
    with pytest.raises(ValueError):
        k(X[:1], X[:1], np.array([[2]]), np.array([[0]]))


# ------------------------------------------------------------------ round-2 host fixes
def test_binary_kernel_parameters_are_write_through():
    """kernel/core.py:466-548: the combined kernel exposes write-through views with setters (gp.params = ...,
    gp.free_params[i] = v, restoring parameters after a per-theta loop)."""
    k1 = g.SquaredExponentialKernel(initial_params=[0.9, 0.7], param_bounds=[(0, 10)] * 2)
    k2 = g.SquaredExponentialKernel(initial_params=[0.5, 0.3], fixed_params=[False, True], param_bounds=[(0, 10)] * 2)
    ks = k1 + k2
    assert list(ks.params) == [0.9, 0.7, 0.5, 0.3] and ks.num_free_params == 3
    assert list(ks.free_param_idxs) == [0, 1, 2]
    ks.params = [1.0, 2.0, 3.0, 4.0]
    assert list(k1.params) == [1.0, 2.0] and list(k2.params) == [3.0, 4.0]
    ks.free_params = [1.5, 2.5, 3.5]
    assert list(k1.params) == [1.5, 2.5] and list(k2.params) == [3.5, 4.0]
    ks.free_params[2] = 0.25
    assert k2.params[0] == 0.25
    ks.params[1] = 7.0
    assert k1.params[1] == 7.0
    gp = g.GaussianProcess(ks, X=[0.0, 1.0, 2.0], y=[0.0, 1.0, 0.5], err_y=0.1)
    gp.params = [1.0, 1.0, 1.0, 4.0, 0.0]          # kernel (4) + ZeroKernel (1)
    assert list(ks.params) == [1.0, 1.0, 1.0, 4.0]
    gp.free_params = [0.2, 0.3, 0.4]
    assert list(k1.params) == [0.2, 0.3] and k2.params[0] == 0.4
    gp.free_params[0] = 0.9
    assert k1.params[0] == 0.9
    saved = np.array(gp.free_params[:], dtype=float)
    gp._set_free_params([5.0, 6.0, 7.0])
    gp._set_free_params(saved)
    assert list(gp.free_params[:]) == list(saved)


def test_batched_entry_falls_back_where_the_device_kernel_cannot_go():
    """Free Matern parameters beyond the persistent kernel's gradient slots take the per-theta loop instead of raising;
    Matern rows with an invalid order (nu <= 0) evaluate to inf; T and M > 2048 stay on the batched C entry."""
    rs = np.random.RandomState(0)
    X = np.sort(rs.rand(40)) * 3
    k = g.MaternKernel(num_dim=1, initial_params=[1.0, 2.5, 0.8], param_bounds=[(0, 10), (-5, 10), (0, 10)])
    gp = with_fake(g.GaussianProcess(k, X=X, y=np.sin(X), err_y=0.05))
    assert gp._batchable(False)
    th = np.array([[1.0, 2.5, 0.8], [1.1, -1.0, 0.7], [0.9, 1.5, 0.9]])
    f = gp.update_hyperparameters_batch(th, with_deriv=False)
    assert np.isfinite(f[0]) and np.isinf(f[1]) and np.isfinite(f[2])
    assert "ll_batched" in gp._dev_obj.calls
    k6 = g.MaternKernel(num_dim=6, initial_params=[1.0, 2.5] + [1.0] * 6, fixed_params=[False, True] + [False] * 6,
                        param_bounds=[(0, 10)] * 8)
    assert k6.batchable(False) and not k6.batchable(True)    # l_6 is parameter 7: outside the gradient slots
    X6 = rs.rand(12, 6)
    gp6 = with_fake(g.GaussianProcess(k6, X=X6, y=np.sin(X6).sum(1), err_y=0.05, use_hyper_deriv=True))
    assert not gp6._batchable(True) and gp6._batchable(False)
    gp6.BATCHED_KERNEL_MAX_M = 4                           # beyond the persistent kernel: the call runs the thetas serially
    assert gp6._batchable(True)
    # transformed observations stay on the batched entry (gpt_ll_batched runs them back to back on the device)
    kt = g.SquaredExponentialKernel(initial_params=[1.0, 0.5], param_bounds=[(0, 10)] * 2)
    gpt_ = with_fake(g.GaussianProcess(kt))
    gpt_.add_data(np.linspace(0, 1, 6), [0.5, 0.7], err_y=0.1, T=np.array([[0.5, 0.5, 0, 0, 0, 0], [0, 0, 0, 0.25, 0.25, 0.5]]))
    gpt_._dev_obj.calls = []
    thT = np.array([[1.0, 0.5], [1.2, 0.4]])
    fT = gpt_.update_hyperparameters_batch(thT, with_deriv=False)
    assert gpt_._dev_obj.calls.count("ll_batched") == 1 and "ll" not in gpt_._dev_obj.calls
    assert_close(fT[1], gpt_.update_hyperparameters(thT[1]), rtol=1e-12)


def test_condense_duplicates_prunes_unused_quadrature_points_and_remove_outliers_with_T():
    """gaussian_process.py:535-541 (all-zero T columns are dropped) and :583-621 (remove_outliers with T returns T_bad);
    compared with the reference itself where it is available."""
    rs = np.random.RandomState(5)
    Xq = np.linspace(0, 1, 9)
    T = np.zeros((3, 9))
    T[0, :4] = 0.25
    T[1, 2:6] = 0.25
    T[2, 3:7] = 0.25                                       # columns 7, 8 never enter
    y = np.array([1.0, 1.2, 5.0])
    k = g.SquaredExponentialKernel(initial_params=[1.0, 0.5], param_bounds=[(0, 10)] * 2)
    gp = with_fake(g.GaussianProcess(k))
    gp.add_data(Xq, y, err_y=0.1, T=T)
    gp.condense_duplicates()
    assert gp.T.shape == (3, 7) and gp.X.shape == (7, 1) and gp.n.shape == (7, 1)
    assert np.array_equal(gp.X.ravel(), Xq[:7])
    from oracle.ref_shim import load_reference, reference_available
    if reference_available():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = load_reference()
            kr = r.SquaredExponentialKernel(initial_params=[1.0, 0.5], param_bounds=[(0, 10)] * 2)
            gr = r.GaussianProcess(kr)
            gr.add_data(Xq, y, err_y=0.1, T=T)
            gr.condense_duplicates()
            assert np.array_equal(gr.T, gp.T) and np.array_equal(gr.X, gp.X) and np.array_equal(gr.n, gp.n)
            out_r = gr.remove_outliers(thresh=3)
            with_fake(gp)
            out_g = gp.remove_outliers(thresh=3)
            assert len(out_r) == len(out_g) == 6
            for a, b in zip(out_r, out_g):
                assert np.array_equal(np.asarray(a), np.asarray(b))
            for a, b in ((gr.X, gp.X), (gr.n, gp.n), (gr.y, gp.y), (gr.err_y, gp.err_y), (gr.T, gp.T)):
                assert np.array_equal(a, b)
    else:
        out = gp.remove_outliers(thresh=3)
        assert len(out) == 6 and out[4].shape == (3,)


def test_predict_noise_diagonal_without_pair_lists():
    """predict(noise=True, return_std=True) adds sigma_n^2 [n* == n_noise] per test point (kernel/noise.py:103-110)."""
    rs = np.random.RandomState(1)
    X = np.sort(rs.rand(12)) * 2
    k = g.SquaredExponentialKernel(initial_params=[1.0, 0.5], param_bounds=[(0, 10)] * 2)
    nk = g.DiagonalNoiseKernel(1, initial_noise=0.3, noise_bound=(0, 5))
    gp = with_fake(g.GaussianProcess(k, noise_k=nk, X=X, y=np.sin(X), err_y=0.05))
    Xs = np.linspace(0, 2, 7)
    m0, s0 = gp.predict(Xs, noise=False)
    m1, s1 = gp.predict(Xs, noise=True)
    assert_close(s1 ** 2, s0 ** 2 + 0.09, rtol=1e-12)
    _, s1d = gp.predict(Xs, n=1, noise=True)               # the noise sits on order 0 only
    _, s0d = gp.predict(Xs, n=1, noise=False)
    assert_close(s1d, s0d, rtol=1e-12)
    _, c1 = gp.predict(Xs, noise=True, return_cov=True)
    _, c0 = gp.predict(Xs, noise=False, return_cov=True)
    assert_close(c1, c0 + 0.09 * np.eye(7), rtol=1e-12, atol=1e-15)


def test_lockstep_multistart_equals_sequential_starts():
    """optimize_hyperparameters runs its random starts in lock-step on the batched entry; every start must end where
    its own sequential scipy.optimize.minimize run ends (the optimiser cannot see the batching)."""
    rs = np.random.RandomState(2)
    X = np.sort(rs.rand(25)) * 3
    y = np.sin(2 * X) + 0.05 * rs.randn(25)
    outs = []
    for batched in (True, False):
        k = g.SquaredExponentialKernel(initial_params=[1.0, 0.5], param_bounds=[(0.1, 5), (0.05, 3)])
        gp = with_fake(g.GaussianProcess(k, X=X, y=y, err_y=0.05, use_hyper_deriv=True))
        np.random.seed(11)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res, n = gp.optimize_hyperparameters(random_starts=5, batched_starts=batched,
                                                 opt_kwargs={"options": {"maxiter": 30}})
        outs.append((res, n, list(gp._dev_obj.calls), np.array(gp.free_params[:], dtype=float)))
    (rb, nb, calls_b, pb), (rq, nq, calls_q, pq) = outs
    assert nb == nq == 5
    assert_close(rb.x, rq.x, rtol=1e-10)
    assert_close(rb.fun, rq.fun, rtol=1e-12)
    assert_close(pb, pq, rtol=1e-10)                        # the GP is left at the optimum in both modes
    assert calls_b.count("ll_batched") > 0 and calls_q.count("ll_batched") == 0
    assert calls_b.count("ll_batched") < calls_q.count("ll") / 2   # one launch serves all waiting starts


def test_compute_from_MCMC_uses_one_batched_prediction_and_matches_the_per_sample_loop():
    """gaussian_process.py:1944-1969: update_hyperparameters + predict per retained sample.  Mean / std requests go
    through ONE batched device call (gpt_predict_batched); the per-sample loop (still used for covariances / samples)
    gives the same lists, samples outside the prior support are dropped by both, predict_MCMC combines them by the law of
    total variance."""
    gp = with_fake(_kat1_gp())
    rs = np.random.RandomState(5)
    th0 = np.array(gp.free_params[:], dtype=float)
    trace = th0 * np.exp(0.05 * rs.randn(6, len(th0)))
    trace[2, 0] = -1.0                                     # zero prior probability: predicted all the same (reference)
    trace[4, 0] = 1e170                                    # sigma_f^2 overflows: not positive definite, dropped
    Xs = rs.rand(5, 2)
    n_calls = len(gp._dev_obj.calls)
    res_b = gp.compute_from_MCMC(Xs, flat_trace=trace)
    calls = gp._dev_obj.calls[n_calls:]
    assert calls.count("predict_batched") == 1 and calls.count("predict") == 0
    gp._mcmc_predict_by_loop = True
    res_l = gp.compute_from_MCMC(Xs, flat_trace=trace)
    gp._mcmc_predict_by_loop = False
    assert len(res_b['mean']) == len(res_l['mean']) == 5   # six samples, one dropped
    for a, b in zip(res_b['mean'], res_l['mean']):
        assert_close(a, b, rtol=1e-10, atol=1e-12)
    for a, b in zip(res_b['std'], res_l['std']):
        assert_close(a, b, rtol=1e-7, atol=1e-10)
    assert_close(np.asarray(gp.free_params[:], float), th0, rtol=0, atol=0)   # the GP's own parameters are untouched
    out = gp.predict_MCMC(Xs, flat_trace=trace)
    means = np.array(res_l['mean'])
    stds = np.array(res_l['std'])
    assert_close(out['mean'], means.mean(axis=0), rtol=1e-10)
    assert_close(out['std'], np.sqrt((stds ** 2).mean(axis=0) + np.var(means, axis=0, ddof=1)), rtol=1e-7)
    # derivative predictions and per-theta mean-function parameters take the same route
    res_d = gp.compute_from_MCMC(Xs, n=np.tile([1, 0], (5, 1)), flat_trace=trace[:2])
    gp._mcmc_predict_by_loop = True
    res_dl = gp.compute_from_MCMC(Xs, n=np.tile([1, 0], (5, 1)), flat_trace=trace[:2])
    gp._mcmc_predict_by_loop = False
    for a, b in zip(res_d['mean'], res_dl['mean']):
        assert_close(a, b, rtol=1e-9, atol=1e-11)


def test_predict_MCMC_matches_the_reference():
    """The reference's own predict_MCMC / compute_from_MCMC on a 12-sample trace (golden mcmc_predict_se1d; one sample outside
    the prior support, which its per-sample wrapper predicts all the same; burn / thin applied to a given trace): per-sample
    means and standard deviations, the marginalised mean / std (law of total variance), derivative predictions."""
    gd = load_golden("mcmc_predict_se1d")
    k = g.SquaredExponentialKernel(initial_params=[1.0, 0.7], param_bounds=[(0.05, 5), (0.1, 3)])
    gp = with_fake(g.GaussianProcess(k))
    nv = int((gd["n"][:, 0] == 0).sum())
    gp.add_data(gd["X"][:nv, 0], gd["y"][:nv], err_y=gd["err_y"][:nv])
    gp.add_data(gd["X"][nv:, 0], gd["y"][nv:], err_y=gd["err_y"][nv:], n=1)
    for by_loop in (False, True):
        gp._mcmc_predict_by_loop = by_loop
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
