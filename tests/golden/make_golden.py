"""Generate golden vectors from the UNMODIFIED reference (run in the dev container only).

    python tests/golden/make_golden.py

Loads /root/reference/gptools through oracle/ref_shim.py and dumps, for a set of
seeded inputs covering SURVEY.md section 8 rows a1-a10 / configs C1-C5 (at sizes the
reference evaluates in seconds), the reference's K entries, ll, ll gradient,
alpha, predictive mean/std/cov and draw_sample output into tests/golden/*.npz.
The library versions that produced the numbers are stored in every file.
"""
import json
import os
import pickle
import sys
import warnings

import numpy as np
import scipy
from numpy.random import RandomState

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle.ref_shim import load_reference, REF_ROOT  # noqa: E402

g = load_reference()

META = json.dumps({
    "numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0],
    "reference": "markchil/gptools %s (unmodified, via oracle/ref_shim.py)" % g.__version__,
})


def save(name, **arrs):
    arrs["meta"] = np.array(META)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print("wrote", name, {k: np.shape(v) for k, v in arrs.items() if k != "meta"})


def gp_state(gp):
    """Inputs exactly as the reference holds them after add_data."""
    d = {"X": gp.X, "n": gp.n, "y": gp.y, "err_y": gp.err_y}
    if gp.T is not None:
        d["T"] = gp.T
    return d


def ll_and_grad(gp, with_grad):
    gp.use_hyper_deriv = with_grad
    gp.K_up_to_date = False
    gp.compute_K_L_alpha_ll()
    out = {"ll": gp.ll, "log_prior": gp.hyperprior(gp.params), "alpha": gp.alpha.ravel(),
           "K": gp.K, "L": gp.L}
    if with_grad:
        out["ll_deriv"] = gp.ll_deriv
    return out


# ---------------------------------------------------------------- KAT-1: SE 2-D + gradients
def case_se2d():
    rs = RandomState(0)
    X = rs.rand(6, 2)
    y = np.sin(X).sum(1)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.7, 1.1], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k)
    gp.add_data(X, y, err_y=0.01)
    gp.add_data(X, np.cos(X[:, 0]), n=np.tile([1, 0], (6, 1)), err_y=0.01)
    gp.add_data(X, np.cos(X[:, 1]), n=np.tile([0, 1], (6, 1)), err_y=0.01)
    out = ll_and_grad(gp, True)
    for p in range(3):
        out["dK%d" % p] = gp.compute_Kij(gp.X, None, gp.n, None, hyper_deriv=p)
    Xs = rs.rand(4, 2)
    res = gp.predict(Xs, full_output=True)
    out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
    res = gp.predict(Xs, n=np.tile([1, 0], (4, 1)), full_output=True)
    out.update(mean_d1=res["mean"], std_d1=res["std"], cov_d1=res["cov"])
    out["Kstar"] = gp.compute_Kij(gp.X, Xs, gp.n, np.zeros((4, 2), dtype=int))
    save("se2d_kat1", params=k.params.copy(), **gp_state(gp), **out)


# ---------------------------------------------------------------- SE pair list, orders 0-4
def case_se_pairs():
    rs = RandomState(11)
    npair = 600
    for D in (1, 2, 3):
        Xi = rs.randn(npair, D)
        Xj = rs.randn(npair, D)
        Xj[:40] = Xi[:40]                      # tau == 0 pairs
        Xj[40:60, 0] = Xi[40:60, 0]            # tau == 0 in one dimension only
        ni = rs.randint(0, 3, size=(npair, D))
        nj = rs.randint(0, 3, size=(npair, D))
        ni[100:200] = 0
        nj[200:300] = 0
        params = np.concatenate(([1.7], 0.5 + rs.rand(D)))
        k = g.SquaredExponentialKernel(num_dim=D, initial_params=params, param_bounds=[(0, 10)] * (D + 1))
        out = {"val": k(Xi, Xj, ni, nj)}
        for p in range(D + 1):
            out["hd%d" % p] = k(Xi, Xj, ni, nj, hyper_deriv=p)
        # value-only list (the reference's fast path, squared_exponential.py:110-113,169-171)
        z = np.zeros_like(ni)
        out["val0"] = k(Xi, Xj, z, z)
        for p in range(D + 1):
            out["hd0_%d" % p] = k(Xi, Xj, z, z, hyper_deriv=p)
        save("se_pairs_D%d" % D, Xi=Xi, Xj=Xj, ni=ni, nj=nj, params=params, **out)


# ---------------------------------------------------------------- KAT-2: Matern52 1-D, + 2-D K (tests/test_matern.py shape)
def case_matern52():
    rs = RandomState(1)
    X = np.sort(rs.rand(8))
    k = g.Matern52Kernel(num_dim=1, initial_params=[2.0, 0.4], param_bounds=[(0, 10)] * 2)
    gp = g.GaussianProcess(k)
    gp.add_data(X, np.sin(5 * X), err_y=0.02)
    gp.add_data(X[::2], 5 * np.cos(5 * X[::2]), n=1, err_y=0.05)
    out = ll_and_grad(gp, False)
    Xs = np.array([0.1, 0.5, 0.9])
    res = gp.predict(Xs, full_output=True)
    out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
    res = gp.predict(Xs, n=1, full_output=True)
    out.update(mean_d1=res["mean"], std_d1=res["std"], cov_d1=res["cov"])
    save("matern52_kat2", params=k.params.copy(), **gp_state(gp), **out)

    # the reference's own test (tests/test_matern.py:4-31) with seeded length scales
    rs = RandomState(0)
    X2 = rs.randn(5, 2)
    ls = np.exp(RandomState(5).randn(2) * 0.5)
    k2 = g.Matern52Kernel(num_dim=2, initial_params=[1.0, ls[0], ls[1]], param_bounds=[(0, 10)] * 3)
    gp2 = g.GaussianProcess(k2)
    yv = rs.randn(5)
    gp2.add_data(X2, yv, err_y=0.1)
    gp2.add_data(X2, rs.randn(5), n=np.tile([1, 0], (5, 1)), err_y=0.1)
    gp2.add_data(X2, rs.randn(5), n=np.tile([0, 1], (5, 1)), err_y=0.1)
    out = ll_and_grad(gp2, False)
    # generic Matern at nu = 5/2 on the same data (kernel/matern.py:251-465)
    k3 = g.MaternKernel(num_dim=2, initial_params=[1.0, 2.5, ls[0], ls[1]], param_bounds=[(0, 10)] * 4)
    K_generic = g.GaussianProcess(k3).compute_Kij(gp2.X, None, gp2.n, None)
    save("matern52_2d_testshape", params=k2.params.copy(), params_generic=k3.params.copy(),
         K_generic=K_generic, **gp_state(gp2), **out)


# ---------------------------------------------------------------- generic Matern incl. the series zone
def case_matern_generic():
    rs = RandomState(3)
    X = np.sort(rs.rand(30)) * 3.0
    X[5] = X[4] + 1e-3       # y = 2 nu r^2 inside (0, 5e-4]
    X[11] = X[10] + 4e-3
    X[20] = X[19]            # exact duplicate -> origin limits
    for nu in (2.5, 3.5, 1.5):
        k = g.MaternKernel(num_dim=1, initial_params=[1.4, nu, 0.6], param_bounds=[(0, 10)] * 3)
        gp = g.GaussianProcess(k)
        gp.add_data(X, np.sin(2 * X), err_y=0.05)
        if nu > 2:
            gp.add_data(X[::3], 2 * np.cos(2 * X[::3]), n=1, err_y=0.05)
        out = ll_and_grad(gp, False)
        Xs = np.linspace(0.05, 2.95, 7)
        res = gp.predict(Xs, full_output=True)
        out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
        save("matern_generic_nu%s" % str(nu).replace(".", "p"), params=k.params.copy(), **gp_state(gp), **out)
    # 2-D, value + both gradient components, with near-coincident points
    rs = RandomState(4)
    X2 = rs.rand(12, 2)
    X2[3] = X2[2] + [2e-3, 0.0]
    X2[7] = X2[6]
    k = g.MaternKernel(num_dim=2, initial_params=[0.9, 2.5, 0.5, 0.8], param_bounds=[(0, 10)] * 4)
    gp = g.GaussianProcess(k)
    gp.add_data(X2, np.sin(X2).sum(1), err_y=0.05)
    gp.add_data(X2, np.cos(X2[:, 0]), n=np.tile([1, 0], (12, 1)), err_y=0.05)
    gp.add_data(X2, np.cos(X2[:, 1]), n=np.tile([0, 1], (12, 1)), err_y=0.05)
    out = ll_and_grad(gp, False)
    save("matern_generic_2d", params=k.params.copy(), **gp_state(gp), **out)


def case_matern_real_nu():
    """Round 2: orders that are not half-integers (K_nu of real order on the device) and an integer order (the
    reference averages its power series over nu -+ 0.001 there, utils.py:1480-1484, 1498-1502)."""
    rs = RandomState(3)
    X = np.sort(rs.rand(30)) * 3.0
    X[5] = X[4] + 1e-3       # y = 2 nu r^2 inside (0, 5e-4]
    X[11] = X[10] + 4e-3
    X[20] = X[19]            # exact duplicate -> origin limits
    for nu in (2.2, 3.0):
        k = g.MaternKernel(num_dim=1, initial_params=[1.4, nu, 0.6], param_bounds=[(0, 10)] * 3)
        gp = g.GaussianProcess(k)
        gp.add_data(X, np.sin(2 * X), err_y=0.05)
        gp.add_data(X[::3], 2 * np.cos(2 * X[::3]), n=1, err_y=0.05)
        out = ll_and_grad(gp, False)
        Xs = np.linspace(0.05, 2.95, 7)
        res = gp.predict(Xs, full_output=True)
        out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
        resd = gp.predict(Xs, n=1, full_output=True)
        out.update(dmean=resd["mean"], dstd=resd["std"])
        save("matern_generic_nu%s" % str(nu).replace(".", "p"), params=k.params.copy(), **gp_state(gp), **out)
    rs = RandomState(4)
    X2 = rs.rand(12, 2)
    X2[3] = X2[2] + [2e-3, 0.0]
    X2[7] = X2[6]
    k = g.MaternKernel(num_dim=2, initial_params=[0.9, 2.2, 0.5, 0.8], param_bounds=[(0, 10)] * 4)
    gp = g.GaussianProcess(k)
    gp.add_data(X2, np.sin(X2).sum(1), err_y=0.05)
    gp.add_data(X2, np.cos(X2[:, 0]), n=np.tile([1, 0], (12, 1)), err_y=0.05)
    gp.add_data(X2, np.cos(X2[:, 1]), n=np.tile([0, 1], (12, 1)), err_y=0.05)
    out = ll_and_grad(gp, False)
    save("matern_generic_2d_nu2p2", params=k.params.copy(), **gp_state(gp), **out)
    # hyper-derivative oracle (Richardson differences of the reference's ll and K), smooth inputs
    rs = RandomState(3)
    X = np.sort(rs.rand(14)) * 3.0
    for nu in (2.2, 3.0):
        k = g.MaternKernel(num_dim=1, initial_params=[1.4, nu, 0.6], param_bounds=[(0, 10)] * 3)
        gp = g.GaussianProcess(k)
        gp.add_data(X, np.sin(2 * X), err_y=0.05)
        gp.add_data(X[::3], 2 * np.cos(2 * X[::3]), n=1, err_y=0.05)
        _hyperfd("hyperfd_matern_generic_nu%s" % str(nu).replace(".", "p"), gp, [0, 2], 4e-3)


def case_hyper_mp():
    """Hyper-parameter derivatives of the covariance pinned BEYOND finite-difference accuracy: the reference's own
    mpmath covariance functions (kernel/matern.py:44-100 matern_function, kernel/gibbs.py:41-160 GibbsFunction1dArb with
    tanh_warp_arb -- the functions its ArbitraryKernel differentiates with mpmath.diff, kernel/core.py:946-953) are
    differentiated with mpmath.diff at 40 digits with respect to the inputs (observation derivative orders) AND one
    hyperparameter.  Pairs are kept outside the generic Matern kernel's series zone (y > 5e-4), where the closed
    form is the exact function."""
    import mpmath
    from gptools.kernel.matern import matern_function
    from gptools.kernel.gibbs import GibbsFunction1dArb, tanh_warp_arb
    mpmath.mp.dps = 40
    rs = RandomState(11)

    def run(name, fun, D, params, hyper_idx, npairs):
        Xi = rs.rand(npairs, D) * 2.0
        Xj = Xi + (0.05 + 0.6 * rs.rand(npairs, D)) * rs.choice([-1.0, 1.0], size=(npairs, D))
        orders = []
        for q in range(npairs):
            oi = np.zeros(D, dtype=int)
            oj = np.zeros(D, dtype=int)
            c = q % 4
            if c in (1, 3):
                oi[rs.randint(D)] = 1
            if c in (2, 3):
                oj[rs.randint(D)] = 1
            orders.append((oi, oj))
        ni = np.array([o[0] for o in orders])
        nj = np.array([o[1] for o in orders])
        val = np.zeros(npairs)
        dval = np.zeros((len(hyper_idx), npairs))

        def f(*a):
            xi, xj, th = a[:D], a[D:2 * D], a[2 * D:]
            if D == 1:
                return fun(xi[0], xj[0], *th)
            return fun(tuple(xi), tuple(xj), *th)
        for q in range(npairs):
            point = [mpmath.mpf(float(v)) for v in Xi[q]] + [mpmath.mpf(float(v)) for v in Xj[q]] + \
                    [mpmath.mpf(float(v)) for v in params]
            base = tuple(int(v) for v in ni[q]) + tuple(int(v) for v in nj[q])
            zero = (0,) * len(params)
            val[q] = float(mpmath.diff(f, tuple(point), base + zero)) if sum(base) else float(f(*point))
            for k_, pidx in enumerate(hyper_idx):
                hd = [0] * len(params)
                hd[pidx] = 1
                dval[k_, q] = float(mpmath.diff(f, tuple(point), base + tuple(hd)))
        save(name, params=np.array(params, dtype=float), Xi=Xi, Xj=Xj, ni=ni, nj=nj, val=val, idx=np.array(hyper_idx),
             dval=dval)

    run("hypermp_matern_nu2p5_1d", matern_function, 1, [1.4, 2.5, 0.6], [0, 2], 24)
    run("hypermp_matern_nu2p2_1d", matern_function, 1, [1.4, 2.2, 0.6], [0, 2], 24)
    run("hypermp_matern_nu2p5_2d", matern_function, 2, [0.9, 2.5, 0.5, 0.8], [0, 2, 3], 24)
    gibbs = GibbsFunction1dArb(tanh_warp_arb)
    run("hypermp_gibbs_tanh", gibbs, 1, [1.5, 0.6, 0.1, 0.05, 0.9], [0, 1, 2, 3, 4], 24)


# ---------------------------------------------------------------- KAT-3: Gibbs-tanh + T + draw_sample
def case_gibbs():
    k = g.GibbsKernel1dTanh(initial_params=[1.5, 0.6, 0.1, 0.05, 0.9],
                            param_bounds=[(0, 10), (0, 5), (0, 5), (0, 1), (0, 2)])
    Xq = np.linspace(0, 1.1, 12)
    T = np.zeros((3, 12))
    T[0, :6] = T[1, 3:9] = T[2, 6:] = 1 / 6.0
    gp = g.GaussianProcess(k)
    gp.add_data(Xq, [2.5, 2.0, 1.0], err_y=0.05, T=T)
    gp.add_data(0, 0, n=1)
    out = ll_and_grad(gp, False)
    Xs = np.array([0.0, 0.5, 1.0])
    res = gp.predict(Xs, full_output=True)
    out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
    res1 = gp.predict(Xs, n=1, full_output=True)
    out.update(mean_d1=res1["mean"], std_d1=res1["std"], cov_d1=res1["cov"])
    rv = RandomState(2).randn(3, 2)
    out["rand_vars"] = rv
    out["draw"] = gp.draw_sample(Xs, rand_vars=rv, method="cholesky")
    save("gibbs_kat3", params=k.params.copy(), **gp_state(gp), **out)

    # C5-shaped, reduced: 400 quadrature points -> 50 line integrals + core slope constraint
    rs = RandomState(0)
    Nq, Mo, W = 400, 50, 40
    Xq = np.linspace(0, 1.1, Nq)
    T = np.zeros((Mo, Nq))
    starts = rs.randint(0, Nq - W, size=Mo)
    for i, s in enumerate(starts):
        T[i, s:s + W] = 1.1 / Nq
    yy = rs.rand(Mo) * 0.3 + 0.1
    gp = g.GaussianProcess(k)
    gp.add_data(Xq, yy, err_y=0.02, T=T)
    gp.add_data(0, 0, n=1)
    out = ll_and_grad(gp, False)
    del out["K"]
    Xs = np.linspace(0, 1.1, 40)
    res = gp.predict(Xs, full_output=True)
    rv = rs.randn(40, 8)
    out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"], rand_vars=rv,
               draw=gp.draw_sample(Xs, rand_vars=rv, method="cholesky"))
    save("gibbs_c5_small", params=k.params.copy(), **gp_state(gp), **out)


# ---------------------------------------------------------------- config 5 at full size (4001 latent -> 501 observations)
def case_c5_full():
    k = g.GibbsKernel1dTanh(initial_params=[1.5, 0.6, 0.1, 0.05, 0.9],
                            param_bounds=[(0, 10), (0, 5), (0, 5), (0, 1), (0, 2)])
    rs = RandomState(0)
    Nq, Mo, W = 4000, 500, 400
    Xq = np.linspace(0, 1.1, Nq)
    T = np.zeros((Mo, Nq))
    starts = rs.randint(0, Nq - W, size=Mo)
    for i, s in enumerate(starts):
        T[i, s:s + W] = 1.1 / Nq
    yy = rs.rand(Mo) * 0.3 + 0.1
    gp = g.GaussianProcess(k)
    gp.add_data(Xq, yy, err_y=0.02, T=T)
    gp.add_data(0, 0, n=1)
    gp.compute_K_L_alpha_ll()
    Xs = np.linspace(0, 1.1, 400)
    res = gp.predict(Xs, full_output=True)
    rv = rs.randn(400, 4)
    draw = gp.draw_sample(Xs, rand_vars=rv, method="cholesky", mean=res["mean"], cov=res["cov"])
    # inputs are regenerated by the test from the same seeds (T alone is 16 MB); only results are stored
    save("gibbs_c5_full", params=k.params.copy(), starts=starts, y=gp.y, ll=gp.ll, log_prior=gp.hyperprior(gp.params),
         alpha=gp.alpha.ravel(), Xs=Xs, mean=res["mean"], std=res["std"], cov_diag=np.diag(res["cov"]).copy(),
         rand_vars=rv, draw=draw)


# ---------------------------------------------------------------- KAT-4: demo / config 1
def case_demo():
    with open(os.path.join(REF_ROOT, "demo", "sample_data_core.pkl"), "rb") as f:
        d = pickle.load(f, encoding="latin1")
    X, y, err = (np.asarray(d[kk], dtype=float) for kk in ("X", "y", "err_y"))
    hp = g.UniformJointPrior([(0, 20)]) * g.GammaJointPriorAlt([1.0], [0.7])
    k = g.SquaredExponentialKernel(initial_params=[1.8849006111246833, 0.97760159723344708], hyperprior=hp)
    gp = g.GaussianProcess(k)
    gp.add_data(X, y, err_y=err)
    gp.add_data(0, 0, n=1)
    out = ll_and_grad(gp, True)
    Xs = np.linspace(0, 1.1, 400)
    m, s = gp.predict(Xs)
    m1, s1 = gp.predict(Xs, n=1)
    out.update(Xs=Xs, mean=m, std=s, mean_d1=m1, std_d1=s1)
    save("demo_c1_kat4", params=k.params.copy(), **gp_state(gp), **out)

    # synthetic 200-point variant
    rs = RandomState(0)
    X = np.sort(rs.rand(200)) * 1.05
    y = 3 - 1.2 * X ** 2 + 0.1 * rs.randn(200)
    hp = g.UniformJointPrior([(0, 20)]) * g.GammaJointPriorAlt([1.0], [0.7])
    k = g.SquaredExponentialKernel(initial_params=[1.9, 1.0], hyperprior=hp)
    gp = g.GaussianProcess(k)
    gp.add_data(X, y, err_y=0.1)
    gp.add_data(0, 0, n=1)
    out = ll_and_grad(gp, True)
    del out["K"], out["L"]
    m, s = gp.predict(Xs)
    m1, s1 = gp.predict(Xs, n=1)
    out.update(Xs=Xs, mean=m, std=s, mean_d1=m1, std_d1=s1)
    save("c1_synth200", params=k.params.copy(), **gp_state(gp), **out)


# ---------------------------------------------------------------- KAT-5: config 3 (headline), two theta of the batch
def c3_problem():
    rs = RandomState(0)
    Xv = rs.rand(256, 2)
    Xd1 = rs.rand(128, 2)
    Xd2 = rs.rand(128, 2)
    f = lambda x: np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k)
    gp.add_data(Xv, f(Xv) + 0.05 * rs.randn(256), err_y=0.05)
    gp.add_data(Xd1, 3 * np.cos(3 * Xd1[:, 0]) * np.cos(2 * Xd1[:, 1]) + 0.05 * rs.randn(128), err_y=0.05,
                n=np.tile([1, 0], (128, 1)))
    gp.add_data(Xd2, -2 * np.sin(3 * Xd2[:, 0]) * np.sin(2 * Xd2[:, 1]) + 0.05 * rs.randn(128), err_y=0.05,
                n=np.tile([0, 1], (128, 1)))
    th = np.array([1.0, 0.3, 0.4]) * np.exp(0.1 * RandomState(1).randn(4096, 3))
    return gp, th


def case_c3():
    gp, th = c3_problem()
    gp.use_hyper_deriv = True
    idx = np.array([0, 1, 2, 3, 1000, 2047, 4094, 4095])
    lls, grads = [], []
    for b in idx:
        f, gr = gp.update_hyperparameters(th[b])
        lls.append(-f)
        grads.append(-gr)
    gp.update_hyperparameters(th[0])
    Xs = RandomState(2).rand(5, 2)
    m, s = gp.predict(Xs)
    save("c3_kat5", theta_idx=idx, theta=th[idx], ll=np.array(lls), ll_deriv=np.array(grads),
         log_prior=gp.hyperprior(gp.params), Xs=Xs, mean=m, std=s, alpha0=gp.alpha.ravel(),
         **gp_state(gp))


# ---------------------------------------------------------------- config 2, reduced (Matern52 and generic Matern, 1-D)
def case_c2():
    rs = RandomState(0)
    X = np.sort(rs.rand(200)) * 10
    yv = np.sin(X) + 0.05 * rs.randn(200)
    yd = np.cos(X) + 0.05 * rs.randn(200)
    Xs = np.linspace(0, 10, 500)
    for name, k in (
        ("c2_small_matern52", g.Matern52Kernel(num_dim=1, initial_params=[1.0, 0.8], param_bounds=[(0, 10)] * 2)),
        ("c2_small_matern_generic", g.MaternKernel(num_dim=1, initial_params=[1.0, 2.5, 0.8],
                                                   param_bounds=[(0, 10)] * 3)),
    ):
        gp = g.GaussianProcess(k)
        gp.add_data(X, yv, err_y=0.05)
        gp.add_data(X, yd, err_y=0.05, n=1)
        out = ll_and_grad(gp, False)
        del out["K"], out["L"]
        m, s = gp.predict(Xs)
        m1, s1 = gp.predict(Xs, n=1)
        out.update(Xs=Xs, mean=m, std=s, mean_d1=m1, std_d1=s1)
        save(name, params=k.params.copy(), **gp_state(gp), **out)


# ---------------------------------------------------------------- DiagonalNoiseKernel + gradient (gaussian_process.py:1480-1488)
def case_noise():
    rs = RandomState(7)
    X = rs.rand(20, 2)
    y = np.sin(3 * X[:, 0]) + X[:, 1] + 0.1 * rs.randn(20)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.1, 0.4, 0.6], param_bounds=[(0, 10)] * 3)
    nk = g.DiagonalNoiseKernel(num_dim=2, initial_noise=0.12, noise_bound=(0, 5))
    gp = g.GaussianProcess(k, noise_k=nk)
    gp.add_data(X, y, err_y=0.03)
    gp.add_data(X[:5], np.ones(5), n=np.tile([0, 1], (5, 1)), err_y=0.2)
    out = ll_and_grad(gp, True)
    Xs = rs.rand(6, 2)
    res = gp.predict(Xs, full_output=True)
    out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
    resn = gp.predict(Xs, noise=True, full_output=True)
    out.update(mean_noise=resn["mean"], std_noise=resn["std"], cov_noise=resn["cov"])
    save("se_diagnoise", params=k.params.copy(), noise_sigma=nk.params.copy(), **gp_state(gp), **out)


# ---------------------------------------------------------------- hyper-derivative oracle
def _fd(f, theta, i, h):
    """Central difference with one Richardson step: (4 D(h/2) - D(h)) / 3, error O(h^4)."""
    def D(hh):
        tp, tm = theta.copy(), theta.copy()
        tp[i] += hh
        tm[i] -= hh
        return (f(tp) - f(tm)) / (2.0 * hh)
    return (4.0 * D(h / 2.0) - D(h)) / 3.0


def _hyperfd(name, gp, idx, rel_h):
    """The reference raises NotImplementedError for hyper_deriv of these kernels (kernel/matern.py:543,
    kernel/core.py:723, kernel/gibbs.py:319), so the golden gradient is the finite difference of the
    reference's own ll (compute_K_L_alpha_ll, hyperprior included -- uniform here, so constant) and K."""
    k = gp.k
    theta0 = np.array(k.params, dtype=float)

    def set_theta(th):
        k.params[:] = th
        gp.K_up_to_date = False

    def ll(th):
        set_theta(th)
        gp.compute_K_L_alpha_ll()
        return float(gp.ll)

    def Kmat(th):
        set_theta(th)
        return np.array(gp.compute_Kij(gp.X, None, gp.n, None))

    grad = np.array([_fd(ll, theta0, i, rel_h * abs(theta0[i])) for i in idx])
    dK = np.array([_fd(Kmat, theta0, i, rel_h * abs(theta0[i])) for i in idx])
    set_theta(theta0)
    gp.compute_K_L_alpha_ll()
    save(name, params=theta0, idx=np.array(idx, dtype=int), ll=gp.ll, ll_grad_fd=grad, dK_fd=dK,
         rel_h=rel_h, **gp_state(gp))


def case_hyperfd():
    # Matern 5/2, 1-D, values + derivatives (inputs of matern52_kat2)
    rs = RandomState(1)
    X = np.sort(rs.rand(8))
    k = g.Matern52Kernel(num_dim=1, initial_params=[2.0, 0.4], param_bounds=[(0, 10)] * 2)
    gp = g.GaussianProcess(k)
    gp.add_data(X, np.sin(5 * X), err_y=0.02)
    gp.add_data(X[::2], 5 * np.cos(5 * X[::2]), n=1, err_y=0.05)
    _hyperfd("hyperfd_matern52_1d", gp, [0, 1], 1e-3)
    # Matern 5/2, 2-D, value + both gradient components (inputs of matern52_2d_testshape)
    rs = RandomState(0)
    X2 = rs.randn(5, 2)
    ls = np.exp(RandomState(5).randn(2) * 0.5)
    k2 = g.Matern52Kernel(num_dim=2, initial_params=[1.0, ls[0], ls[1]], param_bounds=[(0, 10)] * 3)
    gp2 = g.GaussianProcess(k2)
    gp2.add_data(X2, rs.randn(5), err_y=0.1)
    gp2.add_data(X2, rs.randn(5), n=np.tile([1, 0], (5, 1)), err_y=0.1)
    gp2.add_data(X2, rs.randn(5), n=np.tile([0, 1], (5, 1)), err_y=0.1)
    _hyperfd("hyperfd_matern52_2d", gp2, [0, 1, 2], 1e-3)
    # generic Matern (nu fixed), smooth inputs (no series-zone pairs: the reference's kvp round-off would
    # dominate a finite difference there), values + derivatives where nu allows
    rs = RandomState(3)
    X = np.sort(rs.rand(14)) * 3.0
    for nu in (2.5, 3.5, 1.5):
        k = g.MaternKernel(num_dim=1, initial_params=[1.4, nu, 0.6], param_bounds=[(0, 10)] * 3)
        gp = g.GaussianProcess(k)
        gp.add_data(X, np.sin(2 * X), err_y=0.05)
        if nu > 2:
            gp.add_data(X[::3], 2 * np.cos(2 * X[::3]), n=1, err_y=0.05)
        _hyperfd("hyperfd_matern_generic_nu%s" % str(nu).replace(".", "p"), gp, [0, 2], 4e-3)
    rs = RandomState(4)
    X2 = rs.rand(8, 2)
    k = g.MaternKernel(num_dim=2, initial_params=[0.9, 2.5, 0.5, 0.8], param_bounds=[(0, 10)] * 4)
    gp = g.GaussianProcess(k)
    gp.add_data(X2, np.sin(X2).sum(1), err_y=0.05)
    gp.add_data(X2, np.cos(X2[:, 0]), n=np.tile([1, 0], (8, 1)), err_y=0.05)
    _hyperfd("hyperfd_matern_generic_2d", gp, [0, 2, 3], 4e-3)
    # Gibbs-tanh: direct observations incl. a derivative, and the T path (inputs of gibbs_kat3)
    k = g.GibbsKernel1dTanh(initial_params=[1.5, 0.6, 0.1, 0.05, 0.9],
                            param_bounds=[(0, 10), (0, 5), (0, 5), (0, 1), (0, 2)])
    Xd = np.linspace(0.05, 1.05, 11)
    gp = g.GaussianProcess(k)
    gp.add_data(Xd, 2.5 - 1.5 * np.tanh((Xd - 0.9) / 0.1), err_y=0.05)
    gp.add_data(Xd[::4], -1.0 * np.ones(3), n=1, err_y=0.3)
    gp.add_data(0, 0, n=1)
    _hyperfd("hyperfd_gibbs_direct", gp, [0, 1, 2, 3, 4], 1e-3)
    k = g.GibbsKernel1dTanh(initial_params=[1.5, 0.6, 0.1, 0.05, 0.9],
                            param_bounds=[(0, 10), (0, 5), (0, 5), (0, 1), (0, 2)])
    Xq = np.linspace(0, 1.1, 12)
    T = np.zeros((3, 12))
    T[0, :6] = T[1, 3:9] = T[2, 6:] = 1 / 6.0
    gp = g.GaussianProcess(k)
    gp.add_data(Xq, [2.5, 2.0, 1.0], err_y=0.05, T=T)
    gp.add_data(0, 0, n=1)
    _hyperfd("hyperfd_gibbs_T", gp, [0, 1, 2, 3, 4], 1e-3)


# ---------------------------------------------------------------- kernel algebra: ProductKernel (kernel/core.py:601-670)
def case_product():
    rs = RandomState(11)
    k1 = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.4, 0.9], param_bounds=[(0, 10)] * 3)
    k2 = g.SquaredExponentialKernel(num_dim=2, initial_params=[0.8, 1.1, 0.5], param_bounds=[(0, 10)] * 3)
    kp = k1 * k2
    ks = k1 + k2
    Mp = 60
    Xi, Xj = rs.rand(Mp, 2), rs.rand(Mp, 2)
    Xj[:5] = Xi[:5]
    ni, nj = rs.randint(0, 3, size=(Mp, 2)), rs.randint(0, 2, size=(Mp, 2))
    save("kernel_algebra_se2d", Xi=Xi, Xj=Xj, ni=ni, nj=nj, params1=k1.params.copy(), params2=k2.params.copy(),
         prod=kp(Xi, Xj, ni, nj), sum=ks(Xi, Xj, ni, nj),
         sum_hd1=ks(Xi, Xj, ni, nj, hyper_deriv=1), sum_hd4=ks(Xi, Xj, ni, nj, hyper_deriv=4))


# ---------------------------------------------------------------- kernel algebra on the device: mixed trees, GP level
def case_composite():
    """(SE + Matern52) * SE in 2-D on pair lists with first-derivative orders, and a GaussianProcess whose kernel is
    SE * Matern52 + SE in 1-D with value and derivative observations (ll, alpha, K, predictions) -- the reference's own
    SumKernel / ProductKernel (kernel/core.py:549-670)."""
    rs = RandomState(41)
    ka = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.4, 0.9], param_bounds=[(0, 10)] * 3)
    kb = g.Matern52Kernel(num_dim=2, initial_params=[0.7, 0.8, 0.6], param_bounds=[(0, 10)] * 3)
    kc = g.SquaredExponentialKernel(num_dim=2, initial_params=[0.9, 1.5, 1.2], param_bounds=[(0, 10)] * 3)
    k = (ka + kb) * kc
    Mp = 80
    Xi, Xj = rs.rand(Mp, 2), rs.rand(Mp, 2)
    Xj[:6] = Xi[:6]
    ni, nj = np.zeros((Mp, 2), dtype=int), np.zeros((Mp, 2), dtype=int)
    ni[np.arange(Mp), rs.randint(0, 2, Mp)] = rs.randint(0, 2, Mp)   # at most one first derivative per side
    nj[np.arange(Mp), rs.randint(0, 2, Mp)] = rs.randint(0, 2, Mp)
    save("composite_pairs_2d", Xi=Xi, Xj=Xj, ni=ni, nj=nj, params=np.array([float(v) for v in k.params]),
         K=k(Xi, Xj, ni, nj))
    # GP level, 1-D
    k1 = g.SquaredExponentialKernel(initial_params=[1.1, 0.9], param_bounds=[(0, 10)] * 2)
    k2 = g.Matern52Kernel(initial_params=[0.8, 0.5], param_bounds=[(0, 10)] * 2)
    k3 = g.SquaredExponentialKernel(initial_params=[0.4, 0.15], param_bounds=[(0, 10)] * 2)
    kg = k1 * k2 + k3
    X = np.sort(rs.rand(40)) * 3.0
    y = np.sin(2 * X) + 0.3 * np.sin(15 * X) + 0.05 * rs.randn(40)
    gp = g.GaussianProcess(kg)
    gp.add_data(X, y, err_y=0.05)
    gp.add_data(X[::5], 2 * np.cos(2 * X[::5]), err_y=0.3, n=1)
    out = ll_and_grad(gp, False)
    Xs = np.linspace(0.0, 3.0, 11)
    res = gp.predict(Xs, full_output=True)
    out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
    res1 = gp.predict(Xs, n=1, full_output=True)
    out.update(mean_d1=res1["mean"], std_d1=res1["std"])
    save("composite_gp_1d", params=np.array([float(v) for v in kg.params]), **gp_state(gp), **out)
    # SE + SE in 2-D with value and derivative observations and the reference's OWN ll gradient (SumKernel forwards
    # hyper_deriv to the operand that owns the parameter, kernel/core.py:576-582; both operands are SE kernels)
    ks = (g.SquaredExponentialKernel(num_dim=2, initial_params=[1.2, 0.5, 0.7], param_bounds=[(0, 10)] * 3) +
          g.SquaredExponentialKernel(num_dim=2, initial_params=[0.4, 0.15, 0.2], param_bounds=[(0, 10)] * 3))
    Xg = rs.rand(30, 2)
    gps = g.GaussianProcess(ks)
    gps.add_data(Xg, np.sin(3 * Xg[:, 0]) * np.cos(2 * Xg[:, 1]) + 0.05 * rs.randn(30), err_y=0.05)
    gps.add_data(Xg[::5], 3 * np.cos(3 * Xg[::5, 0]) * np.cos(2 * Xg[::5, 1]), err_y=0.2, n=np.tile([1, 0], (6, 1)))
    outs = ll_and_grad(gps, True)
    save("composite_sum_grad_2d", params=np.array([float(v) for v in ks.params]), **gp_state(gps), **outs)


# ---------------------------------------------------------------- compute_from_MCMC / predict_MCMC (gaussian_process.py:1840-2254)
def case_mcmc_predict():
    """predict_MCMC over a given trace of hyperparameter samples: the law of total variance over the per-sample
    predictions (one sample lies outside the prior support: the reference's per-sample wrapper predicts it all the same,
    gaussian_process.py:2301-2330).  num_proc = 2: with one process the reference iterates a Python-3 map object twice."""
    rs = RandomState(3)
    X = np.sort(rs.rand(40)) * 4
    y = np.sin(2 * X) + 0.1 * rs.randn(40)
    k = g.SquaredExponentialKernel(initial_params=[1.0, 0.7], param_bounds=[(0.05, 5), (0.1, 3)])
    gp = g.GaussianProcess(k, X=X, y=y, err_y=0.1)
    gp.add_data(X[::8], 2 * np.cos(2 * X[::8]), err_y=0.3, n=1)
    trace = np.array([1.0, 0.7]) * np.exp(0.1 * rs.randn(12, 2))
    trace[3, 1] = 9.0
    Xs = np.linspace(0, 4, 9)
    out = gp.predict_MCMC(Xs, flat_trace=trace, num_proc=2, return_samples=False)
    out1 = gp.predict_MCMC(Xs, n=1, flat_trace=trace, num_proc=2, return_samples=False)
    res = gp.compute_from_MCMC(Xs, flat_trace=trace, num_proc=2)
    thin = gp.predict_MCMC(Xs, flat_trace=trace, burn=2, thin=3, num_proc=2, return_samples=False)
    save("mcmc_predict_se1d", trace=trace, Xs=Xs, mean=out['mean'], std=out['std'], mean_d1=out1['mean'],
         std_d1=out1['std'], means=np.array(res['mean']), stds=np.array(res['std']), mean_thin=thin['mean'],
         std_thin=thin['std'], **gp_state(gp))


# ---------------------------------------------------------------- compute_ll_matrix (gaussian_process.py:1607-1692)
def case_ll_matrix():
    """The log-posterior over a regular grid of the free hyperparameters (KAT-1 problem: SE 2-D, value + both gradient
    observations), as the reference's own nested loop computes it."""
    rs = RandomState(0)
    X = rs.rand(6, 2)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.7, 1.1], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k)
    gp.add_data(X, np.sin(X).sum(1), err_y=0.01)
    gp.add_data(X, np.cos(X[:, 0]), n=np.tile([1, 0], (6, 1)), err_y=0.01)
    gp.add_data(X, np.cos(X[:, 1]), n=np.tile([0, 1], (6, 1)), err_y=0.01)
    bounds, num_pts = [(1.0, 1.5), (0.5, 0.9), (0.9, 1.3)], [2, 3, 2]
    ll_vals, pv = gp.compute_ll_matrix(bounds, num_pts)
    save("ll_matrix_kat1", bounds=np.array(bounds), num_pts=np.array(num_pts), ll_vals=ll_vals, p0=pv[0], p1=pv[1], p2=pv[2])


# ---------------------------------------------------------------- other Gibbs length-scale profiles (kernel/gibbs.py:508-902)
def case_gibbs_profiles():
    rs = RandomState(21)
    X = np.sort(rs.rand(24)) * 1.1
    y = 2.5 - 1.5 * np.tanh((X - 0.8) / 0.1) + 0.05 * rs.randn(24)
    Xs = np.linspace(0.0, 1.1, 9)
    kernels = {
        "double_tanh": g.GibbsKernel1dDoubleTanh(
            initial_params=[1.5, 0.6, 0.3, 0.08, 0.1, 0.05, 0.4, 0.9], param_bounds=[(0, 10)] * 8),
        "cubic_bucket": g.GibbsKernel1dCubicBucket(
            initial_params=[1.5, 0.5, 0.1, 0.3, 0.8, 0.2, 0.15, 0.1], param_bounds=[(0, 10)] * 8),
        "quintic_bucket": g.GibbsKernel1dQuinticBucket(
            initial_params=[1.5, 0.5, 0.1, 0.3, 0.8, 0.2, 0.15, 0.1], param_bounds=[(0, 10)] * 8),
        # GibbsKernel1dExpGauss cannot be run here: exp_gauss_warp slices with len(msb) / 3 (kernel/gibbs.py:833),
        # a float under Python 3, and the reference is loaded unmodified
    }
    for name, k in kernels.items():
        gp = g.GaussianProcess(k)
        gp.add_data(X, y, err_y=0.05)
        gp.add_data(X[::5], -1.0 * np.ones(len(X[::5])), err_y=0.5, n=1)
        gp.add_data(0, 0, n=1)
        out = ll_and_grad(gp, False)
        res = gp.predict(Xs, full_output=True)
        out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
        res1 = gp.predict(Xs, n=1, full_output=True)
        out.update(mean_d1=res1["mean"], std_d1=res1["std"])
        lp = list(k.params[1:])
        out.update(l=k.l_func(gp.X[:, 0], 0, *lp), l1=k.l_func(gp.X[:, 0], 1, *lp))
        save("gibbs_profile_" + name, params=k.params.copy(), **gp_state(gp), **out)


# ---------------------------------------------------------------- input warping (kernel/warping.py:464-720)
def case_warped():
    rs = RandomState(31)
    X = np.sort(rs.rand(20)) * 3.0 + 1.0            # in [1, 4]
    y = np.sin(2 * X) + 0.05 * rs.randn(20)
    kse = g.SquaredExponentialKernel(initial_params=[1.2, 0.3], param_bounds=[(0, 10)] * 2)
    kb = g.BetaWarpedKernel(kse, initial_params=[1.8, 0.7], param_bounds=[(0.01, 10)] * 2)
    k = g.LinearWarpedKernel(kb, [0.5], [4.5])      # map [0.5, 4.5] -> [0, 1], then the beta CDF
    gp = g.GaussianProcess(k)
    gp.add_data(X, y, err_y=0.05)
    gp.add_data(X[::4], 2 * np.cos(2 * X[::4]), err_y=0.1, n=1)
    out = ll_and_grad(gp, False)
    Xs = np.linspace(1.0, 4.0, 7)
    res = gp.predict(Xs, full_output=True)
    out.update(Xs=Xs, mean=res["mean"], std=res["std"], cov=res["cov"])
    res1 = gp.predict(Xs, n=1, full_output=True)
    out.update(mean_d1=res1["mean"], std_d1=res1["std"])
    save("warped_beta_linear_se", params=np.array([float(v) for v in k.params[:]]),
         fixed=np.array([bool(v) for v in k.fixed_params[:]]),
         **gp_state(gp), **out)


if __name__ == "__main__":
    cases = [case_se2d, case_se_pairs, case_matern52, case_matern_generic, case_gibbs, case_c5_full, case_demo,
             case_c3, case_c2, case_noise, case_hyperfd, case_product, case_gibbs_profiles, case_warped,
             case_matern_real_nu, case_hyper_mp, case_composite, case_mcmc_predict, case_ll_matrix]
    only = set(sys.argv[1:])          # e.g. `make_golden.py case_hyperfd` regenerates one family
    for c in cases:
        if not only or c.__name__ in only:
            c()
