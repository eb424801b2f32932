"""Config 4 (large single GP: 2-D SE kernel, value + both gradient components at every location) at sizes the
reference itself cannot reach (its pair-list temporaries are ~155 GB at M = 49152).

* M = 24576: the device path against the oracle restatement -- covariance block-assembled with the oracle's kernel
  (oracle/gp_oracle.py, the numpy restatement pinned to the reference's goldens), factored with
  scipy.linalg.cholesky: ll, alpha, and predictive mean / std on a 1000-point slice at 1e-9.
* M = 49152 (the full configuration): 10^6 test points, mean + std; size-independent properties (chunking independence,
  fused mean-only path == mean of the mean + std path, 0 <= std <= sigma_f, residual K_tot alpha = y on sampled rows).
"""
import concurrent.futures
import os

import numpy as np
import pytest
import scipy.linalg

import gptools_b200 as g
from helpers import assert_close
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

EPS = np.finfo(float).eps


def c4_data(nloc, seed=0):
    rs = np.random.RandomState(seed)
    X0 = rs.rand(nloc, 2)
    X = np.vstack([X0, X0, X0])
    n = np.vstack([np.zeros((nloc, 2), dtype=int), np.tile([1, 0], (nloc, 1)), np.tile([0, 1], (nloc, 1))])
    y = np.concatenate([np.sin(3 * X0[:, 0]) * np.cos(2 * X0[:, 1]),
                        3 * np.cos(3 * X0[:, 0]) * np.cos(2 * X0[:, 1]),
                        -2 * np.sin(3 * X0[:, 0]) * np.sin(2 * X0[:, 1])]) + 0.05 * rs.randn(3 * nloc)
    return X, n, y


def oracle_K_lower(params, X, n, block=512):
    """Lower triangle (row blocks against columns [0, r1)) of K(X, X) from the oracle's kernel, row blocks in
    parallel threads (numpy releases the GIL in its ufuncs)."""
    M = X.shape[0]
    K = np.zeros((M, M))

    def work(r0):
        r1 = min(r0 + block, M)
        K[r0:r1, :r1] = orc.compute_Kij(orc.KERNEL_SE, params, X[r0:r1], X[:r1], n[r0:r1], n[:r1])
        return r0
    nthreads = max(1, len(os.sched_getaffinity(0)))
    with concurrent.futures.ThreadPoolExecutor(nthreads) as ex:
        list(ex.map(work, range(0, M, block)))
    return K


def test_c4_shape_M24576_against_oracle():
    nloc = 8192
    X, n, y = c4_data(nloc)
    M = 3 * nloc
    params = np.array([1.0, 0.05, 0.05])
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=params, param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=y, err_y=0.05, n=n)
    gp.compute_K_L_alpha_ll()
    ll_dev = gp.ll - gp.hyperprior(gp.params)
    alpha_dev = gp.alpha.ravel().copy()
    Xs = np.random.RandomState(2).rand(1000, 2)
    mean_dev, std_dev = gp.predict(Xs)
    dmean_dev, dstd_dev = gp.predict(Xs[:200], n=np.tile([1, 0], (200, 1)))

    # oracle: same formulas as gaussian_process.py:1418-1469 on the block-assembled covariance
    K = oracle_K_lower(params, X, n)
    K[np.diag_indices(M)] += 0.05 ** 2 + 1e2 * EPS
    L = scipy.linalg.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    alpha = scipy.linalg.cho_solve((L, True), y, check_finite=False)
    ll = -0.5 * y.dot(alpha) - np.log(np.diag(L)).sum() - 0.5 * M * np.log(2.0 * np.pi)
    assert_close(ll_dev, ll, rtol=1e-9, what="ll at M = 24576")
    assert_close(alpha_dev, alpha, rtol=0.0, atol=1e-9 * np.abs(alpha).max(), what="alpha at M = 24576")
    # the device's L, a few rows
    Ldev = gp._dev().get_L()
    for r in (0, 1, 8191, 8192, 20000, M - 1):
        assert_close(Ldev[r, :r + 1], L[r, :r + 1], rtol=0.0, atol=1e-10 * params[0], what="L row %d" % r)
    del Ldev

    def oracle_predict(Xp, np_):
        Ks = orc.compute_Kij(orc.KERNEL_SE, params, X, Xp, n, np_)           # (M, M*), gaussian_process.py:966
        mean = Ks.T.dot(alpha)
        v = scipy.linalg.solve_triangular(L, Ks, lower=True, check_finite=False)
        kss = np.diag(orc.compute_Kij(orc.KERNEL_SE, params, Xp, None, np_, None))
        return mean, kss - np.sum(v * v, axis=0), kss

    m_o, var_o, kss = oracle_predict(Xs, np.zeros((1000, 2), dtype=int))
    assert_close(mean_dev, m_o, rtol=1e-9, atol=1e-9, what="predictive mean")
    assert np.all(np.abs(std_dev ** 2 - var_o) <= 1e-9 * kss), np.abs(std_dev ** 2 - var_o).max()
    m_o, var_o, kss = oracle_predict(Xs[:200], np.tile([1, 0], (200, 1)))
    assert_close(dmean_dev, m_o, rtol=1e-9, atol=1e-9 * np.abs(m_o).max(), what="predictive mean of d/dx1")
    assert np.all(np.abs(dstd_dev ** 2 - var_o) <= 1e-9 * kss)


def test_c4_full_size_million_point_prediction():
    nloc = 16384
    X, n, y = c4_data(nloc)
    M = 3 * nloc
    params = np.array([1.0, 0.05, 0.05])
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=params, param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=y, err_y=0.05, n=n)
    gp.compute_K_L_alpha_ll()
    assert np.isfinite(gp.ll)
    alpha = gp.alpha.ravel()
    # residual K_tot alpha = y on sampled rows, rows of K from the oracle's kernel
    rows = np.array([0, 1, 4097, 16383, 16384, 30000, 49151])
    Krows = orc.compute_Kij(orc.KERNEL_SE, params, X[rows], X, n[rows], n)
    resid = Krows.dot(alpha) + (0.05 ** 2 + 1e2 * EPS) * alpha[rows] - y[rows]
    assert np.abs(resid).max() <= 1e-8 * np.abs(y).max(), resid

    Ms = 10 ** 6
    Xs = np.random.RandomState(2).rand(Ms, 2)
    mean, std = gp.predict(Xs)
    assert mean.shape == (Ms,) and std.shape == (Ms,)
    assert np.isfinite(mean).all() and np.isfinite(std).all()
    assert np.all(std >= 0) and np.all(std <= params[0] * (1 + 1e-12))
    f = np.sin(3 * Xs[:, 0]) * np.cos(2 * Xs[:, 1])
    inside = np.all((Xs > 0.05) & (Xs < 0.95), axis=1)
    assert np.abs(mean - f)[inside].max() < 0.05           # 16384 noisy values + gradients pin the surface
    # fused mean-only path (K* never stored) == mean of the mean + std path
    mean_only = gp.predict(Xs, return_std=False)
    assert_close(mean_only, mean, rtol=1e-9, atol=1e-10, what="fused mean-only path")
    # chunking independence: a strided subset predicted on its own
    sl = slice(0, Ms, 997)
    m2, s2 = gp.predict(Xs[sl])
    assert_close(m2, mean[sl], rtol=1e-12, atol=1e-13, what="chunking independence (mean)")
    assert_close(s2, std[sl], rtol=1e-9, atol=1e-12, what="chunking independence (std)")
    # the subset against the oracle's formulas with the DEVICE alpha (mean) -- K* columns from the oracle's kernel
    sub = Xs[sl][:64]
    Ks = orc.compute_Kij(orc.KERNEL_SE, params, X, sub, n, np.zeros((64, 2), dtype=int))
    assert_close(m2[:64], Ks.T.dot(alpha), rtol=1e-9, atol=1e-10, what="mean against oracle K* with device alpha")
