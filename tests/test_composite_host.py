"""Kernel algebra (SumKernel / ProductKernel trees, reference kernel/core.py:424-670) as the device evaluates it:
the flattened sum-of-products form, checked WITHOUT a GPU --
  * the oracle's restatement of the reference's derivative-subset enumeration against reference goldens,
  * the device source (csrc/covfn.cuh: comp_eval, compiled for the host) against the same goldens,
  * the host-side flattening of kernel trees and its fall-backs."""
import ctypes
import os
import subprocess
from itertools import product as cartesian
from math import comb

import numpy as np
import pytest

import gptools_b200 as g
from gptools_b200._lib import CompositeId
from helpers import assert_close, load_golden, richardson_fd
from oracle import gp_oracle as orc
from test_covfn_host import lib, pairs, _ptr  # noqa: F401  (fixture + helpers)

SE, M52, MAT, GIBBS = 0, 1, 2, 3
MIXED = ((SE, M52, SE), (3, 3, 3), (0b101, 0b110))      # (SE + Matern52) * SE, 2-D


def comp_pairs(lib, D, structure, params, Xi, Xj, ni, nj, hyper_deriv=-1):
    kids, nps, masks = (np.ascontiguousarray(v, dtype=np.int32) for v in structure)
    Xi = np.ascontiguousarray(Xi, dtype=np.float64)
    Xj = np.ascontiguousarray(Xj, dtype=np.float64)
    ni = np.ascontiguousarray(ni, dtype=np.int32)
    nj = np.ascontiguousarray(nj, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = np.empty(Xi.shape[0])
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    lib.gpt_hostcheck_composite_pairs.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, ip, dp, ctypes.c_int,
                                                  ctypes.c_long, dp, dp, ip, ip, dp]
    rc = lib.gpt_hostcheck_composite_pairs(D, len(kids), _ptr(kids, ctypes.c_int32), _ptr(nps, ctypes.c_int32),
                                           len(masks), _ptr(masks, ctypes.c_int32), _ptr(params, ctypes.c_double),
                                           hyper_deriv, Xi.shape[0], _ptr(Xi, ctypes.c_double),
                                           _ptr(Xj, ctypes.c_double), _ptr(ni, ctypes.c_int32),
                                           _ptr(nj, ctypes.c_int32), _ptr(out, ctypes.c_double))
    assert rc == 0
    return out


class _Struct(object):
    def __init__(self, structure):
        self.structure = structure


# ------------------------------------------------------------------ oracle restatement vs the reference
def test_oracle_composite_matches_reference_goldens():
    gd = load_golden("kernel_algebra_se2d")
    a = (gd["Xi"], gd["Xj"], gd["ni"], gd["nj"])
    p = np.concatenate([gd["params1"], gd["params2"]])
    prod = orc.kernel_pairs(_Struct(((SE, SE), (3, 3), (0b11,))), p, *a)
    assert_close(prod, gd["prod"], rtol=1e-10, atol=1e-12 * np.abs(gd["prod"]).max(), what="product")
    summ = _Struct(((SE, SE), (3, 3), (0b01, 0b10)))
    assert_close(orc.kernel_pairs(summ, p, *a), gd["sum"], rtol=1e-12, atol=1e-14 * np.abs(gd["sum"]).max())
    assert_close(orc.kernel_pairs(summ, p, *a, hyper_deriv=1), gd["sum_hd1"], rtol=1e-9,
                 atol=1e-12 * np.abs(gd["sum_hd1"]).max())
    assert_close(orc.kernel_pairs(summ, p, *a, hyper_deriv=4), gd["sum_hd4"], rtol=1e-9,
                 atol=1e-12 * np.abs(gd["sum_hd4"]).max())
    gm = load_golden("composite_pairs_2d")
    out = orc.kernel_pairs(_Struct(MIXED), gm["params"], gm["Xi"], gm["Xj"], gm["ni"], gm["nj"])
    assert_close(out, gm["K"], rtol=1e-11, atol=1e-13 * np.abs(gm["K"]).max(), what="(SE + Matern52) * SE")


def test_oracle_composite_gp_matches_reference():
    """SE * Matern52 + SE with value and derivative observations: K, ll, alpha of the reference's own GaussianProcess."""
    gd = load_golden("composite_gp_1d")
    st = _Struct(((SE, M52, SE), (2, 2, 2), (0b011, 0b100)))
    r = orc.compute_K_L_alpha_ll(st, gd["params"], gd["X"], gd["n"], gd["y"], gd["err_y"])
    assert_close(r["K"], gd["K"], rtol=1e-11, atol=1e-13 * np.abs(gd["K"]).max(), what="K")
    assert_close(r["ll"], float(gd["ll"]) - float(gd["log_prior"]), rtol=1e-10, what="ll")
    assert_close(np.ravel(r["alpha"]), gd["alpha"], rtol=1e-7, atol=1e-9 * np.abs(gd["alpha"]).max(), what="alpha")


def test_oracle_composite_gradient_matches_reference():
    """SE + SE in 2-D with derivative observations: ll and the reference's own ll gradient (SumKernel hyper_deriv,
    kernel/core.py:576-582) -- pins the oracle's composite hyper-derivative path that the GPU tests compare against."""
    gd = load_golden("composite_sum_grad_2d")
    st = _Struct(((SE, SE), (3, 3), (0b01, 0b10)))
    r = orc.compute_K_L_alpha_ll(st, gd["params"], gd["X"], gd["n"], gd["y"], gd["err_y"], grad_idx=list(range(6)))
    assert_close(r["ll"], float(gd["ll"]) - float(gd["log_prior"]), rtol=1e-10, what="ll")
    assert_close(r["ll_deriv"], gd["ll_deriv"], rtol=1e-8, atol=1e-10 * np.abs(gd["ll_deriv"]).max(), what="ll gradient")


# ------------------------------------------------------------------ the device source, compiled for the host
def test_device_composite_source_matches_reference_goldens(lib):
    gd = load_golden("kernel_algebra_se2d")
    a = (gd["Xi"], gd["Xj"], gd["ni"], gd["nj"])          # orders up to (2, 1) per dimension
    p = np.concatenate([gd["params1"], gd["params2"]])
    prod = comp_pairs(lib, 2, ((SE, SE), (3, 3), (0b11,)), p, *a)
    assert_close(prod, gd["prod"], rtol=1e-10, atol=1e-12 * np.abs(gd["prod"]).max(), what="product")
    summ = ((SE, SE), (3, 3), (0b01, 0b10))
    assert_close(comp_pairs(lib, 2, summ, p, *a), gd["sum"], rtol=1e-12, atol=1e-14 * np.abs(gd["sum"]).max())
    assert_close(comp_pairs(lib, 2, summ, p, *a, hyper_deriv=1), gd["sum_hd1"], rtol=1e-9,
                 atol=1e-12 * np.abs(gd["sum_hd1"]).max())
    assert_close(comp_pairs(lib, 2, summ, p, *a, hyper_deriv=4), gd["sum_hd4"], rtol=1e-9,
                 atol=1e-12 * np.abs(gd["sum_hd4"]).max())
    gm = load_golden("composite_pairs_2d")
    out = comp_pairs(lib, 2, MIXED, gm["params"], gm["Xi"], gm["Xj"], gm["ni"], gm["nj"])
    assert_close(out, gm["K"], rtol=1e-10, atol=1e-12 * np.abs(gm["K"]).max(), what="(SE + Matern52) * SE")


def test_device_composite_gp_matrix_matches_reference(lib):
    gd = load_golden("composite_gp_1d")
    X, n = gd["X"], gd["n"]
    M = len(X)
    K = comp_pairs(lib, 1, ((SE, M52, SE), (2, 2, 2), (0b011, 0b100)), gd["params"], np.repeat(X, M, axis=0),
                   np.tile(X, (M, 1)), np.repeat(n, M, axis=0), np.tile(n, (M, 1))).reshape(M, M)
    assert_close(K, gd["K"], rtol=1e-10, atol=1e-12 * np.abs(gd["K"]).max(), what="K")


def _leibniz(lib, D, leaves, params_list, hd_leaf, hd_local, Xi, Xj, ni, nj):
    """Independent Python evaluation of a product of leaves with the general Leibniz rule, the hyper-derivative taken
    in leaf `hd_leaf` -- the leaves themselves through the (separately pinned) single-kernel host build."""
    if len(leaves) == 1:
        return pairs(lib, leaves[0], params_list[0], Xi, Xj, ni, nj, hd_local if hd_leaf == 0 else -1)
    out = np.zeros(len(Xi))
    nij = np.hstack((ni, nj))
    for row in np.unique(nij, axis=0):
        sel = (nij == row).all(axis=1)
        cnt = int(sel.sum())
        for a in cartesian(*[range(int(m) + 1) for m in row]):
            a = np.array(a, dtype=int)
            w = np.prod([comb(int(m), int(x)) for m, x in zip(row, a)])
            n1, n2 = np.tile(a, (cnt, 1)), np.tile(row - a, (cnt, 1))
            head = pairs(lib, leaves[0], params_list[0], Xi[sel], Xj[sel], n1[:, :D], n1[:, D:],
                         hd_local if hd_leaf == 0 else -1)
            tail = _leibniz(lib, D, leaves[1:], params_list[1:], hd_leaf - 1, hd_local, Xi[sel], Xj[sel], n2[:, :D],
                            n2[:, D:])
            out[sel] += w * head * tail
    return out


@pytest.mark.parametrize("nfac", [2, 3, 4])
def test_device_composite_product_hyper_derivatives(lib, nfac):
    """hyper_deriv inside products (the reference raises NotImplementedError there): the device's odometer over
    derivative splits against an independent recursive Leibniz sum, for 2, 3 and 4 factors and every parameter."""
    rs = np.random.RandomState(5 + nfac)
    D = 2 if nfac < 4 else 1                  # at most 10 parameters in total
    npar = D + 1
    leaves = [SE, M52, SE, SE][:nfac]
    plist = [np.array([1.0 + 0.1 * q, 0.5 + 0.2 * q, 0.9 - 0.1 * q][:npar]) for q in range(nfac)]
    Mp = 40
    Xi, Xj = rs.rand(Mp, D), rs.rand(Mp, D)
    Xj[:4] = Xi[:4]
    ni, nj = np.zeros((Mp, D), dtype=int), np.zeros((Mp, D), dtype=int)
    ni[np.arange(Mp), rs.randint(0, D, Mp)] = rs.randint(0, 2, Mp)
    nj[np.arange(Mp), rs.randint(0, D, Mp)] = rs.randint(0, 2, Mp)
    structure = (leaves, [npar] * nfac, [(1 << nfac) - 1])
    params = np.concatenate(plist)
    want = _leibniz(lib, D, leaves, plist, -1, -1, Xi, Xj, ni, nj)
    got = comp_pairs(lib, D, structure, params, Xi, Xj, ni, nj)
    assert_close(got, want, rtol=1e-12, atol=1e-13 * np.abs(want).max(), what="value")
    for hd in range(npar * nfac):
        want = _leibniz(lib, D, leaves, plist, hd // npar, hd % npar, Xi, Xj, ni, nj)
        got = comp_pairs(lib, D, structure, params, Xi, Xj, ni, nj, hyper_deriv=hd)
        assert_close(got, want, rtol=1e-12, atol=1e-13 * np.abs(want).max(), what="hyper_deriv %d" % hd)
    # and against central differences of the value
    f = lambda th: comp_pairs(lib, D, structure, th, Xi, Xj, ni, nj)
    for hd in (0, 4):
        fd = richardson_fd(f, params, hd, 1e-3)
        got = comp_pairs(lib, D, structure, params, Xi, Xj, ni, nj, hyper_deriv=hd)
        assert_close(got, fd, rtol=1e-7, atol=1e-8 * np.abs(fd).max(), what="FD %d" % hd)


# ------------------------------------------------------------------ host-side flattening
def _se(p, D=1, **kw):
    return g.SquaredExponentialKernel(num_dim=D, initial_params=p, param_bounds=[(0, 10)] * (D + 1), **kw)


def test_kernel_trees_flatten_to_sum_of_products():
    a, b, c = _se([1.0, 0.5]), g.Matern52Kernel(initial_params=[0.8, 0.4], param_bounds=[(0, 10)] * 2), _se([0.3, 0.2])
    kid, params = (a * b + c).device_descriptor()
    assert isinstance(kid, CompositeId) and int(kid) == 5
    assert kid.structure == ((SE, M52, SE), (2, 2, 2), (0b011, 0b100))
    assert list(params) == [1.0, 0.5, 0.8, 0.4, 0.3, 0.2]
    kid2, _ = ((a + b) * c).device_descriptor()
    assert kid2.structure == ((SE, M52, SE), (2, 2, 2), (0b101, 0b110))
    assert kid != kid2 and kid == (a * b + c).device_descriptor()[0] and kid == 5
    import copy, pickle
    assert pickle.loads(pickle.dumps(kid)) == kid and copy.deepcopy(kid2).structure == kid2.structure
    d = _se([0.7, 0.9])
    kid3, _ = ((a + b) * (c + d)).device_descriptor()
    assert kid3.structure[2] == (0b0101, 0b1001, 0b0110, 0b1010)
    # parameter updates reach the descriptor
    k = a + c
    k.params = [2.0, 0.6, 0.4, 0.1]
    assert list(k.device_descriptor()[1]) == [2.0, 0.6, 0.4, 0.1]


def test_kernel_trees_outside_the_device_limits_fall_back_to_the_host():
    ks = [_se([1.0 + 0.1 * i, 0.5]) for i in range(6)]
    five = ks[0] + ks[1] + ks[2] + ks[3] + ks[4]
    assert five.device_descriptor() is None                      # more than 4 leaves

    class HostK(g.Kernel):
        def __init__(self):
            super(HostK, self).__init__(num_dim=1, num_params=1, initial_params=[1.0], param_bounds=[(0, 10)])

        def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
            return np.ones(np.atleast_2d(Xi).shape[0])

    assert (ks[0] + HostK()).device_descriptor() is None         # user-defined operand
    big = g.MaternKernel(num_dim=3, initial_params=[1, 2.5, 1, 1, 1], param_bounds=[(0, 10)] * 5)
    big2 = g.MaternKernel(num_dim=3, initial_params=[1, 1.5, 1, 1, 1], param_bounds=[(0, 10)] * 5)
    big3 = g.SquaredExponentialKernel(num_dim=3, initial_params=[1, 1, 1, 1], param_bounds=[(0, 10)] * 4)
    assert (big + big2).device_descriptor() is not None          # 10 parameters: the limit
    assert (big + big2 + big3).device_descriptor() is None       # 14 parameters
    aux = g.GibbsKernel1dDoubleTanh(initial_params=[1.5, 0.6, 0.3, 0.08, 0.1, 0.05, 0.4, 0.9], param_bounds=[(0, 10)] * 8)
    assert (aux + ks[0]).device_descriptor() is None             # host-evaluated length-scale profile


def test_composite_hyper_deriv_checks_follow_the_leaves():
    a = _se([1.0, 0.5])
    m = g.MaternKernel(initial_params=[1.0, 2.5, 0.7], param_bounds=[(0, 10)] * 3)
    k = a * m
    k.check_hyper_deriv([0, 1, 2, 3, 4])
    assert k.fd_hyper_idxs == (3,)                               # nu of the Matern operand: finite differences
    assert not k.batchable(True) and k.batchable(False)
    rows = np.array([[1.0, 0.5, 1.0, 2.5, 0.7], [1.0, 0.5, 1.0, -1.0, 0.7]])
    assert list(k.batch_rows_supported(rows)) == [True, False]


# ------------------------------------------------------------------ GaussianProcess host logic on a composite kernel
def test_gaussian_process_host_logic_with_composite_kernel():
    """The GP drives a composite kernel like any device kernel: one descriptor, device mode, batched entry, gradients
    in the order of the concatenated parameter vector (oracle-backed test double instead of the GPU)."""
    from fake_device import FakeDevice
    gd = load_golden("composite_gp_1d")
    p = gd["params"]
    k = (_se(p[0:2]) * g.Matern52Kernel(initial_params=list(p[2:4]), param_bounds=[(0, 10)] * 2) + _se(p[4:6]))
    gp = g.GaussianProcess(k)
    gp._dev_obj = FakeDevice()
    nv = int((gd["n"][:, 0] == 0).sum())
    gp.add_data(gd["X"][:nv], gd["y"][:nv], err_y=gd["err_y"][:nv])
    gp.add_data(gd["X"][nv:], gd["y"][nv:], err_y=gd["err_y"][nv:], n=1)
    assert gp._device_mode() and gp._batchable(False)
    gp.compute_K_L_alpha_ll()
    assert_close(gp.ll, float(gd["ll"]), rtol=1e-9, what="ll")
    mean, std = gp.predict(gd["Xs"])
    assert_close(mean, gd["mean"], rtol=1e-8, atol=1e-9)
    assert_close(std, gd["std"], rtol=1e-5, atol=1e-8)
    th = np.array(gp.free_params[:], dtype=float)
    f = gp.update_hyperparameters_batch(np.vstack([th, 1.1 * th]), with_deriv=False)
    assert_close(f[0], -float(gd["ll"]), rtol=1e-9)
    assert gp._dev_obj.calls.count("ll_batched") == 1
    # an SE-only tree has every hyper-derivative in the oracle: gradient plumbing through the GP
    k2 = _se([1.1, 0.9]) * _se([0.8, 0.5]) + _se([0.4, 0.15], fixed_params=[False, True])
    gp2 = g.GaussianProcess(k2, use_hyper_deriv=True)
    gp2._dev_obj = FakeDevice()
    gp2.add_data(gd["X"][:nv], gd["y"][:nv], err_y=gd["err_y"][:nv])
    th2 = np.array(gp2.free_params[:], dtype=float)
    assert len(th2) == 5
    with pytest.warns(UserWarning):
        f0, df0 = gp2.update_hyperparameters(th2)
    for i in range(5):
        fd = richardson_fd(lambda t: gp2.update_hyperparameters(t)[0], th2, i, 1e-4)
        assert abs(df0[i] - fd) <= 1e-6 * np.abs(df0).max(), (i, df0[i], fd)
