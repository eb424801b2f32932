"""Input warping: k(w(x), w(x')) for an elementwise warping function w (kernel/warping.py).

Host-side composition, like the reference: the warped coordinates are handed to the wrapped kernel's own call
(``gpt_cov_pairs`` for the accelerated kernels) and first-derivative observations pick up the chain-rule factor
w'(x) (kernel/warping.py:491-505).  A GaussianProcess built on a WarpedKernel takes the host-kernel path
(``gpt_ll_from_K`` / ``gpt_predict_from_Kstar``): K and K* are assembled by calling this class on the pair lists,
factorisation and solves run on the device.
"""
import numpy as np
import scipy.special

from .._params import ParamHolder, count_fun_params
from ..error_handling import GPArgumentError
from ..utils import LogNormalJointPrior
from .core import BinaryKernel, Kernel

__all__ = ["WarpingFunction", "beta_cdf_warp", "linear_warp", "WarpedKernel", "BetaWarpedKernel",
           "LinearWarpedKernel"]


class WarpingFunction(ParamHolder):
    """``fun(X, d, n, *params)``: warp the coordinates ``X`` of dimension ``d``, derivative order ``n``
    (kernel/warping.py:63-313).  Holds its parameters like a kernel does."""

    def __init__(self, fun, num_dim=1, num_params=None, initial_params=None, fixed_params=None, param_bounds=None,
                 param_names=None, enforce_bounds=False, hyperprior=None):
        self.fun = fun
        self.num_dim = int(num_dim)
        if num_params is None:
            num_params = count_fun_params(fun, 3)   # (X, d, n) come first
            if num_params is None:
                if initial_params is None:
                    raise GPArgumentError("Warping functions taking *args need num_params or initial_params")
                num_params = len(initial_params)
        self._init_params(num_params, initial_params, fixed_params, param_bounds, param_names, enforce_bounds,
                          hyperprior, arg_error=GPArgumentError)

    def __call__(self, X, d, n):
        return self.fun(X, d, n, *self.params)


def beta_cdf_warp(X, d, n, *args):
    r"""w(x) = I_x(a_d, b_d), the regularised incomplete beta function, on [0, 1]; parameters a_0, b_0, a_1, b_1, ...
    (kernel/warping.py:315-364; Snoek et al., ICML 2014).  n = 1 is the beta density."""
    X = np.asarray(X, dtype=float)
    a, b = args[2 * d], args[2 * d + 1]
    if n == 0:
        return scipy.special.betainc(a, b, X)
    if n == 1:
        return (1.0 - X) ** (b - 1.0) * X ** (a - 1.0) / scipy.special.beta(a, b)
    raise NotImplementedError("Only derivatives up to order 1 are supported!")


def linear_warp(X, d, n, *args):
    r"""w(x) = (x - a_d) / (b_d - a_d); parameters a_0, b_0, a_1, b_1, ... (kernel/warping.py:367-401)."""
    X = np.asarray(X, dtype=float)
    a, b = args[2 * d], args[2 * d + 1]
    if n == 0:
        return (X - a) / (b - a)
    if n == 1:
        return np.ones_like(X) / (b - a)
    return np.zeros_like(X)


class WarpedKernel(BinaryKernel):
    """k(w_1(x_1), ..., w_D(x_D); same for x') (kernel/warping.py:464-631); parameters = kernel's, then the
    warping function's.  Derivative orders up to one per point."""

    def __init__(self, k, w):
        if not isinstance(k, Kernel):
            raise TypeError("k must be a Kernel")
        if not isinstance(w, WarpingFunction):
            w = WarpingFunction(w, num_dim=k.num_dim)
        if k.num_dim != w.num_dim:
            raise ValueError("k and w must have the same number of dimensions!")
        # BinaryKernel's parameter plumbing works on the pair (k1, k2) = (kernel, warping function)
        self.k1 = self.k = k
        self.k2 = self.w = w
        self.num_dim = k.num_dim

    def w_func(self, X, d, n):
        """The (possibly nested) warping and its first derivative (kernel/warping.py:507-531)."""
        if n == 0:
            wX = self.w(X, d, 0)
            return self.k.w_func(wX, d, 0) if isinstance(self.k, WarpedKernel) else wX
        if n == 1:
            out = self.w(X, d, 1)
            if isinstance(self.k, WarpedKernel):
                out = out * self.k.w_func(self.w(X, d, 0), d, 1)
            return out
        raise ValueError("Derivative orders greater than one are not supported!")

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        Xi = np.atleast_2d(np.asarray(Xi, dtype=float))
        Xj = np.atleast_2d(np.asarray(Xj, dtype=float))
        ni = np.atleast_2d(np.asarray(ni, dtype=int))
        nj = np.atleast_2d(np.asarray(nj, dtype=int))
        if (ni > 1).any() or (nj > 1).any():
            raise ValueError("Derivative orders greater than one are not supported!")
        if hyper_deriv is not None and hyper_deriv >= self.k.num_params:
            raise NotImplementedError("Derivatives with respect to the warping parameters are not available")
        wXi = np.column_stack([self.w(Xi[:, d], d, 0) for d in range(self.num_dim)])
        wXj = np.column_stack([self.w(Xj[:, d], d, 0) for d in range(self.num_dim)])
        out = np.array(self.k(wXi, wXj, ni, nj, hyper_deriv=hyper_deriv, symmetric=symmetric), dtype=float)
        for d in range(self.num_dim):
            mi, mj = ni[:, d] == 1, nj[:, d] == 1
            out[mi] *= self.w(Xi[mi, d], d, 1)
            out[mj] *= self.w(Xj[mj, d], d, 1)
        return out


class BetaWarpedKernel(WarpedKernel):
    """Warp with the beta CDF, inputs in the unit hypercube (kernel/warping.py:633-676).  Without bounds or a
    hyperprior every alpha, beta gets the log-normal(0, 0.5) prior of the reference."""

    def __init__(self, k, **w_kwargs):
        names = []
        for d in range(k.num_dim):
            names += ['\\alpha_{:d}'.format(d), '\\beta_{:d}'.format(d)]
        if 'hyperprior' not in w_kwargs and 'param_bounds' not in w_kwargs:
            w_kwargs['hyperprior'] = LogNormalJointPrior([0, 0] * k.num_dim, [0.5, 0.5] * k.num_dim)
        w = WarpingFunction(beta_cdf_warp, num_dim=k.num_dim, num_params=2 * k.num_dim, param_names=names, **w_kwargs)
        super(BetaWarpedKernel, self).__init__(k, w)


class LinearWarpedKernel(WarpedKernel):
    """Warp with w(x) = (x - a) / (b - a), a and b fixed (kernel/warping.py:678-720): maps a box onto the unit
    hypercube, e.g. ahead of a BetaWarpedKernel."""

    def __init__(self, k, a, b):
        a = np.atleast_1d(np.asarray(a, dtype=float))
        b = np.atleast_1d(np.asarray(b, dtype=float))
        if len(a) != k.num_dim or len(b) != k.num_dim:
            raise ValueError("a and b must have length equal to k.num_dim!")
        params, names, bounds = [], [], []
        for d in range(k.num_dim):
            params += [a[d], b[d]]
            names += ['a_{:d}'.format(d), 'b_{:d}'.format(d)]
            # the reference brackets the fixed values by +-1e-3 (kernel/warping.py:707): the (constant) prior density
            # of these fixed parameters enters ll, so the same bounds are needed for identical ll values
            bounds += [(a[d] - 1e-3, a[d] + 1e-3), (b[d] - 1e-3, b[d] + 1e-3)]
        w = WarpingFunction(linear_warp, num_dim=k.num_dim, num_params=2 * k.num_dim, initial_params=params,
                            fixed_params=np.ones(2 * k.num_dim, dtype=bool), param_names=names, param_bounds=bounds)
        super(LinearWarpedKernel, self).__init__(k, w)
