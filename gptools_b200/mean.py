"""Parametric mean functions (host side; reference: gptools/mean.py).

The mean function enters the likelihood path only as O(M) vectors: ``y - T mu(X, n)``
(gaussian_process.py:1455-1461), ``mu(Xstar, n)`` added to the predictive mean (:972-974) and
``dmu/dtheta . alpha`` in the gradient (:1507-1514).  It is evaluated in numpy and handed to the device
as the right-hand side."""
import numpy as np

from ._params import ParamHolder, count_fun_params
from .utils import unique_rows

__all__ = ["MeanFunction", "constant", "ConstantMeanFunction", "linear", "LinearMeanFunction"]


class MeanFunction(ParamHolder):
    """Wrap ``fun(X, n, p1, p2, ..., hyper_deriv=None)`` as a mean function (mean.py:29-170).

    ``fun`` receives the rows of ``X`` that share one derivative-order vector ``n`` (a 1-D array of
    length D) and returns the mean (or its derivative with respect to parameter ``hyper_deriv``)."""

    def __init__(self, fun, num_params=None, initial_params=None, fixed_params=None, param_bounds=None,
                 param_names=None, enforce_bounds=False, hyperprior=None):
        self.fun = fun
        if num_params is None:
            num_params = count_fun_params(fun, 2)
            if num_params is None:
                if hyperprior is not None:
                    num_params = len(hyperprior.bounds)
                elif param_names is not None:
                    num_params = len(param_names)
                elif param_bounds is not None:
                    num_params = len(param_bounds)
                else:
                    raise ValueError("If the mean function uses a variable number of arguments, you must also "
                                     "specify an explicit hyperprior, list of param_names and/or list of param_bounds.")
        self._init_params(num_params, initial_params, fixed_params, param_bounds, param_names, enforce_bounds,
                          hyperprior, arg_error=ValueError)

    def __call__(self, X, n, hyper_deriv=None):
        n = np.atleast_2d(np.asarray(n, dtype=int))
        X = np.atleast_2d(np.asarray(X))
        mu = np.zeros(X.shape[0])
        for nn in unique_rows(n):
            idxs = (n == nn).all(axis=1)
            mu[idxs] = self.fun(X[idxs, :], nn, *self.params, hyper_deriv=hyper_deriv)
        return mu


def constant(X, n, mu, hyper_deriv=None):
    """Constant mean (mean.py:286-309): value mu for n == 0, zero for any derivative."""
    if (n == 0).all():
        if hyper_deriv is not None:
            return np.ones(X.shape[0])
        return mu * np.ones(X.shape[0])
    return np.zeros(X.shape[0])


class ConstantMeanFunction(MeanFunction):
    def __init__(self, **kwargs):
        if 'hyperprior' not in kwargs and 'param_bounds' not in kwargs:
            kwargs['param_bounds'] = [(-1e3, 1e3)]
        super(ConstantMeanFunction, self).__init__(constant, param_names=['\\mu'], **kwargs)


def linear(X, n, *args, **kwargs):
    """Linear mean ``slopes . x + offset`` of arbitrary dimension (same contract as mean.py:443-475):
    ``args = (m_1, ..., m_D, b)``; first derivatives return the slope, higher ones zero."""
    hyper_deriv = kwargs.pop('hyper_deriv', None)
    slopes = np.asarray(args[:-1], dtype=float)
    offset = args[-1]
    order = int(np.sum(n))
    npts = X.shape[0]
    if order > 1:
        return np.zeros(npts)
    if hyper_deriv is None:
        if order == 0:
            return (slopes * X).sum(axis=1) + offset
        return np.full(npts, slopes[np.asarray(n) == 1][0])
    # derivative with respect to parameter number hyper_deriv
    if order == 0:
        if hyper_deriv < len(slopes):
            return np.array(X[:, hyper_deriv], dtype=float)
        if hyper_deriv == len(slopes):
            return np.ones(npts)
        raise ValueError("Invalid value for hyper_deriv, " + str(hyper_deriv))
    hit = hyper_deriv < len(n) and n[hyper_deriv] == 1
    return np.ones(npts) if hit else np.zeros(npts)


class LinearMeanFunction(MeanFunction):
    def __init__(self, num_dim=1, **kwargs):
        names = ['m_{:d}'.format(i + 1) for i in range(num_dim)] + ['b']
        if 'hyperprior' not in kwargs and 'param_bounds' not in kwargs:
            kwargs['param_bounds'] = [(-1e3, 1e3)] * (num_dim + 1)
        super(LinearMeanFunction, self).__init__(linear, num_params=num_dim + 1, param_names=names, **kwargs)
