"""Kernel base class and the device-dispatch plumbing.

``Kernel`` keeps the reference's constructor, parameter views and call signature
(kernel/core.py:44-421).  A kernel whose closed form exists in the CUDA library exposes a *device
descriptor* ``(kernel_id, params)``; GaussianProcess then runs assembly, factorisation, gradient and
prediction on the device without ever calling the Python ``__call__``.  ``__call__`` itself (the
flattened pair-list contract, kernel/core.py:220-257) is also served by the device
(``gpt_cov_pairs``), so user code and tests that call kernels directly keep working.  There is no
CPU evaluation of the accelerated kernels anywhere in the package.

User-defined kernels simply subclass ``Kernel`` and implement ``__call__`` in numpy; they have no
descriptor and GaussianProcess assembles K by calling them on the pair lists exactly like the reference
(gaussian_process.py:1591-1602), then ships K to the device for the factorisation.
"""
import numpy as np

from .._params import ParamHolder
from ..error_handling import GPArgumentError
from ..utils import CombinedBounds

__all__ = ["Kernel", "DeviceKernel", "BinaryKernel", "SumKernel", "ProductKernel"]

_default_device = None


def default_device():
    """Process-wide Device used by direct kernel calls (lazy; one per process)."""
    global _default_device
    if _default_device is None:
        from .._lib import Device
        _default_device = Device()
    return _default_device


class Kernel(ParamHolder):
    """Covariance kernel base class (kernel/core.py:44-218 for the parameter semantics)."""

    kernel_id = None  # set by device-accelerated subclasses

    def __init__(self, num_dim=1, num_params=0, initial_params=None, fixed_params=None, param_bounds=None,
                 param_names=None, enforce_bounds=False, hyperprior=None):
        if num_dim < 1 or not isinstance(num_dim, (int, np.integer)):
            raise ValueError("num_dim must be an integer > 0!")
        self.num_dim = int(num_dim)
        self._init_params(num_params, initial_params, fixed_params, param_bounds, param_names, enforce_bounds,
                          hyperprior, warn_default_bounds=True, arg_error=GPArgumentError)

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        """Covariance of d^ni f(Xi) and d^nj f(Xj) for M flattened pairs -> (M,) (kernel/core.py:220-257)."""
        raise NotImplementedError("This is an abstract method -- please use one of the implementing subclasses!")

    def device_descriptor(self):
        """(kernel_id, params) when the closed form lives in the CUDA library, else None."""
        return None

    def __add__(self, other):
        return SumKernel(self, other)

    def __mul__(self, other):
        return ProductKernel(self, other)

    def _compute_r2l2(self, tau, return_l=False):
        """sum_d tau_d^2 / l_d^2 with 0/0 -> 0 (kernel/core.py:384-421); a host helper for user-defined
        kernels -- the accelerated kernels compute this inside their device functions."""
        l_mat = np.tile(self.params[-self.num_dim:], (tau.shape[0], 1))
        with np.errstate(divide="ignore", invalid="ignore"):
            tau_over_l = tau / l_mat
        tau_over_l[(tau == 0) & (l_mat == 0)] = 0.0
        r2l2 = np.sum(tau_over_l ** 2, axis=1)
        return (r2l2, l_mat) if return_l else r2l2


def richardson_difference(f, params, i, h):
    """(4 D(h/2) - D(h)) / 3 with D the central difference of ``f`` along ``params[i]``: error O(h^4)."""
    def D(hh):
        pp, pm = np.array(params, dtype=float), np.array(params, dtype=float)
        pp[i] += hh
        pm[i] -= hh
        return (np.asarray(f(pp), dtype=float) - np.asarray(f(pm), dtype=float)) / (2.0 * hh)
    return (4.0 * D(0.5 * h) - D(h)) / 3.0


class DeviceKernel(Kernel):
    """A kernel evaluated by libgptb200 (``kernel_id`` names the device function in csrc/covfn.cuh)."""

    supports_hyper_deriv = False

    #: parameter indices whose hyper-derivative has no closed form on the device and is taken by Richardson-extrapolated
    #: central differences of the device's own evaluation instead (MaternKernel: the order nu)
    fd_hyper_idxs = ()
    #: relative step of those differences
    FD_HYPER_REL_STEP = 2e-3

    def device_descriptor(self):
        self._check_params_for_device()
        return (self.kernel_id, np.array(self.params, dtype=float))

    def _check_params_for_device(self):
        pass

    def _check_orders(self, ni, nj):
        pass

    def device_points(self, X, n):
        """The point / derivative-order arrays the device function consumes.  Identity for most kernels; kernels
        whose closed form takes per-point auxiliary columns (GibbsKernel1d: l(x), l'(x)) append them here."""
        return X, n

    def device_points_key(self):
        """Changes whenever ``device_points`` would return different auxiliary columns (None: never)."""
        return None

    #: gradient slots of the batched device kernel for kernels other than SE (csrc/batched4.cu: 1 + GPT_MAX_DIM)
    BATCHED_GRAD_SLOTS = 7

    def batchable(self, with_deriv):
        """False when ``gpt_ll_batched`` cannot serve this kernel's free parameters (the caller then evaluates one
        theta at a time through ``gpt_ll``)."""
        if with_deriv and set(self.fd_hyper_idxs) & set(int(i) for i in self.free_param_idxs):
            return False
        if with_deriv and self.kernel_id != 0:
            return not np.any(np.asarray(self.free_param_idxs) >= self.BATCHED_GRAD_SLOTS)
        return True

    def batch_rows_supported(self, param_rows):
        """Per-row validity of a (B, num_params) array of FULL parameter vectors for the device closed forms: rows
        flagged False evaluate to ``inf`` like the per-theta path, which raises before its device call."""
        return np.ones(np.atleast_2d(param_rows).shape[0], dtype=bool)

    def check_hyper_deriv(self, idxs):
        """Raise NotImplementedError (the reference's exception, kernel/core.py:723) when the derivative with
        respect to any of the parameter indices ``idxs`` is not available on the device."""
        if len(idxs) > 0 and not self.supports_hyper_deriv:
            raise NotImplementedError("Hyperparameter derivatives have not been implemented!")

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        if hyper_deriv is not None:
            self.check_hyper_deriv([int(hyper_deriv)])
        Xi = np.atleast_2d(np.asarray(Xi, dtype=float))
        Xj = np.atleast_2d(np.asarray(Xj, dtype=float))
        ni = np.atleast_2d(np.asarray(ni, dtype=int))
        nj = np.atleast_2d(np.asarray(nj, dtype=int))
        self._check_orders(ni, nj)
        kid, params = self.device_descriptor()
        Xi, ni = self.device_points(Xi, ni)
        Xj, nj = self.device_points(Xj, nj)
        if hyper_deriv is not None and int(hyper_deriv) in self.fd_hyper_idxs:
            return richardson_difference(lambda p_: default_device().cov_pairs(kid, p_, Xi, Xj, ni, nj), params,
                                         int(hyper_deriv), self.FD_HYPER_REL_STEP * abs(params[int(hyper_deriv)]))
        return default_device().cov_pairs(kid, params, Xi, Xj, ni, nj, hyper_deriv=hyper_deriv)


class BinaryKernel(Kernel):
    """Two kernels combined, with concatenated hyperparameters (kernel/core.py:424-548).

    When every operand is a device kernel the whole tree is evaluated ON THE DEVICE: it is flattened into a sum of
    products of its leaves (``_flatten``) and handed to the library as kernel id GPT_COMPOSITE
    (``gpt_define_composite``), so assembly, factorisation, gradients, prediction and the batched call run exactly as
    for a single kernel, with the general Leibniz rule over derivative orders applied per matrix entry.  Trees the
    library does not take (more than 4 leaves / 8 terms / 10 parameters, user-defined operands, Gibbs kernels with a
    host length-scale function) fall back to host-side composition of the operands' own calls, like the reference."""

    supports_hyper_deriv = True
    BATCHED_GRAD_SLOTS = 10

    # ---- device composition ---------------------------------------------------------------------------------
    def _flatten(self):
        """(leaves, terms): the leaf kernels in parameter order and the product terms (tuples of leaf indices) whose
        sum equals this kernel; None when an operand has no device closed form."""
        def walk(k):
            if isinstance(k, BinaryKernel):
                a, b = walk(k.k1), walk(k.k2)
                if a is None or b is None:
                    return None
                leaves = a[0] + b[0]
                shift = len(a[0])
                tb = [tuple(i + shift for i in t) for t in b[1]]
                if isinstance(k, SumKernel):
                    return leaves, a[1] + tb
                if isinstance(k, ProductKernel):
                    return leaves, [ta + t2 for ta in a[1] for t2 in tb]
                return None
            if isinstance(k, DeviceKernel) and k.kernel_id in (0, 1, 2, 3) and k.device_points_key() is None:
                return [k], [(0,)]
            return None
        return walk(self)

    def device_descriptor(self):
        from .._lib import CompositeId, MAX_LEAVES, MAX_TERMS, MAX_PARAMS
        flat = self._flatten()
        if flat is None:
            return None
        leaves, terms = flat
        if len(leaves) > MAX_LEAVES or len(terms) > MAX_TERMS or self.num_params > MAX_PARAMS:
            return None
        params = []
        for k in leaves:
            params.append(k.device_descriptor()[1])
        masks = [sum(1 << i for i in t) for t in terms]
        cid = CompositeId([k.kernel_id for k in leaves], [k.num_params for k in leaves], masks)
        return (cid, np.concatenate(params))

    def _leaf_offsets(self):
        leaves = self._flatten()[0]
        offs = np.cumsum([0] + [k.num_params for k in leaves])
        return leaves, offs

    @property
    def fd_hyper_idxs(self):
        """Parameter indices (into the concatenated vector) differentiated by finite differences: those of the leaves."""
        flat = self._flatten()
        if flat is None:
            return ()
        leaves, offs = self._leaf_offsets()
        return tuple(int(offs[q]) + int(i) for q, k in enumerate(leaves) for i in k.fd_hyper_idxs)

    FD_HYPER_REL_STEP = 2e-3

    def check_hyper_deriv(self, idxs):
        """Per-leaf availability."""
        leaves, offs = self._leaf_offsets()
        for i in idxs:
            q = int(np.searchsorted(offs, int(i), side="right") - 1)
            leaves[q].check_hyper_deriv([int(i) - int(offs[q])])

    def _check_orders(self, ni, nj):
        for k in self._flatten()[0]:
            k._check_orders(ni, nj)

    def device_points(self, X, n):
        return X, n

    def device_points_key(self):
        return None

    def batchable(self, with_deriv):
        # the library decides: persistent many-theta kernel, or theta after theta beyond its gradient slots; free
        # parameters differentiated by finite differences take the per-theta path
        return not (with_deriv and set(self.fd_hyper_idxs) & set(int(i) for i in self.free_param_idxs))

    def batch_rows_supported(self, param_rows):
        param_rows = np.atleast_2d(param_rows)
        leaves, offs = self._leaf_offsets()
        ok = np.ones(param_rows.shape[0], dtype=bool)
        for q, k in enumerate(leaves):
            ok &= k.batch_rows_supported(param_rows[:, offs[q]:offs[q + 1]])
        return ok

    def _device_call(self, Xi, Xj, ni, nj, hyper_deriv):
        """``__call__`` through ``gpt_cov_pairs`` when the tree is device-evaluable, else None."""
        desc = self.device_descriptor()
        if desc is None:
            return None
        Xi = np.atleast_2d(np.asarray(Xi, dtype=float))
        Xj = np.atleast_2d(np.asarray(Xj, dtype=float))
        ni = np.atleast_2d(np.asarray(ni, dtype=int))
        nj = np.atleast_2d(np.asarray(nj, dtype=int))
        if hyper_deriv is not None:
            self.check_hyper_deriv([int(hyper_deriv)])
        self._check_orders(ni, nj)
        if hyper_deriv is not None and int(hyper_deriv) in self.fd_hyper_idxs:
            return richardson_difference(lambda p_: default_device().cov_pairs(desc[0], p_, Xi, Xj, ni, nj), desc[1],
                                         int(hyper_deriv), self.FD_HYPER_REL_STEP * abs(desc[1][int(hyper_deriv)]))
        return default_device().cov_pairs(desc[0], desc[1], Xi, Xj, ni, nj, hyper_deriv=hyper_deriv)

    def __init__(self, k1, k2):
        if not isinstance(k1, Kernel) or not isinstance(k2, Kernel):
            raise TypeError("Argument to SumKernel must be instances of type Kernel.")
        if k1.num_dim != k2.num_dim:
            raise ValueError("Both kernels must have the same number of dimensions!")
        self.k1 = k1
        self.k2 = k2
        self.num_dim = k1.num_dim

    @property
    def num_params(self):
        return self.k1.num_params + self.k2.num_params

    # Write-through views of the operands' arrays, with setters that split and forward (kernel/core.py:466-548):
    # ``gp.params[i] = v``, ``gp.free_params = ...`` and friends must reach k1 / k2.
    def _split_set(self, attr, value, n1):
        value = list(value) if not isinstance(value, np.ndarray) else value
        if len(value) != n1 + len(getattr(self.k2, attr)):
            raise ValueError("Length of %s must be %d!" % (attr, n1 + len(getattr(self.k2, attr))))
        setattr(self.k1, attr, value[:n1])
        setattr(self.k2, attr, value[n1:])

    @property
    def params(self):
        return CombinedBounds(self.k1.params, self.k2.params)

    @params.setter
    def params(self, value):
        value = np.asarray(value, dtype=float)
        n1 = self.k1.num_params
        if len(value) != self.num_params:
            raise ValueError("Length of params must be %d!" % self.num_params)
        self.k1.params[:] = value[:n1]
        self.k2.params[:] = value[n1:]

    @property
    def fixed_params(self):
        return CombinedBounds(self.k1.fixed_params, self.k2.fixed_params)

    @fixed_params.setter
    def fixed_params(self, value):
        value = np.asarray(value, dtype=bool)
        n1 = self.k1.num_params
        if len(value) != self.num_params:
            raise ValueError("Length of fixed_params must be %d!" % self.num_params)
        self.k1.fixed_params = value[:n1]
        self.k2.fixed_params = value[n1:]

    @property
    def param_names(self):
        return CombinedBounds(self.k1.param_names, self.k2.param_names)

    @param_names.setter
    def param_names(self, value):
        n1 = self.k1.num_params
        if len(value) != self.num_params:
            raise ValueError("Length of param_names must be %d!" % self.num_params)
        self.k1.param_names = np.asarray(value[:n1], dtype=str)
        self.k2.param_names = np.asarray(value[n1:], dtype=str)

    @property
    def num_free_params(self):
        return self.k1.num_free_params + self.k2.num_free_params

    @property
    def free_param_idxs(self):
        return np.concatenate((np.asarray(self.k1.free_param_idxs, dtype=int),
                               np.asarray(self.k2.free_param_idxs, dtype=int) + self.k1.num_params))

    @property
    def free_params(self):
        return CombinedBounds(self.k1.free_params, self.k2.free_params)

    @free_params.setter
    def free_params(self, value):
        value = np.asarray(value, dtype=float)
        n1 = self.k1.num_free_params
        if len(value) != self.num_free_params:
            raise ValueError("Length of free_params must be %d!" % self.num_free_params)
        self.k1.free_params = value[:n1]
        self.k2.free_params = value[n1:]

    @property
    def free_param_bounds(self):
        return CombinedBounds(self.k1.free_param_bounds, self.k2.free_param_bounds)

    @free_param_bounds.setter
    def free_param_bounds(self, value):
        n1 = self.k1.num_free_params
        self.k1.free_param_bounds = value[:n1]
        self.k2.free_param_bounds = value[n1:]

    @property
    def free_param_names(self):
        return CombinedBounds(self.k1.free_param_names, self.k2.free_param_names)

    @free_param_names.setter
    def free_param_names(self, value):
        n1 = self.k1.num_free_params
        self.k1.free_param_names = value[:n1]
        self.k2.free_param_names = value[n1:]

    @property
    def hyperprior(self):
        return self.k1.hyperprior * self.k2.hyperprior

    @property
    def enforce_bounds(self):
        return self.k1.enforce_bounds or self.k2.enforce_bounds

    def set_hyperparams(self, new_params):
        new_params = np.asarray(new_params, dtype=float)
        n1 = self.k1.num_free_params
        if len(new_params) != n1 + self.k2.num_free_params:
            raise ValueError("Length of new_params must be {:d}!".format(n1 + self.k2.num_free_params))
        self.k1.set_hyperparams(new_params[:n1])
        self.k2.set_hyperparams(new_params[n1:])

class SumKernel(BinaryKernel):
    """k1 + k2 (kernel/core.py:549-600).  Used by the GP when the noise kernel is neither ZeroKernel nor
    DiagonalNoiseKernel (gaussian_process.py:1489-1490).  Device operands: evaluated on the device (BinaryKernel);
    otherwise the host composition below."""

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        out = self._device_call(Xi, Xj, ni, nj, hyper_deriv)
        if out is not None:
            return out
        if hyper_deriv is None:
            return (self.k1(Xi, Xj, ni, nj, symmetric=symmetric) + self.k2(Xi, Xj, ni, nj, symmetric=symmetric))
        if hyper_deriv < self.k1.num_params:
            return self.k1(Xi, Xj, ni, nj, hyper_deriv=hyper_deriv, symmetric=symmetric)
        return self.k2(Xi, Xj, ni, nj, hyper_deriv=hyper_deriv - self.k1.num_params, symmetric=symmetric)


class ProductKernel(BinaryKernel):
    """k1 * k2 with the general Leibniz rule over the derivative orders (kernel/core.py:601-670).

    For a pair with combined orders m = (ni, nj) (2 D slots) the reference sums k1^(s) k2^(m - s) over every subset
    of the multiset of derivative slots; collecting equal terms gives
    sum_{a <= m} prod_i C(m_i, a_i) k1^(a) k2^(m - a), which is what is evaluated here (one call of each operand per
    distinct split instead of one per subset).  Device operands: the same sum runs per entry on the device
    (csrc/covfn.cuh: comp_eval), including ``hyper_deriv``; in the host composition ``hyper_deriv`` raises
    NotImplementedError like the reference."""

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        out = self._device_call(Xi, Xj, ni, nj, hyper_deriv)
        if out is not None:
            return out
        if hyper_deriv is not None:
            raise NotImplementedError("hyper_deriv keyword not yet supported!")
        from itertools import product as cartesian
        from math import comb
        Xi = np.atleast_2d(np.asarray(Xi, dtype=float))
        Xj = np.atleast_2d(np.asarray(Xj, dtype=float))
        ni = np.atleast_2d(np.asarray(ni, dtype=int))
        nj = np.atleast_2d(np.asarray(nj, dtype=int))
        D = self.num_dim
        nij = np.hstack((ni, nj))
        result = np.zeros(Xi.shape[0])
        patterns, inverse = np.unique(nij, axis=0, return_inverse=True)
        inverse = np.ravel(inverse)
        for p_idx, m in enumerate(patterns):
            sel = inverse == p_idx
            xi, xj = Xi[sel], Xj[sel]
            cnt = int(sel.sum())
            for a in cartesian(*[range(int(mi) + 1) for mi in m]):
                a = np.array(a, dtype=int)
                weight = 1
                for mi, ai in zip(m, a):
                    weight *= comb(int(mi), int(ai))
                n1 = np.tile(a, (cnt, 1))
                n2 = np.tile(m - a, (cnt, 1))
                result[sel] += weight * (self.k1(xi, xj, n1[:, :D], n1[:, D:]) * self.k2(xi, xj, n2[:, :D], n2[:, D:]))
        return result
