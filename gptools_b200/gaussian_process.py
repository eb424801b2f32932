"""GaussianProcess: the reference's public class (gptools/gaussian_process.py:56-1605) with its numerical
core -- covariance assembly, Cholesky, alpha, log-likelihood, gradient, prediction, sampling -- executed by
the CUDA library through the C-ABI (include/gptb200.h).  Host code here is bookkeeping only: argument
canonicalisation (bit-exact with the reference's add_data / predict preamble), hyperparameter plumbing,
hyperpriors and mean functions (O(P) / O(M) numpy), and the optimizer / sampler drivers.

There is no CPU fallback for the numerics: every path below ends in a ``Device`` call.
"""
import multiprocessing
import sys
import traceback
import warnings

import numpy as np
import numpy.linalg
import scipy.linalg
import scipy.optimize
import scipy.stats

from .error_handling import GPArgumentError, GPImpossibleParamsError
from ._lib import CompositeId
from .kernel import DiagonalNoiseKernel, Kernel, ZeroKernel
from .kernel.core import default_device
from .utils import CombinedBounds

__all__ = ["GaussianProcess", "Constraint"]

EPS = sys.float_info.epsilon


def _has_iter(v):
    try:
        iter(v)
    except TypeError:
        return False
    return True


class GaussianProcess(object):
    """Gaussian process with derivative observations and linearly transformed observations.

    Same constructor and public attributes as the reference (gaussian_process.py:196-238):
    ``GaussianProcess(k, noise_k=None, X=None, y=None, err_y=0, n=0, T=None, diag_factor=1e2, mu=None,
    use_hyper_deriv=False, verbose=False)``; one extra keyword, ``device``, selects the CUDA device
    (default: ``LOCAL_RANK`` or 0).

    ``K``, ``noise_K``, ``L`` and ``alpha`` are fetched from the device on first access after each update
    instead of being copied back eagerly (at M = 49152 the factor alone is 19 GB).
    """

    def __init__(self, k, noise_k=None, X=None, y=None, err_y=0, n=0, T=None, diag_factor=1e2, mu=None,
                 use_hyper_deriv=False, verbose=False, device=None):
        if not isinstance(k, Kernel):
            raise TypeError("Argument k must be an instance of Kernel when constructing GaussianProcess!")
        if noise_k is None:
            noise_k = ZeroKernel(k.num_dim)
        elif not isinstance(noise_k, Kernel):
            raise TypeError("Keyword noise_k must be an instance of Kernel when constructing GaussianProcess!")
        self.mu = mu
        self.diag_factor = diag_factor
        self.k = k
        self.noise_k = noise_k
        self.use_hyper_deriv = use_hyper_deriv
        self.verbose = verbose
        self.y = np.array([], dtype=float)
        self.X = None
        self.err_y = np.array([], dtype=float)
        self.n = None
        self.T = None
        self._device_index = device
        self._dev_obj = None
        self._dev_data_version = -1
        self._data_version = 0
        self._dev_y_key = None
        self._dev_kernel_key = None
        self._cache = {}
        self._up_to_date = False
        self.ll = None
        self.ll_deriv = None
        if X is not None:
            if y is None:
                raise GPArgumentError("Must pass both X and y when constructing GaussianProcess!")
            self.add_data(X, y, err_y=err_y, n=n, T=T)
        elif y is not None:
            raise GPArgumentError("Must pass both X and y when constructing GaussianProcess!")

    # ------------------------------------------------------------------------------------------
    # state flags / lazily fetched device results
    # ------------------------------------------------------------------------------------------
    @property
    def K_up_to_date(self):
        return self._up_to_date

    @K_up_to_date.setter
    def K_up_to_date(self, value):
        self._up_to_date = bool(value)
        if not value:
            self._cache = {}

    def _lazy(self, name, fetch):
        self.compute_K_L_alpha_ll()
        if name not in self._cache:
            self._cache[name] = fetch()
        return self._cache[name]

    @property
    def K(self):
        """Latent covariance K(X, X) without noise (N x N)."""
        if self._device_mode():
            return self._lazy("K", lambda: self._dev().get_K())
        return self._lazy("K", lambda: self.compute_Kij(self.X, None, self.n, None, noise=False))

    @property
    def noise_K(self):
        def build():
            N = self.X.shape[0]
            if isinstance(self.noise_k, ZeroKernel):
                return np.zeros((N, N))
            if isinstance(self.noise_k, DiagonalNoiseKernel):
                return self.noise_k.params[0] ** 2.0 * np.eye(N)
            return self.compute_Kij(self.X, None, self.n, None, noise=True)
        return self._lazy("noise_K", build)

    @property
    def L(self):
        """Lower Cholesky factor of the total observation covariance (M x M)."""
        return self._lazy("L", lambda: self._dev().get_L())

    @property
    def alpha(self):
        """K_tot^{-1} (y - T mu), shape (M, 1) like the reference."""
        return self._lazy("alpha", lambda: self._dev().get_alpha()[:, None])

    def __getstate__(self):
        # device handles are per process: drop on pickle, re-create lazily (GP objects are pickled to
        # worker pools by the reference's drivers, gaussian_process.py:730, 1951)
        d = dict(self.__dict__)
        d["_dev_obj"] = None
        d["_dev_data_version"] = -1
        d["_dev_y_key"] = None
        d["_dev_kernel_key"] = None
        d["_cache"] = {}
        d["_up_to_date"] = False
        return d

    # ------------------------------------------------------------------------------------------
    # hyperparameter plumbing (gaussian_process.py:246-374)
    # ------------------------------------------------------------------------------------------
    def _parts(self):
        return [self.k, self.noise_k] + ([self.mu] if self.mu is not None else [])

    def _combined(self, attr):
        out = CombinedBounds(getattr(self.k, attr), getattr(self.noise_k, attr))
        if self.mu is not None:
            out = CombinedBounds(out, getattr(self.mu, attr))
        return out

    def _scatter(self, attr, value, counts):
        pos = 0
        for part, cnt in zip(self._parts(), counts):
            setattr(part, attr, value[pos:pos + cnt])
            pos += cnt

    def _num_params_each(self):
        return [p.num_params for p in self._parts()]

    def _num_free_each(self):
        return [p.num_free_params for p in self._parts()]

    @property
    def hyperprior(self):
        hp = self.k.hyperprior * self.noise_k.hyperprior
        if self.mu is not None:
            hp = hp * self.mu.hyperprior
        return hp

    @property
    def fixed_params(self):
        return self._combined("fixed_params")

    @fixed_params.setter
    def fixed_params(self, value):
        self._scatter("fixed_params", np.asarray(value, dtype=bool), self._num_params_each())

    @property
    def params(self):
        return self._combined("params")

    @params.setter
    def params(self, value):
        self.K_up_to_date = False
        self._scatter("params", np.asarray(value, dtype=float), self._num_params_each())

    @property
    def param_bounds(self):
        return self.hyperprior.bounds

    @param_bounds.setter
    def param_bounds(self, value):
        self.hyperprior.bounds = value

    @property
    def param_names(self):
        return self._combined("param_names")

    @param_names.setter
    def param_names(self, value):
        self._scatter("param_names", value, self._num_params_each())

    @property
    def free_params(self):
        return self._combined("free_params")

    @free_params.setter
    def free_params(self, value):
        self.K_up_to_date = False
        self._scatter("free_params", np.asarray(value, dtype=float), self._num_free_each())

    @property
    def free_param_bounds(self):
        return self._combined("free_param_bounds")

    @free_param_bounds.setter
    def free_param_bounds(self, value):
        self._scatter("free_param_bounds", np.asarray(value, dtype=float), self._num_free_each())

    @property
    def free_param_names(self):
        return self._combined("free_param_names")

    @free_param_names.setter
    def free_param_names(self, value):
        self.K_up_to_date = False
        self._scatter("free_param_names", np.asarray(value, dtype=str), self._num_free_each())

    @property
    def num_dim(self):
        return self.k.num_dim

    # ------------------------------------------------------------------------------------------
    # data (gaussian_process.py:376-503) -- row order of X, n, T, y, err_y is part of the contract
    # ------------------------------------------------------------------------------------------
    def add_data(self, X, y, err_y=0, n=0, T=None):
        y = np.atleast_1d(np.asarray(y, dtype=float))
        if len(y.shape) != 1:
            raise ValueError("Training targets y must have only one dimension with length greater than one! "
                             "Shape of y given is {}".format(y.shape))
        if not _has_iter(err_y):
            err_y = err_y * np.ones_like(y, dtype=float)
        else:
            err_y = np.asarray(err_y, dtype=float)
            if err_y.shape != y.shape:
                raise ValueError("When using array-like err_y, shape must match shape of y! Shape of err_y given "
                                 "is {}, shape of y given is {}.".format(err_y.shape, y.shape))
        if (err_y < 0).any():
            raise ValueError("All elements of err_y must be non-negative!")

        X = np.atleast_2d(np.asarray(X, dtype=float))
        if self.num_dim == 1 and X.shape[0] == 1:
            X = X.T
        if T is None and X.shape != (len(y), self.num_dim):
            raise ValueError("Shape of training inputs must be (len(y), k.num_dim)! X given has shape {}, shape of "
                             "y is {} and num_dim={:d}.".format(X.shape, y.shape, self.num_dim))

        if not _has_iter(n):
            n = n * np.ones_like(X, dtype=int)
        else:
            n = np.atleast_2d(np.asarray(n, dtype=int))
            if self.num_dim == 1 and n.shape[1] != 1:
                n = n.T
            if n.shape != X.shape:
                raise ValueError("When using array-like n, shape must be (len(y), k.num_dim)! Shape of n given is "
                                 "{}, shape of y given is {} and num_dim={:d}.".format(n.shape, y.shape, self.num_dim))
        if (n < 0).any():
            raise ValueError("All elements of n must be non-negative integers!")

        if T is None and self.T is not None:
            T = np.eye(len(y))
        if T is not None:
            T = np.atleast_2d(np.asarray(T, dtype=float))
            if T.ndim != 2:
                raise ValueError("T must have exactly 2 dimensions!")
            if T.shape[0] != len(y):
                raise ValueError("T must have as many rows are there are elements in y!")
            if T.shape[1] != X.shape[0]:
                raise ValueError("There must be as many columns in T as there are rows in X!")
            if self.T is None and self.X is not None:
                self.T = np.eye(len(self.y))
            self.T = T if self.T is None else scipy.linalg.block_diag(self.T, T)

        self.X = X if self.X is None else np.vstack((self.X, X))
        self.y = np.append(self.y, y)
        self.err_y = np.append(self.err_y, err_y)
        self.n = n if self.n is None else np.vstack((self.n, n))
        self._data_version += 1
        self.K_up_to_date = False

    def condense_duplicates(self):
        """Merge duplicate (X, n) rows through the transformation matrix and, when a transformation matrix is present,
        drop the quadrature points whose weights are all zero (gaussian_process.py:505-541)."""
        from .utils import unique_rows
        unique, inv = unique_rows(np.hstack((self.X, self.n)), return_inverse=True)
        if len(unique) != len(self.X):
            if self.T is None:
                self.T = np.eye(len(self.y))
            new_T = np.zeros((len(self.y), unique.shape[0]))
            for j in range(len(inv)):
                new_T[:, inv[j]] += self.T[:, j]
            self.T = new_T
            self.n = np.asarray(unique[:, self.X.shape[1]:], dtype=int)
            self.X = unique[:, :self.X.shape[1]]
            self._data_version += 1
            self.K_up_to_date = False
        if self.T is not None:
            # columns of T (= quadrature points) that actually enter (gaussian_process.py:535-541)
            good_cols = (self.T != 0.0).any(axis=0)
            if not good_cols.all():
                self.T = self.T[:, good_cols]
                self.X = self.X[good_cols, :]
                self.n = self.n[good_cols, :]
                self._data_version += 1
                self.K_up_to_date = False

    def remove_outliers(self, thresh=3, **predict_kwargs):
        """Drop observations more than ``thresh`` * ``err_y`` from the GP mean (gaussian_process.py:543-621).

        Returns ``(X_bad, y_bad, err_y_bad, n_bad, bad_idxs)`` and, when the GP has a transformation matrix, a sixth
        element ``T_bad``.  With T the reference keeps a quadrature point only if EVERY remaining row weights it
        (``(T != 0).all(axis=0)``, :603-604 and :613-614); that rule is reproduced as it stands."""
        mean = self.predict(self.X, n=self.n, noise=False, return_std=False, output_transform=self.T, **predict_kwargs)
        with np.errstate(divide="ignore", invalid="ignore"):
            deltas = np.absolute(mean - self.y) / self.err_y
        deltas[self.err_y == 0] = 0
        bad_idxs = (deltas >= thresh)
        good_idxs = ~bad_idxs
        y_bad = self.y[bad_idxs]
        err_y_bad = self.err_y[bad_idxs]
        if self.T is not None:
            T_bad = self.T[bad_idxs, :]
            non_zero_cols = (T_bad != 0).all(axis=0)
            T_bad = T_bad[:, non_zero_cols]
            X_bad = self.X[non_zero_cols, :]
            n_bad = self.n[non_zero_cols, :]
        else:
            X_bad = self.X[bad_idxs, :]
            n_bad = self.n[bad_idxs, :]
        if self.T is None:
            self.X = self.X[good_idxs, :]
            self.n = self.n[good_idxs, :]
        else:
            self.T = self.T[good_idxs, :]
            non_zero_cols = (self.T != 0).all(axis=0)
            self.T = self.T[:, non_zero_cols]
            self.X = self.X[non_zero_cols, :]
            self.n = self.n[non_zero_cols, :]
        self.y = self.y[good_idxs]
        self.err_y = self.err_y[good_idxs]
        self._data_version += 1
        self.K_up_to_date = False
        if self.T is None:
            return (X_bad, y_bad, err_y_bad, n_bad, bad_idxs)
        return (X_bad, y_bad, err_y_bad, n_bad, bad_idxs, T_bad)

    # ------------------------------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------------------------------
    def _dev(self):
        if self._dev_obj is None:
            from ._lib import Device
            self._dev_obj = Device(self._device_index)
            self._dev_data_version = -1
            self._dev_y_key = None
        return self._dev_obj

    def _device_mode(self):
        """True when the whole path (assembly included) runs on the device: an accelerated kernel plus the
        diagonal / zero noise kernels the reference special-cases (gaussian_process.py:1434-1439)."""
        return self.k.device_descriptor() is not None and isinstance(self.noise_k, DiagonalNoiseKernel)

    def _noise_sigma(self):
        if isinstance(self.noise_k, ZeroKernel):
            return 0.0
        return float(self.noise_k.params[0])

    def _y_minus_mean(self):
        """y - T mu(X, n) (gaussian_process.py:1455-1461)."""
        if self.mu is None:
            return self.y
        mu_alph = self.mu(self.X, self.n)
        if self.T is not None:
            mu_alph = self.T.dot(mu_alph)
        return self.y - mu_alph

    def _sync_device(self):
        dev = self._dev()
        y_alph = self._y_minus_mean()
        aux_key = self.k.device_points_key() if self.k.device_descriptor() is not None else None
        if self._dev_data_version != self._data_version or getattr(self, "_dev_aux_key", None) != aux_key:
            if self.X is None or len(self.y) == 0:
                raise GPArgumentError("No training data: call add_data first")
            if self.k.device_descriptor() is not None:
                Xd, nd = self.k.device_points(self.X, self.n)
            else:
                # host-evaluated kernel: the device never reads the points (K, dK, K* arrive assembled), so the input
                # dimension is not limited by GPT_MAX_DIM
                Xd = np.zeros((self.X.shape[0], 1))
                nd = np.zeros((self.X.shape[0], 1), dtype=int)
            dev.set_data(Xd, nd, y_alph, self.err_y, self.T)
            self._dev_aux_key = aux_key
            self._dev_data_version = self._data_version
            self._dev_y_key = None if self.mu is None else tuple(self.mu.params)
            self._dev_kernel_key = None
        elif self.mu is not None and self._dev_y_key != tuple(self.mu.params):
            dev.set_y(y_alph)
            self._dev_y_key = tuple(self.mu.params)
        desc = self.k.device_descriptor()
        kid, nparams = (desc[0], len(desc[1])) if desc is not None else (0, self.num_dim + 1)
        key = (kid, nparams, float(self.diag_factor))
        if self._dev_kernel_key != key:
            dev.set_kernel(kid, nparams, float(self.diag_factor))
            self._dev_kernel_key = key
        return dev, y_alph

    # ------------------------------------------------------------------------------------------
    # the hot path
    # ------------------------------------------------------------------------------------------
    def compute_Kij(self, Xi, Xj, ni, nj, noise=False, hyper_deriv=None, k=None):
        """Covariance matrix between (Xi, ni) and (Xj, nj); Xj=None means symmetric
        (gaussian_process.py:1535-1605).  Accelerated kernels are assembled tile-wise on the device;
        other kernels are called on the flattened pair lists exactly like the reference."""
        if k is None:
            k = self.noise_k if noise else self.k
        Xi = np.atleast_2d(np.asarray(Xi, dtype=float))
        ni = np.atleast_2d(np.asarray(ni, dtype=int))
        desc = k.device_descriptor()
        if desc is not None:
            if hyper_deriv is not None:
                k.check_hyper_deriv([int(hyper_deriv)])
            k._check_orders(ni, ni if nj is None else np.atleast_2d(np.asarray(nj, dtype=int)))
            Xi_d, ni_d = k.device_points(Xi, ni)
            # a kernel other than the GP's own is evaluated on the process-wide handle: defining another composite
            # structure on the GP's handle would drop its resident factorisation
            dev = default_device() if (k is not self.k and isinstance(desc[0], CompositeId)) else self._dev()
            if Xj is None:
                return dev.compute_Kij(desc[0], desc[1], Xi_d, ni_d, hyper_deriv=hyper_deriv)
            Xj_d, nj_d = k.device_points(np.atleast_2d(np.asarray(Xj, dtype=float)),
                                         np.atleast_2d(np.asarray(nj, dtype=int)))
            return dev.compute_Kij(desc[0], desc[1], Xi_d, ni_d, Xj_d, nj_d, hyper_deriv=hyper_deriv)
        symmetric = Xj is None
        if symmetric:
            Xj, nj = Xi, ni
        Xj = np.atleast_2d(np.asarray(Xj, dtype=float))
        nj = np.atleast_2d(np.asarray(nj, dtype=int))
        Mi, Mj = Xi.shape[0], Xj.shape[0]
        Kij = k(np.repeat(Xi, Mj, axis=0), np.tile(Xj, (Mi, 1)), np.repeat(ni, Mj, axis=0), np.tile(nj, (Mi, 1)),
                hyper_deriv=hyper_deriv, symmetric=symmetric)
        return np.reshape(Kij, (Mi, -1))

    def _grad_layout(self):
        """Positions of the free parameters in ll_deriv: kernel, noise kernel, mean function (in that order).

        The reference writes the mean-function entries at ``i + len(knk.free_params)`` (gaussian_process.py:1514),
        which collides with the DiagonalNoiseKernel entry (:1485); here they follow the noise entries."""
        nk, nn = self.k.num_free_params, self.noise_k.num_free_params
        return nk, nn

    def compute_K_L_alpha_ll(self):
        """K, L, alpha and the log-posterior (gaussian_process.py:1418-1522), on the device.

        Raises numpy.linalg.LinAlgError when K_tot is not positive definite, like scipy.linalg.cholesky."""
        if self.K_up_to_date:
            return
        dev, y_alph = self._sync_device()
        nk_free, nn_free = self._grad_layout()
        n_free = len(self.free_params)
        want_grad = bool(self.use_hyper_deriv)
        if want_grad:
            warnings.warn("Use of hyperparameter derivatives is experimental!")
        ll_deriv = np.zeros(n_free) if want_grad else None
        if self._device_mode():
            kid, kparams = self.k.device_descriptor()
            grad_idx = None
            fd_cols = {}
            if want_grad:
                self.k.check_hyper_deriv(list(self.k.free_param_idxs))
                free_idx = [int(i) for i in self.k.free_param_idxs]
                fd_set = set(int(i) for i in getattr(self.k, "fd_hyper_idxs", ()))
                grad_idx = [i for i in free_idx if i not in fd_set]
                if nn_free > 0:
                    grad_idx.append(len(kparams))
                # parameters without a closed-form hyper-derivative on the device (the Matern order nu): Richardson
                # central differences of the device's ll, evaluated BEFORE the main call so that the factorisation
                # left resident belongs to the GP's own parameters
                from .kernel.core import richardson_difference

                def ll_at(p_):
                    val, _, st = dev.ll(p_, self._noise_sigma(), grad_idx=None)
                    if st != 0:
                        raise numpy.linalg.LinAlgError("%d-th leading minor of the array is not positive definite" % st)
                    return val
                for col, i in enumerate(free_idx):
                    if i in fd_set:
                        fd_cols[col] = float(richardson_difference(ll_at, kparams, i,
                                                                   self.k.FD_HYPER_REL_STEP * abs(kparams[i])))
            ll, grad, status = dev.ll(kparams, self._noise_sigma(), grad_idx=grad_idx)
            if status != 0:
                raise numpy.linalg.LinAlgError(
                    "%d-th leading minor of the array is not positive definite" % status)
            if want_grad and grad is not None:
                an_cols = [c for c in range(nk_free) if c not in fd_cols] + list(range(nk_free, nk_free + nn_free))
                for c, v in zip(an_cols, grad):
                    ll_deriv[c] = v
                for c, v in fd_cols.items():
                    ll_deriv[c] = v
        else:
            K = self.compute_Kij(self.X, None, self.n, None, noise=False)
            if isinstance(self.noise_k, ZeroKernel):
                Kn = K
            elif isinstance(self.noise_k, DiagonalNoiseKernel):
                Kn = K + self.noise_k.params[0] ** 2.0 * np.eye(self.X.shape[0])
            else:
                Kn = K + self.compute_Kij(self.X, None, self.n, None, noise=True)
            ll, status = dev.ll_from_K(Kn)
            if status != 0:
                raise numpy.linalg.LinAlgError(
                    "%d-th leading minor of the array is not positive definite" % status)
            self._cache["K"] = K
            if want_grad:
                if isinstance(self.noise_k, DiagonalNoiseKernel):
                    knk = self.k
                    if nn_free > 0:
                        # gaussian_process.py:1484-1488: dK = 2 sigma_n I_M over the observations, also with T
                        ll_deriv[nk_free] = dev.noise_grad(self.noise_k.params[0])
                else:
                    knk = self.k + self.noise_k
                idxs = np.arange(0, len(knk.params), dtype=int)[~np.asarray(knk.fixed_params, dtype=bool)]
                for i, pi in enumerate(idxs):
                    dK = self.compute_Kij(self.X, None, self.n, None, k=knk, hyper_deriv=int(pi))
                    ll_deriv[i] = dev.grad_from_dK(dK)
        self.ll = ll + self.hyperprior(self.params)
        if want_grad:
            if self.mu is not None:
                alpha = dev.get_alpha()
                for i, pi in enumerate(self.mu.free_param_idxs):
                    dmu = self.mu(self.X, self.n, hyper_deriv=int(pi))
                    if self.T is not None:
                        dmu = self.T.dot(dmu)
                    ll_deriv[nk_free + nn_free + i] = dmu.dot(alpha)
            all_idx = np.arange(0, len(self.params), dtype=int)[~np.asarray(self.fixed_params[:], dtype=bool)]
            params = self.params
            hp = self.hyperprior
            for i, pi in enumerate(all_idx):
                ll_deriv[i] += hp(params, hyper_deriv=int(pi))
            self.ll_deriv = ll_deriv
        self._up_to_date = True

    def update_hyperparameters(self, new_params, hyper_deriv_handling='default', exit_on_bounds=True,
                               inf_on_error=True):
        """Set the free hyperparameters and return -ll [and -grad] (gaussian_process.py:1332-1416)."""
        use_hyper_deriv = self.use_hyper_deriv
        if hyper_deriv_handling == 'value':
            self.use_hyper_deriv = False
        elif hyper_deriv_handling == 'deriv':
            self.use_hyper_deriv = True
        nk, nn = self.k.num_free_params, self.noise_k.num_free_params
        self.k.set_hyperparams(new_params[:nk])
        self.noise_k.set_hyperparams(new_params[nk:nk + nn])
        if self.mu is not None:
            self.mu.set_hyperparams(new_params[nk + nn:])
        self.K_up_to_date = False
        try:
            if exit_on_bounds and np.isinf(self.hyperprior(self.params)):
                raise GPImpossibleParamsError("Impossible values for params!")
            self.compute_K_L_alpha_ll()
        except Exception as e:
            self.use_hyper_deriv = use_hyper_deriv
            if not inf_on_error:
                raise e
            if not isinstance(e, GPImpossibleParamsError) and self.verbose:
                warnings.warn("Unhandled exception when updating GP! Exception was:\n{:s}\nState of params is: "
                              "{:s}".format(traceback.format_exc(), str(self.free_params[:])))
            if use_hyper_deriv and hyper_deriv_handling == 'default':
                return (np.inf, np.zeros(len(self.free_params)))
            if hyper_deriv_handling == 'deriv':
                return np.zeros(len(self.free_params))
            return np.inf
        self.use_hyper_deriv = use_hyper_deriv
        if use_hyper_deriv and hyper_deriv_handling == 'default':
            return (-1.0 * self.ll, -1.0 * self.ll_deriv)
        if hyper_deriv_handling == 'deriv':
            return -1.0 * self.ll_deriv
        return -1.0 * self.ll

    # aliases named in the task description
    def compute_ll(self):
        self.compute_K_L_alpha_ll()
        return self.ll

    # ------------------------------------------------------------------------------------------
    # many hyperparameter vectors at once (the batched device entry; no counterpart in the reference,
    # where every theta is a separate update_hyperparameters call, possibly in a worker process)
    # ------------------------------------------------------------------------------------------
    def update_hyperparameters_batch(self, thetas, with_deriv=None, rank=None, world_size=None):
        """Evaluate -ll (and -grad) for B free-parameter vectors in one device launch.

        thetas : (B, num_free_params).  Returns ``neg_ll`` (B,) or ``(neg_ll, neg_grad)``; entries whose
        parameters have zero prior probability or whose covariance is not positive definite are ``inf``
        with zero gradient -- the per-theta result ``update_hyperparameters(inf_on_error=True)`` gives.
        The GP's own hyperparameters are left unchanged.
        """
        thetas = np.atleast_2d(np.asarray(thetas, dtype=float))
        if with_deriv is None:
            with_deriv = bool(self.use_hyper_deriv)
        if not self._batchable(with_deriv):
            # host kernels and kernels whose per-point columns depend on theta
            return self._batch_by_loop(thetas, with_deriv)
        plan = self._batch_prepare(thetas, with_deriv)
        res = plan["dev"].ll_batched(plan["full_eval"], grad_idx=plan["grad_idx"], y_batch=plan["y_batch"],
                                     return_alpha=plan["need_alpha"])
        return self._batch_finish(plan, res[0], res[1], res[2], res[3] if plan["need_alpha"] else None)

    #: observations up to which ``gpt_ll_batched`` runs its persistent many-theta kernel (32 tiles of 64 rows); beyond
    #: that, and with a transformation matrix, the same call runs the thetas back to back through the single-matrix path
    BATCHED_KERNEL_MAX_M = 2048

    def _batchable(self, with_deriv):
        """True when ``gpt_ll_batched`` can evaluate this GP: an accelerated kernel with theta-independent point
        columns and -- for the persistent kernel with gradients -- free kernel parameters inside its gradient slots.
        Everything else (host kernels, GibbsKernel1d with a user length-scale function) takes the per-theta loop."""
        if not self._device_mode() or self.k.device_points_key() is not None:
            return False
        if with_deriv and set(getattr(self.k, "fd_hyper_idxs", ())) & set(int(i) for i in self.k.free_param_idxs):
            return False  # a free parameter whose gradient entry is a finite difference of ll: per-theta path
        persistent = self.T is None and len(self.y) <= self.BATCHED_KERNEL_MAX_M
        if persistent and not self.k.batchable(with_deriv):
            return False
        return True

    def _batch_prepare(self, thetas, with_deriv):
        """Host side of a batched evaluation that does not depend on the device results: full parameter rows,
        per-theta hyperprior, parameter validity, per-theta mean-function residuals."""
        B = thetas.shape[0]
        nk, nn = self.k.num_free_params, self.noise_k.num_free_params
        n_free = len(self.free_params)
        if thetas.shape[1] != n_free:
            raise ValueError("thetas must have shape (B, %d)" % n_free)
        dev, y_alph = self._sync_device()
        kid, kparams = self.k.device_descriptor()
        nparams = len(kparams)
        if self.T is not None or len(self.y) > self.BATCHED_KERNEL_MAX_M or isinstance(kid, CompositeId):
            # these batches may run theta after theta through the single-matrix path, which leaves the factorisation of
            # the LAST theta on the device: the next predict / alpha / L must refactor at the GP's own hyperparameters
            self.K_up_to_date = False
        base_row = np.concatenate([kparams, [self._noise_sigma()]])
        full = np.tile(base_row, (B, 1))
        kfree = self.k.free_param_idxs
        full[:, kfree] = thetas[:, :nk]
        if nn > 0:
            full[:, nparams] = thetas[:, nk]
        # hyperprior per theta (host, O(P) each) over ALL parameters, fixed ones included (gaussian_process.py:1469)
        all_params = np.tile(np.asarray(self.params[:], dtype=float), (B, 1))
        free_mask = ~np.asarray(self.fixed_params[:], dtype=bool)
        all_params[:, free_mask] = thetas
        hp = self.hyperprior
        logp = np.asarray(hp.logpdf_batch(all_params), dtype=float)
        ok = np.isfinite(logp)
        # rows the device closed forms do not cover (e.g. a Matern nu that is not a half-integer): same result as the
        # per-theta path, which raises before the device call and returns inf
        rows_ok = self.k.batch_rows_supported(full[:, :nparams])
        ok &= rows_ok
        y_batch = None
        if self.mu is not None and self.mu.num_free_params > 0:
            y_batch = np.empty((B, len(self.y)))
            saved = np.array(self.mu.params, dtype=float)
            try:
                for b in range(B):
                    self.mu.set_hyperparams(thetas[b, nk + nn:])
                    y_batch[b] = self._y_minus_mean()
            finally:
                self.mu.params[:] = saved
        grad_idx = None
        if with_deriv:
            self.k.check_hyper_deriv(list(kfree))
            grad_idx = list(kfree) + ([nparams] if nn > 0 else [])
        need_alpha = bool(with_deriv and self.mu is not None and self.mu.num_free_params > 0)
        full_eval = full if ok.all() else np.where(ok[:, None], full, np.tile(base_row, (B, 1)))
        return dict(dev=dev, thetas=thetas, B=B, nk=nk, nn=nn, n_free=n_free, with_deriv=with_deriv, logp=logp, ok=ok,
                    y_batch=y_batch, grad_idx=grad_idx, need_alpha=need_alpha, full_eval=full_eval,
                    all_params=all_params, free_mask=free_mask, full=full, rows_ok=rows_ok, base_row=base_row)

    def _batch_plan_from_gathered(self, thetas, with_deriv, logp):
        """The part of a ``_batch_prepare`` plan that ``_batch_finish`` reads, for a batch whose per-row log-prior was
        computed elsewhere (-inf marks rows outside the prior support or not evaluable): the sharded entry prepares
        every row on one rank only and gathers the log-prior with the device results."""
        B = thetas.shape[0]
        nk, nn = self.k.num_free_params, self.noise_k.num_free_params
        all_params = np.empty((B, len(self.params)))
        all_params[:] = np.asarray(self.params[:], dtype=float)
        free_mask = ~np.asarray(self.fixed_params[:], dtype=bool)
        all_params[:, free_mask] = thetas
        logp = None if logp is None else np.asarray(logp, dtype=float)   # None: filled in by the caller after the gather
        return dict(thetas=thetas, B=B, nk=nk, nn=nn, n_free=len(self.free_params), with_deriv=with_deriv, logp=logp,
                    ok=None if logp is None else np.isfinite(logp), need_alpha=False, all_params=all_params,
                    free_mask=free_mask)

    def _batch_finish(self, plan, ll, grad, status, alpha=None):
        """-ll / -grad from the device results of ``_batch_prepare``'s rows (prior terms, inf / zero masks)."""
        B, nk, nn, n_free = plan["B"], plan["nk"], plan["nn"], plan["n_free"]
        thetas, logp, all_params, free_mask = plan["thetas"], plan["logp"], plan["all_params"], plan["free_mask"]
        good = plan["ok"] & (np.asarray(status) == 0)
        neg_ll = np.where(good, -(np.asarray(ll) + logp), np.inf)
        if not plan["with_deriv"]:
            return neg_ll
        g = np.zeros((B, n_free))
        if grad is not None and grad.shape[1] > 0:
            g[:, :grad.shape[1]] = grad
        if plan["need_alpha"]:
            saved = np.array(self.mu.params, dtype=float)
            try:
                for b in np.nonzero(good)[0]:
                    self.mu.set_hyperparams(thetas[b, nk + nn:])
                    for i, pi in enumerate(self.mu.free_param_idxs):
                        g[b, nk + nn + i] = self.mu(self.X, self.n, hyper_deriv=int(pi)).dot(alpha[b])
            finally:
                self.mu.params[:] = saved
        hp = self.hyperprior
        free_idx = np.nonzero(free_mask)[0]
        all_good = bool(good.all())
        good_params = all_params if all_good else all_params[good]
        for i, pi in enumerate(free_idx):
            dlp = hp.dlogpdf_batch(good_params, int(pi))
            if all_good:
                g[:, i] += dlp
            else:
                g[good, i] += dlp
        if not all_good:
            g[~good] = 0.0
        return neg_ll, -g

    def _eval_batch(self, thetas, with_deriv):
        """The unit of work of the L4 drivers (multi-start optimisation, ensemble sampler, ll grids): one batched
        device evaluation, split over the ranks of the default torch.distributed process group when there is one
        (the reference farms the same units to worker processes, gaussian_process.py:723-735, 1757-1763)."""
        if "torch" in sys.modules:
            from . import parallel
            if parallel.world()[1] > 1:
                return parallel.update_hyperparameters_batch_sharded(self, thetas, with_deriv=with_deriv)
        return self.update_hyperparameters_batch(thetas, with_deriv=with_deriv)

    def _set_free_params(self, values):
        """Write the free parameters of kernel, noise kernel and mean function (the split update_hyperparameters
        uses, gaussian_process.py:1369-1375)."""
        values = np.asarray(values, dtype=float)
        nk, nn = self.k.num_free_params, self.noise_k.num_free_params
        self.k.set_hyperparams(values[:nk])
        self.noise_k.set_hyperparams(values[nk:nk + nn])
        if self.mu is not None:
            self.mu.set_hyperparams(values[nk + nn:])
        self.K_up_to_date = False

    def _batch_by_loop(self, thetas, with_deriv):
        saved = np.array(self.free_params[:], dtype=float)
        saved_flag = self.use_hyper_deriv
        out, grads = [], []
        try:
            self.use_hyper_deriv = with_deriv
            for th in thetas:
                r = self.update_hyperparameters(th)
                if with_deriv:
                    out.append(r[0])
                    grads.append(r[1])
                else:
                    out.append(r)
        finally:
            self.use_hyper_deriv = saved_flag
            self._set_free_params(saved)
        if with_deriv:
            return np.asarray(out), np.asarray(grads)
        return np.asarray(out)

    # ------------------------------------------------------------------------------------------
    # prediction (gaussian_process.py:785-1034)
    # ------------------------------------------------------------------------------------------
    def predict(self, Xstar, n=0, noise=False, return_std=True, return_cov=False, full_output=False,
                return_samples=False, num_samples=1, samp_kwargs={}, return_mean_func=False, use_MCMC=False,
                full_MC=False, rejection_func=None, ddof=1, output_transform=None, full_covar=None, **kwargs):
        if full_covar is not None:  # alias named in the task description
            return_cov = bool(full_covar)
        if use_MCMC:
            res = self.predict_MCMC(
                Xstar, n=n, noise=noise, return_std=return_std or full_output, return_cov=return_cov or full_output,
                return_samples=full_output and (return_samples or rejection_func),
                return_mean_func=full_output and return_mean_func, num_samples=num_samples, samp_kwargs=samp_kwargs,
                full_MC=full_MC, rejection_func=rejection_func, ddof=ddof, output_transform=output_transform, **kwargs)
            if full_output:
                return res
            if return_cov:
                return (res['mean'], res['cov'])
            if return_std:
                return (res['mean'], res['std'])
            return res['mean']

        Xstar = np.atleast_2d(np.asarray(Xstar, dtype=float))
        if self.num_dim == 1 and Xstar.shape[0] == 1:
            Xstar = Xstar.T
        if Xstar.shape[1] != self.num_dim:
            raise ValueError("Second dimension of Xstar must be equal to self.num_dim! Shape of Xstar given is "
                             "{}, num_dim is {:d}.".format(Xstar.shape, self.num_dim))
        if output_transform is not None:
            output_transform = np.atleast_2d(np.asarray(output_transform, dtype=float))
            if output_transform.ndim != 2:
                raise ValueError("output_transform must have exactly 2 dimensions! Shape of output_transform given "
                                 "is {}.".format(output_transform.shape))
            if output_transform.shape[1] != Xstar.shape[0]:
                raise ValueError("output_transform must have the same number of columns the number of rows in "
                                 "Xstar! Shape of output_transform given is {}, shape of Xstar is "
                                 "{}.".format(output_transform.shape, Xstar.shape))
        if not _has_iter(n):
            n = n * np.ones(Xstar.shape, dtype=int)
        else:
            n = np.atleast_2d(np.asarray(n, dtype=int))
            if self.num_dim == 1 and n.shape[0] == 1:
                n = n.T
            if n.shape != Xstar.shape:
                raise ValueError("When using array-like n, shape must match shape of Xstar! Shape of n given is "
                                 "{}, shape of Xstar given is {}.".format(n.shape, Xstar.shape))
        if (n < 0).any():
            raise ValueError("All elements of n must be non-negative integers!")

        self.compute_K_L_alpha_ll()
        need_second = return_std or return_cov or full_output or full_MC
        # the full M* x M* covariance is only formed when somebody asks for it; std alone needs diag(K**) - |v|^2
        need_cov = need_second and (return_cov or full_output or full_MC or return_samples or
                                    output_transform is not None)
        if not self._device_mode():
            # host-evaluated kernel (user-defined Python kernel, kernel sum): K* and the prior (co)variance of the
            # test points are assembled exactly like the reference does (gaussian_process.py:966, 984), the solves
            # run on the device against the resident factor
            Kstar = self.compute_Kij(self.X, Xstar, self.n, n)
            kss_diag = Kss = None
            if need_cov:
                Kss = self.compute_Kij(Xstar, None, n, None)
            elif need_second:
                kss_diag = np.asarray(self.k(Xstar, Xstar, n, n, symmetric=False), dtype=float).ravel()
            mean, var, covariance = self._dev().predict_from_Kstar(
                Kstar, kss_diag=kss_diag, Kss=Kss, want_var=need_second and not need_cov, want_cov=need_cov)
        else:
            self.k._check_orders(self.n, n)  # unsupported derivative orders raise before any device call
            Xs_d, ns_d = self.k.device_points(Xstar, n)
            mean, var, covariance = self._dev().predict(Xs_d, ns_d, want_var=need_second and not need_cov,
                                                        want_cov=need_cov)
        mean_func = None
        if self.mu is not None:
            mean_func = self.mu(Xstar, n)
            mean = mean + mean_func
        if output_transform is not None:
            mean = output_transform.dot(mean)
            if return_mean_func and mean_func is not None:
                mean_func = output_transform.dot(mean_func)
        if not need_second:
            return mean
        if noise and not isinstance(self.noise_k, ZeroKernel):
            if isinstance(self.noise_k, DiagonalNoiseKernel) and type(self.noise_k).__call__ is DiagonalNoiseKernel.__call__:
                # sigma_n^2 [n* == n_noise] on the diagonal (kernel/noise.py:103-110) without forming the M*^2 pair lists
                nvar = self.noise_k.params[0] ** 2.0 * np.all(n == np.asarray(self.noise_k.n), axis=1)
                if not need_cov:
                    var = var + nvar
                elif len(np.unique(Xstar, axis=0)) == len(Xstar):
                    covariance = covariance + np.diag(nvar)
                else:  # repeated test points share their noise in the reference (Xi == Xj off the diagonal too)
                    covariance = covariance + self.compute_Kij(Xstar, None, n, None, noise=True)
            elif need_cov:
                covariance = covariance + self.compute_Kij(Xstar, None, n, None, noise=True)
            else:
                var = var + np.diagonal(self.compute_Kij(Xstar, None, n, None, noise=True))
        if need_cov and output_transform is not None:
            covariance = output_transform.dot(covariance.dot(output_transform.T))
        samps = None
        if return_samples or full_MC:
            samps = self.draw_sample(Xstar, n=n, num_samp=num_samples, mean=mean, cov=covariance, **samp_kwargs)
            if rejection_func:
                good = [s for s in samps.T if rejection_func(s)]
                if len(good) == 0:
                    raise ValueError("Did not get any good samples!")
                samps = np.asarray(good, dtype=float).T
            if full_MC:
                mean = np.mean(samps, axis=1)
                covariance = np.cov(samps, rowvar=1, ddof=ddof)
        with np.errstate(invalid="ignore"):
            std = np.sqrt(np.diagonal(covariance)) if need_cov else np.sqrt(var)
        if full_output:
            out = {'mean': mean, 'std': std, 'cov': covariance}
            if samps is not None:
                out['samp'] = samps
            if return_mean_func and self.mu is not None:
                out['mean_func'] = mean_func
                out['cov_func'] = np.zeros((len(mean_func), len(mean_func)), dtype=float)
                out['std_func'] = np.zeros_like(mean_func)
                out['mean_without_func'] = mean - mean_func
                out['cov_without_func'] = covariance
                out['std_without_func'] = std
            return out
        if return_cov:
            return (mean, covariance)
        return (mean, std)

    # ------------------------------------------------------------------------------------------
    # sampling (gaussian_process.py:1155-1330)
    # ------------------------------------------------------------------------------------------
    def draw_sample(self, Xstar, n=0, num_samp=1, rand_vars=None, rand_type='standard normal', diag_factor=1e3,
                    method='cholesky', num_eig=None, mean=None, cov=None, modify_sign=None, **kwargs):
        if mean is None or cov is None:
            out = self.predict(Xstar, n=n, full_output=True, **kwargs)
            mean, cov = out['mean'], out['cov']
        if rand_vars is None and method != 'eig':
            try:
                # same sampler and global RNG as the reference (gaussian_process.py:1269)
                return np.random.multivariate_normal(mean, cov, num_samp).T
            except numpy.linalg.LinAlgError as e:
                if self.verbose:
                    warnings.warn("Failure when drawing from MVN! Falling back on eig. Exception was:\n{}".format(e),
                                  RuntimeWarning)
                method = 'eig'
        if num_eig is None or num_eig > len(mean):
            num_eig = len(mean)
        elif num_eig < 1:
            num_eig = 1
        if rand_vars is None:
            rand_vars = np.random.standard_normal((num_eig, num_samp))
        valid_types = ('standard normal', 'uniform')
        if rand_type not in valid_types:
            raise ValueError("rand_type {} not recognized! Valid options are: {}.".format(rand_type, valid_types))
        if rand_type == 'uniform':
            rand_vars = scipy.stats.norm.ppf(rand_vars)
        rand_vars = np.atleast_2d(np.asarray(rand_vars, dtype=float))
        if method == 'cholesky':
            rv = rand_vars[:num_eig, :]
            if rv.shape[0] != len(mean):
                raise ValueError("rand_vars must have one row per test point for method='cholesky'")
            samp, status = self._dev().draw_sample(mean, cov, rv, diag_factor * EPS)
            if status != 0:
                raise numpy.linalg.LinAlgError(
                    "%d-th leading minor of the array is not positive definite" % status)
            return samp
        if method == 'eig':
            # eigen-decomposition branch: off the hot path, plain LAPACK on the host
            Mx = len(mean)
            eig, Q = scipy.linalg.eigh(cov + diag_factor * EPS * np.eye(Mx),
                                       subset_by_index=(Mx - 1 - (num_eig - 1), Mx - 1))
            if modify_sign is not None:
                tests = {
                    'left value': lambda Q: Q[0, :] < 0.0,
                    'right value': lambda Q: Q[-1, :] < 0.0,
                    'left slope': lambda Q: (Q[1, :] - Q[0, :]) < 0.0,
                    'right slope': lambda Q: (Q[-1, :] - Q[-2, :]) < 0.0,
                    'left concavity': lambda Q: (Q[2, :] - 2 * Q[1, :] + Q[0, :]) < 0.0,
                    'right concavity': lambda Q: (Q[-1, :] - 2 * Q[-2, :] + Q[-3, :]) < 0.0,
                }
                if modify_sign not in tests:
                    raise ValueError("modify_sign {} not recognized!".format(modify_sign))
                Q[:, tests[modify_sign](Q)] *= -1.0
            Lq = Q.dot(np.diag(np.sqrt(eig)))
            return np.atleast_2d(mean).T + Lq.dot(rand_vars[:num_eig, :])
        raise ValueError("method {} not recognized!".format(method))

    # ------------------------------------------------------------------------------------------
    # MAP estimation (gaussian_process.py:623-783)
    # ------------------------------------------------------------------------------------------
    def optimize_hyperparameters(self, method='SLSQP', opt_kwargs={}, verbose=False, random_starts=None,
                                 num_proc=None, max_tries=1, batched_starts=True):
        """Maximise the log-posterior with scipy.optimize.minimize from ``random_starts`` starting points drawn
        from the hyperprior (gaussian_process.py:623-783).

        The reference maps the starts over a pool of ``num_proc`` worker processes (:723-735).  Here the starts advance
        in LOCK-STEP: every start runs the same scipy minimiser (same method, bounds, jac) in its own thread, the
        objective calls of all starts that are waiting are collected into one batch, and the batch is evaluated by a
        single device launch (``update_hyperparameters_batch``; split over the GPUs of an active process group).  The
        result of every start is what its own sequential ``minimize`` run returns -- the optimiser never sees the
        batching.  ``num_proc`` only sets the default number of starts (like the reference when ``random_starts`` is
        None); ``batched_starts=False`` runs the starts one after the other through ``gpt_ll``."""
        opt_kwargs = {} if opt_kwargs is None else dict(opt_kwargs)
        if 'method' in opt_kwargs:
            method = opt_kwargs['method']
            if self.verbose:
                warnings.warn("Key 'method' is present in opt_kwargs, will override option specified with method "
                              "kwarg.", RuntimeWarning)
        else:
            opt_kwargs['method'] = method
        if num_proc is None:
            num_proc = multiprocessing.cpu_count()
        param_ranges = np.asarray(self.free_param_bounds[:], dtype=float)
        param_ranges[np.isnan(param_ranges[:, 0]) | np.isinf(param_ranges[:, 0]), 0] = -1e16
        param_ranges[np.isnan(param_ranges[:, 1]) | np.isinf(param_ranges[:, 1]), 1] = 1e16
        free_mask = ~np.asarray(self.fixed_params[:], dtype=bool)

        def draw_starts():
            if random_starts == 0:
                return [np.array(self.free_params[:], dtype=float)]
            nstart = max(num_proc, 1) if random_starts is None else random_starts
            samples = self.hyperprior.random_draw(size=nstart).T
            return list(samples[:, free_mask])

        if 'bounds' not in opt_kwargs:
            opt_kwargs['bounds'] = param_ranges
        if self.use_hyper_deriv:
            opt_kwargs['jac'] = True
        trial = 0
        res_min = None
        res = []
        while trial < max_tries and res_min is None:
            if trial >= 1 and self.verbose:
                warnings.warn("No solutions found on trial {:d}, retrying random starts.".format(trial - 1),
                              RuntimeWarning)
            starts = draw_starts()
            trial += 1
            if len(starts) > 1 and batched_starts:
                res = self._minimize_starts_lockstep(starts, opt_kwargs)
            else:
                res = []
                for samp in starts:
                    try:
                        r = scipy.optimize.minimize(self.update_hyperparameters, samp, **opt_kwargs)
                    except Exception:
                        if self.verbose:
                            warnings.warn("Minimizer failed, skipping sample. Error is:\n{:s}\nState of params is: "
                                          "{:s}".format(traceback.format_exc(), str(self.free_params[:])), RuntimeWarning)
                        continue
                    res.append(r)
            finite = [r for r in res if not (np.isnan(r.fun) or np.isinf(r.fun))]
            res_min = min(finite, key=lambda r: r.fun) if finite else None
        if res_min is None:
            raise ValueError("Optimizer failed to find a valid solution. Try changing the parameter bounds, picking "
                             "a new initial guess or increasing the number of random starts.")
        self.update_hyperparameters(res_min.x)
        if verbose:
            print("Got {:d} completed starts, optimal result is:".format(len(res)))
            print(res_min)
            print("\nLL\t{:.3g}".format(-1 * res_min.fun))
            for v, l in zip(res_min.x, self.free_param_names):
                print("{:s}\t{:.3g}".format(str(l).replace('\\', ''), v))
        if not res_min.success:
            warnings.warn("Optimizer {:s} reports failure, selected hyperparameters are likely NOT optimal. Status: "
                          "{:d}, Message: '{:s}'. Try adjusting bounds, initial guesses or the number of random "
                          "starts used.".format(method, res_min.status, str(res_min.message)), RuntimeWarning)
        bounds = np.asarray(self.free_param_bounds[:], dtype=float)
        if ((res_min.x <= 1.001 * bounds[:, 0]).any() or (res_min.x >= 0.999 * bounds[:, 1]).any()):
            warnings.warn("Optimizer appears to have hit/exceeded the bounds. Bounds are:\n{:s}\n, solution is:\n"
                          "{:s}. Try adjusting bounds, initial guesses or the number of random starts "
                          "used.".format(str(bounds), str(res_min.x)))
        return (res_min, len(res))

    def _minimize_starts_lockstep(self, starts, opt_kwargs):
        """Run scipy.optimize.minimize from every start with the objective evaluations batched across the starts.

        Each start owns a thread that blocks inside its objective call; once EVERY unfinished start is blocked, the
        pending parameter vectors (ordered by start index, so the batch is deterministic and identical on every rank
        of a process group) go to the device as one batch and the threads are released with their own row."""
        import threading
        nstart = len(starts)
        want_grad = bool(opt_kwargs.get('jac', False))
        cond = threading.Condition()
        pending, results, active = {}, {}, set(range(nstart))
        out = [None] * nstart
        failure = []

        def objective(idx):
            def f(theta):
                with cond:
                    pending[idx] = np.array(theta, dtype=float)
                    cond.notify_all()
                    while idx not in results and not failure:
                        cond.wait()
                    if failure:
                        raise RuntimeError("batched evaluation failed")
                    return results.pop(idx)
            return f

        def worker(idx):
            try:
                out[idx] = scipy.optimize.minimize(objective(idx), starts[idx], **opt_kwargs)
            except Exception:
                if self.verbose:
                    warnings.warn("Minimizer failed, skipping sample. Error is:\n{:s}".format(traceback.format_exc()),
                                  RuntimeWarning)
            finally:
                with cond:
                    active.discard(idx)
                    pending.pop(idx, None)
                    cond.notify_all()

        threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(nstart)]
        for t in threads:
            t.start()
        saved = np.array(self.free_params[:], dtype=float)
        try:
            while True:
                with cond:
                    while active and set(pending) != active:
                        cond.wait()
                    if not active:
                        break
                    order = sorted(pending)
                    thetas = np.array([pending[i] for i in order])
                    pending.clear()
                try:
                    r = self._eval_batch(thetas, want_grad)
                except Exception:
                    with cond:
                        failure.append(traceback.format_exc())
                        cond.notify_all()
                    raise
                with cond:
                    for row, i in enumerate(order):
                        results[i] = (float(r[0][row]), np.array(r[1][row])) if want_grad else float(r[row])
                    cond.notify_all()
        finally:
            for t in threads:
                t.join()
            self._set_free_params(saved)
        return [r for r in out if r is not None]

    # ------------------------------------------------------------------------------------------
    # ll over a grid of hyperparameters (gaussian_process.py:1607-1692): one batched device launch
    # ------------------------------------------------------------------------------------------
    def compute_ll_matrix(self, bounds, num_pts):
        """Log-posterior on a regular grid over the free parameters.  Returns (ll_vals, param_vals) with
        ll_vals of shape (num_pts[0], ..., num_pts[P-1]) like the reference."""
        present = np.array(self.free_params[:], dtype=float)
        bounds = np.atleast_2d(np.asarray(bounds, dtype=float))
        if bounds.shape[1] != 2:
            raise ValueError("Argument bounds must have shape (n, 2)!")
        if bounds.shape[0] == 1:
            bounds = np.tile(bounds, (len(present), 1))
        if not _has_iter(num_pts):
            num_pts = num_pts * np.ones(bounds.shape[0], dtype=int)
        else:
            num_pts = np.asarray(num_pts, dtype=int)
            if len(num_pts) != len(present):
                raise ValueError("Length of num_pts must match the number of free parameters!")
        param_vals = [np.linspace(b[0], b[1], int(npt)) for b, npt in zip(bounds, num_pts)]
        grid = np.stack(np.meshgrid(*param_vals, indexing='ij'), axis=-1).reshape(-1, len(param_vals))
        neg_ll = self._eval_batch(grid, False)
        ll_vals = -np.asarray(neg_ll).reshape([len(v) for v in param_vals])
        return (ll_vals, param_vals)

    # ------------------------------------------------------------------------------------------
    # MCMC over the hyperparameters (gaussian_process.py:1694-1838): ensemble sampler fed by the batched call
    # ------------------------------------------------------------------------------------------
    def sample_hyperparameter_posterior(self, nwalkers=200, nsamp=500, burn=0, thin=1, num_proc=None, sampler=None,
                                        plot_posterior=False, plot_chains=False, sampler_type='ensemble',
                                        ntemps=20, sampler_a=2.0, **plot_kwargs):
        """Run an affine-invariant ensemble sampler (stretch move, the algorithm of emcee.EnsembleSampler used by
        the reference at gaussian_process.py:1757-1787) on the hyperparameter posterior.  Each half-ensemble
        proposal is evaluated as ONE batched device launch instead of ``nwalkers`` Python calls / worker processes.
        Returns the sampler object (``chain`` (nwalkers, nsamp, ndim), ``lnprobability``, ``flatchain``,
        ``acceptance_fraction``), which can be passed back in through ``sampler=`` to continue the chains."""
        from .sampler import EnsembleSampler
        if sampler_type != 'ensemble':
            raise NotImplementedError("only sampler_type='ensemble' is available")
        if plot_posterior or plot_chains:
            warnings.warn("plotting is out of scope of gptools_b200; ignoring plot_* keywords")
        ndim = len(self.free_params)
        nranks = 1
        if "torch" in sys.modules:
            from . import parallel
            nranks = parallel.world()[1]
        if sampler is None:
            # with several ranks every rank runs the same chain (proposals from identical random streams) and
            # evaluates its slice of each half-ensemble: the walkers are the sharded units (gaussian_process.py:1757-1763)
            rstate = np.random.RandomState(parallel.shared_seed()) if nranks > 1 else None
            sampler = EnsembleSampler(nwalkers, ndim, lambda th: -self._eval_batch(th, False), a=sampler_a,
                                      random_state=rstate)
        else:
            sampler.lnprob_batch = lambda th: -self._eval_batch(th, False)
        if sampler.chain.shape[1] == 0:
            theta0 = self.hyperprior.random_draw(size=nwalkers).T
            theta0 = theta0[:, ~np.asarray(self.fixed_params[:], dtype=bool)]
            if nranks > 1:
                theta0 = parallel.broadcast_array(theta0)
        else:
            theta0 = sampler.chain[:, -1, :]
        sampler.run_mcmc(theta0, nsamp)
        return sampler

    def predict_batch(self, thetas, Xstar, n=0, noise=False):
        """Predictive mean and standard deviation at ``Xstar`` for EVERY row of ``thetas`` (B, num_free_params) in one
        device launch -- the per-sample unit of ``compute_from_MCMC`` (update_hyperparameters + predict,
        gaussian_process.py:1944-1969), batched.  The test points ride along as extra tile rows of each theta's
        factorisation (``gpt_predict_batched``).

        Returns ``(mean (B, M*), std (B, M*), good (B,))``; rows with ``good == False`` (covariance not positive
        definite, parameters the closed forms do not take) are NaN, where the per-sample loop skips the sample.  Like the
        reference's per-sample wrapper (gaussian_process.py:2301-2330), a sample outside the prior support is still
        predicted: only a failing prediction drops it.  Returns ``None`` when the
        batch cannot run in the persistent kernel (transformation matrix, more than 2048 observations, host-evaluated
        kernels, theta-dependent point columns): the caller then predicts sample by sample.  The GP's own
        hyperparameters are left unchanged."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=float))
        if (not self._device_mode() or self.k.device_points_key() is not None or self.T is not None or
                len(self.y) > self.BATCHED_KERNEL_MAX_M):
            return None
        Xstar = np.atleast_2d(np.asarray(Xstar, dtype=float))
        if self.num_dim == 1 and Xstar.shape[0] == 1:
            Xstar = Xstar.T
        if Xstar.shape[1] != self.num_dim:
            raise ValueError("Second dimension of Xstar must be equal to self.num_dim!")
        if not _has_iter(n):
            n = n * np.ones(Xstar.shape, dtype=int)
        else:
            n = np.atleast_2d(np.asarray(n, dtype=int))
            if self.num_dim == 1 and n.shape[0] == 1:
                n = n.T
        if n.shape != Xstar.shape or (n < 0).any():
            raise ValueError("n must be non-negative and match the shape of Xstar!")
        self.k._check_orders(self.n, n)
        plan = self._batch_prepare(thetas, False)
        Xs_d, ns_d = self.k.device_points(Xstar, n)
        try:
            rows = np.where(plan["rows_ok"][:, None], plan["full"], plan["base_row"][None, :])
            mean, var, ll, status = plan["dev"].predict_batched(rows, Xs_d, ns_d, y_batch=plan["y_batch"])
        except NotImplementedError:
            return None
        good = plan["rows_ok"] & (status == 0)
        nk, nn = plan["nk"], plan["nn"]
        if self.mu is not None:
            saved = np.array(self.mu.params, dtype=float)
            try:
                if self.mu.num_free_params > 0:
                    for b in np.nonzero(good)[0]:
                        self.mu.set_hyperparams(thetas[b, nk + nn:])
                        mean[b] += self.mu(Xstar, n)
                else:
                    mean += self.mu(Xstar, n)[None, :]
            finally:
                self.mu.params[:] = saved
        if noise and not isinstance(self.noise_k, ZeroKernel):
            if not (isinstance(self.noise_k, DiagonalNoiseKernel) and
                    type(self.noise_k).__call__ is DiagonalNoiseKernel.__call__):
                return None
            sig = rows[:, -1]
            var = var + (sig ** 2.0)[:, None] * np.all(n == np.asarray(self.noise_k.n), axis=1)[None, :]
        with np.errstate(invalid="ignore"):
            std = np.sqrt(var)
        mean[~good] = np.nan
        std[~good] = np.nan
        return mean, std, good

    def compute_from_MCMC(self, X, n=0, return_mean=True, return_std=True, return_cov=False, return_samples=False,
                          return_mean_func=False, num_samples=1, noise=False, samp_kwargs={}, sampler=None,
                          flat_trace=None, burn=0, thin=1, **kwargs):
        """Predict at every retained hyperparameter sample (gaussian_process.py:1840-1990)."""
        output_transform = kwargs.pop('output_transform', None)
        if flat_trace is None:
            if sampler is None:
                sampler = self.sample_hyperparameter_posterior(burn=burn, **kwargs)
            flat_trace = sampler.chain[:, burn::thin, :]
            flat_trace = flat_trace.reshape((-1, flat_trace.shape[2]))
        else:
            flat_trace = np.asarray(flat_trace, dtype=float)[burn::thin, :]   # gaussian_process.py:1938
        saved = np.array(self.free_params[:], dtype=float)
        out = {k_: [] for k_ in ('mean', 'std', 'cov', 'samp', 'mean_func')}
        # the hyperparameter samples are independent units (the reference maps them over a process pool,
        # gaussian_process.py:1944-1969): with a process group every rank takes a contiguous slice
        rank, nranks = 0, 1
        if "torch" in sys.modules:
            from . import parallel
            rank, nranks = parallel.world()
        full_trace = flat_trace
        if nranks > 1:
            lo, hi = parallel.shard_bounds(len(flat_trace), rank, nranks)
            flat_trace = flat_trace[lo:hi]
        # mean / std requests: every sample of the slice in ONE launch (gpt_predict_batched); everything else (full
        # covariances, samples, output transforms, kernels outside the persistent batched kernel) sample by sample
        if (len(flat_trace) > 0 and not return_cov and not return_samples and output_transform is None and
                not getattr(self, "_mcmc_predict_by_loop", False)):
            res_b = self.predict_batch(flat_trace, X, n=n, noise=noise)
            if res_b is not None:
                mean_b, std_b, good = res_b
                saved_mu = None if self.mu is None else np.array(self.mu.params, dtype=float)
                nk_, nn_ = self.k.num_free_params, self.noise_k.num_free_params
                try:
                    for b_ in np.nonzero(good)[0]:
                        out['mean'].append(mean_b[b_])
                        out['std'].append(std_b[b_])
                        if return_mean_func and self.mu is not None:
                            self.mu.set_hyperparams(flat_trace[b_, nk_ + nn_:])
                            Xm = np.atleast_2d(np.asarray(X, dtype=float))
                            if self.num_dim == 1 and Xm.shape[0] == 1:
                                Xm = Xm.T
                            nm = n if _has_iter(n) else n * np.ones(Xm.shape, dtype=int)
                            out['mean_func'].append(self.mu(Xm, np.atleast_2d(np.asarray(nm, dtype=int))))
                finally:
                    if saved_mu is not None:
                        self.mu.params[:] = saved_mu
                flat_trace = flat_trace[:0]
        try:
            for th in flat_trace:
                # the reference's wrapper (gaussian_process.py:2301-2330) ignores the value update_hyperparameters
                # returns (inf for zero prior probability) and drops a sample only when the prediction itself fails
                try:
                    self.update_hyperparameters(th)
                    res = self.predict(X, n=n, noise=noise, full_output=True, return_samples=return_samples,
                                       num_samples=num_samples, samp_kwargs=samp_kwargs,
                                       return_mean_func=return_mean_func, output_transform=output_transform)
                except Exception:
                    continue
                out['mean'].append(res['mean'])
                out['std'].append(res['std'])
                if return_cov:
                    out['cov'].append(res['cov'])
                if return_samples:
                    out['samp'].append(res['samp'])
                if return_mean_func and self.mu is not None:
                    out['mean_func'].append(res['mean_func'])
        finally:
            self.update_hyperparameters(saved)
        if nranks > 1:
            out = parallel.gather_result_lists(out, len(full_trace))
        return {k_: v for k_, v in out.items() if len(v) > 0}

    def predict_MCMC(self, X, ddof=1, full_MC=False, rejection_func=None, **kwargs):
        """Prediction marginalised over the hyperparameter samples by the law of total (co)variance
        (gaussian_process.py:2136-2254)."""
        return_std = kwargs.get('return_std', True)
        return_cov = kwargs.get('return_cov', False)
        kwargs['return_cov'] = True if (return_cov or full_MC) else return_cov
        res = self.compute_from_MCMC(X, **kwargs)
        means = np.array(res['mean'], dtype=float)
        out = {'mean': means.mean(axis=0)}
        if return_cov and 'cov' in res:
            covs = np.array(res['cov'], dtype=float)
            out['cov'] = covs.mean(axis=0) + np.cov(means, rowvar=0, ddof=ddof)
            out['std'] = np.sqrt(np.diagonal(out['cov']))
        elif return_std:
            stds = np.array(res['std'], dtype=float)
            out['std'] = np.sqrt((stds ** 2).mean(axis=0) + np.var(means, axis=0, ddof=ddof))
        if 'samp' in res:
            out['samp'] = np.hstack(res['samp'])
        return out


class Constraint(object):
    """Inequality constraint on the GP mean (or a derivative) for constrained optimizers
    (gaussian_process.py:2488-2656): ``c(params) >= 0`` when the constraint is satisfied."""

    def __init__(self, gp, boundary_val=0.0, n=0, loc='min', type_='gt', bounds=None):
        if not isinstance(gp, GaussianProcess):
            raise TypeError("Argument gp must be an instance of GaussianProcess.")
        self.gp = gp
        self.boundary_val = boundary_val
        self.n = int(n)
        if self.n < 0:
            raise ValueError("n must be a non-negative int!")
        D = gp.num_dim
        if loc in ('min', 'max'):
            self.loc = loc
        else:
            loc = np.atleast_1d(np.asarray(loc, dtype=float))
            if loc.shape != (D,):
                raise ValueError("Argument loc must be 'min', 'max' or an array of length {:d}".format(D))
            self.loc = loc
        if type_ not in ('gt', 'lt'):
            raise ValueError("Argument type_ must be 'gt' or 'lt'.")
        self.type_ = type_
        if bounds is None:
            lo = np.asarray(gp.X.min(axis=0), dtype=float).flatten()
            hi = np.asarray(gp.X.max(axis=0), dtype=float).flatten()
        else:
            bounds = list(bounds)
            if len(bounds) != 2:
                raise ValueError("Argument bounds must have length 2!")
            lo, hi = (np.atleast_1d(np.asarray(b, dtype=float)) for b in bounds)
            if lo.shape != (D,) or hi.shape != (D,):
                raise ValueError("Each element in argument bounds must have length {:d}".format(D))
        self.bounds = list(zip(lo, hi))

    def __call__(self, params):
        self.gp.update_hyperparameters(params)
        if not isinstance(self.loc, str):
            val = self.gp.predict(self.loc, n=self.n, return_std=False)[0]
        else:
            factor = -1.0 if self.loc == 'max' else 1.0
            res = scipy.optimize.minimize(
                lambda X: factor * self.gp.predict(X, n=self.n, return_std=False)[0],
                np.mean(self.bounds, axis=1), method='SLSQP', bounds=self.bounds)
            if not res.success:
                warnings.warn("Solver reports failure, extremum was likely NOT found. Status: {:d}, Message: "
                              "'{:s}'".format(res.status, str(res.message)), RuntimeWarning)
            val = factor * res.fun
        return val - self.boundary_val if self.type_ == 'gt' else self.boundary_val - val
