"""ctypes binding of the C-ABI declared in include/gptb200.h (libgptb200.so).

This is the only place the Python host layer touches native code.  There is NO CPU fallback: if the
library is missing, or no CUDA device is present, every compute entry point raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPTB200_LIB", os.path.join(_HERE, "csrc", "libgptb200.so"))  # override: development builds

GPT_SE, GPT_MATERN52, GPT_MATERN, GPT_GIBBS_TANH = 0, 1, 2, 3
GPT_COMPOSITE = 5
MAX_LEAVES, MAX_TERMS, MAX_PARAMS = 4, 8, 10


class CompositeId(int):
    """Kernel id GPT_COMPOSITE together with the structure ``gpt_define_composite`` takes: the operand kernels in
    parameter order and the product terms (bit masks over the operands) whose sum is the kernel."""

    def __new__(cls, leaf_kids, leaf_nparams, term_masks):
        obj = int.__new__(cls, GPT_COMPOSITE)
        obj.leaf_kids = tuple(int(k) for k in leaf_kids)
        obj.leaf_nparams = tuple(int(k) for k in leaf_nparams)
        obj.term_masks = tuple(int(k) for k in term_masks)
        return obj

    @property
    def structure(self):
        return (self.leaf_kids, self.leaf_nparams, self.term_masks)

    def __reduce__(self):  # pickle / deepcopy: an int subclass would otherwise be rebuilt from its value alone
        return (CompositeId, self.structure)

    def __eq__(self, other):
        if isinstance(other, CompositeId):
            return self.structure == other.structure
        return int(self) == other

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((GPT_COMPOSITE,) + self.structure)

    def __repr__(self):
        return "CompositeId(%r, %r, %r)" % self.structure

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); exactly the symbols of include/gptb200.h
SIGNATURES = {
    "gpt_version": (ctypes.c_int, []),
    "gpt_device_count": (ctypes.c_int, []),
    "gpt_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "gpt_destroy": (None, [_vp]),
    "gpt_last_error": (ctypes.c_char_p, [_vp]),
    "gpt_set_stream": (ctypes.c_int, [_vp, _vp]),
    "gpt_use_own_stream": (ctypes.c_int, [_vp]),
    "gpt_synchronize": (ctypes.c_int, [_vp]),
    "gpt_set_data": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_double_p, _c_int32_p,
                                    _c_double_p, _c_double_p, _c_double_p]),
    "gpt_set_y": (ctypes.c_int, [_vp, _c_double_p]),
    "gpt_set_kernel": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_double]),
    "gpt_define_composite": (ctypes.c_int, [_vp, ctypes.c_int, _c_int32_p, _c_int32_p, ctypes.c_int, _c_int32_p]),
    "gpt_cov_pairs": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int,
                                     ctypes.c_int64, _c_double_p, _c_double_p, _c_int32_p, _c_int32_p, _c_double_p]),
    "gpt_compute_Kij": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int,
                                       ctypes.c_int, _c_double_p, _c_int32_p, ctypes.c_int, _c_double_p, _c_int32_p,
                                       _c_double_p]),
    "gpt_ll": (ctypes.c_int, [_vp, _c_double_p, ctypes.c_double, _c_double_p, _c_double_p, _c_int32_p, ctypes.c_int,
                              _c_int_p]),
    "gpt_ll_from_K": (ctypes.c_int, [_vp, _c_double_p, _c_double_p, _c_int_p]),
    "gpt_grad_from_dK": (ctypes.c_int, [_vp, _c_double_p, _c_double_p]),
    "gpt_noise_grad": (ctypes.c_int, [_vp, ctypes.c_double, _c_double_p]),
    "gpt_get_alpha": (ctypes.c_int, [_vp, _c_double_p]),
    "gpt_get_L": (ctypes.c_int, [_vp, _c_double_p]),
    "gpt_get_K": (ctypes.c_int, [_vp, _c_double_p]),
    "gpt_ll_batched": (ctypes.c_int, [_vp, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                      _c_int32_p, ctypes.c_int, _c_int_p, _c_double_p]),
    "gpt_ll_batched_dev": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _c_int32_p, ctypes.c_int, _vp, _vp]),
    "gpt_predict_batched": (ctypes.c_int, [_vp, ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, _c_double_p,
                                           _c_int32_p, _c_double_p, _c_double_p, _c_double_p, _c_int_p]),
    "gpt_predict": (ctypes.c_int, [_vp, ctypes.c_int, _c_double_p, _c_int32_p, _c_double_p, _c_double_p,
                                   _c_double_p]),
    "gpt_predict_dev": (ctypes.c_int, [_vp, ctypes.c_int, _c_double_p, _c_int32_p, _vp, _vp]),
    "gpt_predict_from_Kstar": (ctypes.c_int, [_vp, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                              _c_double_p, _c_double_p]),
    "gpt_draw_sample": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p,
                                       ctypes.c_double, _c_double_p, _c_int_p]),
    "gpt_launch_count": (ctypes.c_int64, [_vp]),
}

_lib = None


class GPTLibraryError(RuntimeError):
    """The native library is missing / no CUDA device / a CUDA call failed."""


def load_library():
    """dlopen libgptb200.so and attach the signatures.  Needs no GPU (used by the CPU test-suite to check
    that the boundary exports every declared symbol)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GPTLibraryError(
            "libgptb200.so not found at %s -- build it with `python -m gptools_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_c_int32_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def _i32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if shape is not None and a.shape != shape:
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


class Device(object):
    """One gpt_handle: a CUDA device + stream + all device-side state of one GaussianProcess."""

    def __init__(self, device=None):
        self._lib = load_library()
        if self._lib.gpt_device_count() < 1:
            raise GPTLibraryError("no CUDA device visible: gptools_b200 has no CPU fallback")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) % self._lib.gpt_device_count()
        h = _vp()
        rc = self._lib.gpt_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise GPTLibraryError("gpt_create(device=%d) failed with code %d" % (device, rc))
        self._h = h
        self.device = int(device)
        self._data_key = None
        self._kernel_key = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.gpt_last_error(self._h)
            msg = msg.decode("utf-8", "replace") if msg else ""
            if rc == -3:
                raise NotImplementedError("%s: %s" % (what, msg))
            if rc == -1:
                raise ValueError("%s: %s" % (what, msg))
            raise GPTLibraryError("%s failed (code %d): %s" % (what, rc, msg))

    # -- plumbing ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        """Issue all work on the given cudaStream_t (integer handle; 0 / None = legacy default stream)."""
        self._check(self._lib.gpt_set_stream(self._h, _vp(cuda_stream_ptr) if cuda_stream_ptr else None), "gpt_set_stream")

    def use_own_stream(self):
        self._check(self._lib.gpt_use_own_stream(self._h), "gpt_use_own_stream")

    def synchronize(self):
        self._check(self._lib.gpt_synchronize(self._h), "gpt_synchronize")

    def launch_count(self):
        return int(self._lib.gpt_launch_count(self._h))

    # -- data / kernel ------------------------------------------------------------------------
    def set_data(self, X, n, y, err_y, T=None):
        X = _f64(X)
        N, D = X.shape
        n = _i32(n, (N, D))
        y = _f64(y)
        M = y.shape[0]
        err_y = _f64(err_y, (M,))
        if T is not None:
            T = _f64(T, (M, N))
        self._check(self._lib.gpt_set_data(self._h, N, M, D, _dp(X), _ip(n), _dp(y), _dp(err_y), _dp(T)), "gpt_set_data")
        self.N, self.M, self.D = N, M, D

    def set_y(self, y):
        y = _f64(y, (self.M,))
        self._check(self._lib.gpt_set_y(self._h, _dp(y)), "gpt_set_y")

    def define_composite(self, cid):
        """Make ``cid`` (a CompositeId) the structure behind kernel id GPT_COMPOSITE on this handle (no-op when it
        already is)."""
        if getattr(self, "_composite", None) is not None and self._composite == cid:
            return
        kids, nps, masks = _i32(cid.leaf_kids), _i32(cid.leaf_nparams), _i32(cid.term_masks)
        self._check(self._lib.gpt_define_composite(self._h, len(kids), _ip(kids), _ip(nps), len(masks), _ip(masks)),
                    "gpt_define_composite")
        self._composite = cid

    def _own_structure(self):
        """Re-establish the structure of the kernel given to ``set_kernel`` (another composite may have been evaluated
        through ``cov_pairs`` / ``compute_Kij`` on this handle since)."""
        if getattr(self, "_kernel_cid", None) is not None:
            self.define_composite(self._kernel_cid)

    def _prepare_kernel_id(self, kernel_id):
        if isinstance(kernel_id, CompositeId):
            self.define_composite(kernel_id)
        return int(kernel_id)

    def set_kernel(self, kernel_id, nparams, diag_factor):
        self._prepare_kernel_id(kernel_id)
        self._kernel_cid = kernel_id if isinstance(kernel_id, CompositeId) else None
        self._check(self._lib.gpt_set_kernel(self._h, int(kernel_id), int(nparams), float(diag_factor)), "gpt_set_kernel")
        self.kernel_id, self.nparams = int(kernel_id), int(nparams)

    # -- covariance evaluation ------------------------------------------------------------------
    def cov_pairs(self, kernel_id, params, Xi, Xj, ni, nj, hyper_deriv=None):
        Xi = _f64(np.atleast_2d(Xi))
        npairs, D = Xi.shape
        Xj = _f64(np.atleast_2d(Xj), (npairs, D))
        ni = _i32(np.atleast_2d(ni), (npairs, D))
        nj = _i32(np.atleast_2d(nj), (npairs, D))
        params = _f64(params)
        out = np.empty(npairs, dtype=np.float64)
        hd = -1 if hyper_deriv is None else int(hyper_deriv)
        self._check(self._lib.gpt_cov_pairs(self._h, self._prepare_kernel_id(kernel_id), D, len(params), _dp(params), hd, npairs, _dp(Xi),
                                            _dp(Xj), _ip(ni), _ip(nj), _dp(out)), "gpt_cov_pairs")
        return out

    def compute_Kij(self, kernel_id, params, Xi, ni, Xj=None, nj=None, hyper_deriv=None):
        Xi = _f64(np.atleast_2d(Xi))
        Mi, D = Xi.shape
        ni = _i32(np.atleast_2d(ni), (Mi, D))
        if Xj is None:
            Mj = Mi
        else:
            Xj = _f64(np.atleast_2d(Xj))
            Mj = Xj.shape[0]
            nj = _i32(np.atleast_2d(nj), (Mj, D))
        params = _f64(params)
        out = np.empty((Mi, Mj), dtype=np.float64)
        hd = -1 if hyper_deriv is None else int(hyper_deriv)
        self._check(self._lib.gpt_compute_Kij(self._h, self._prepare_kernel_id(kernel_id), D, len(params), _dp(params), hd, Mi, _dp(Xi),
                                              _ip(ni), Mj, _dp(Xj), _ip(nj), _dp(out)), "gpt_compute_Kij")
        return out

    # -- likelihood -----------------------------------------------------------------------------
    def ll(self, params, noise_sigma=0.0, grad_idx=None):
        """Returns (ll, grad or None, status)."""
        self._own_structure()
        params = _f64(params, (self.nparams,))
        ll = ctypes.c_double(0.0)
        status = ctypes.c_int(0)
        if grad_idx is not None and len(grad_idx) > 0:
            gi = _i32(grad_idx)
            grad = np.zeros(len(gi), dtype=np.float64)
            rc = self._lib.gpt_ll(self._h, _dp(params), float(noise_sigma), ctypes.byref(ll), _dp(grad), _ip(gi),
                                  len(gi), ctypes.byref(status))
        else:
            grad = None
            rc = self._lib.gpt_ll(self._h, _dp(params), float(noise_sigma), ctypes.byref(ll), None, None, 0,
                                  ctypes.byref(status))
        self._check(rc, "gpt_ll")
        return ll.value, grad, status.value

    def ll_from_K(self, K_latent):
        K_latent = _f64(K_latent, (self.N, self.N))
        ll = ctypes.c_double(0.0)
        status = ctypes.c_int(0)
        self._check(self._lib.gpt_ll_from_K(self._h, _dp(K_latent), ctypes.byref(ll), ctypes.byref(status)), "gpt_ll_from_K")
        return ll.value, status.value

    def grad_from_dK(self, dK_latent):
        dK_latent = _f64(dK_latent, (self.N, self.N))
        g = ctypes.c_double(0.0)
        self._check(self._lib.gpt_grad_from_dK(self._h, _dp(dK_latent), ctypes.byref(g)), "gpt_grad_from_dK")
        return g.value

    def noise_grad(self, noise_sigma):
        g = ctypes.c_double(0.0)
        self._check(self._lib.gpt_noise_grad(self._h, float(noise_sigma), ctypes.byref(g)), "gpt_noise_grad")
        return g.value

    def get_alpha(self):
        a = np.empty(self.M, dtype=np.float64)
        self._check(self._lib.gpt_get_alpha(self._h, _dp(a)), "gpt_get_alpha")
        return a

    def get_L(self):
        L = np.empty((self.M, self.M), dtype=np.float64)
        self._check(self._lib.gpt_get_L(self._h, _dp(L)), "gpt_get_L")
        return L

    def get_K(self):
        K = np.empty((self.N, self.N), dtype=np.float64)
        self._check(self._lib.gpt_get_K(self._h, _dp(K)), "gpt_get_K")
        return K

    def ll_batched(self, thetas, grad_idx=None, y_batch=None, return_alpha=False):
        """thetas: (B, nparams + 1) = kernel params then sigma_n.  Returns (ll, grad or None, status[, alpha])."""
        self._own_structure()
        thetas = _f64(np.atleast_2d(thetas))
        B = thetas.shape[0]
        if thetas.shape[1] != self.nparams + 1:
            raise ValueError("thetas must be (B, nparams + 1)")
        ll = np.empty(B, dtype=np.float64)
        status = np.empty(B, dtype=np.int32)
        grad = gi = None
        P = 0
        if grad_idx is not None and len(grad_idx) > 0:
            gi = _i32(grad_idx)
            P = len(gi)
            grad = np.empty((B, P), dtype=np.float64)
        yb = _f64(y_batch, (B, self.M)) if y_batch is not None else None
        alpha = np.empty((B, self.M), dtype=np.float64) if return_alpha else None
        self._check(self._lib.gpt_ll_batched(self._h, B, _dp(thetas), _dp(yb), _dp(ll), _dp(grad), _ip(gi), P,
                                             status.ctypes.data_as(_c_int_p), _dp(alpha)), "gpt_ll_batched")
        if return_alpha:
            return ll, grad, status, alpha
        return ll, grad, status

    def ll_batched_dev(self, B, d_thetas, d_ll, d_status, d_grad=0, grad_idx=None, d_y_batch=0, d_alpha=0):
        """Raw device pointers (ints); nothing is copied."""
        self._own_structure()
        gi = _i32(grad_idx) if grad_idx is not None and len(grad_idx) else None
        P = 0 if gi is None else len(gi)
        self._check(self._lib.gpt_ll_batched_dev(self._h, int(B), _vp(d_thetas), _vp(d_y_batch) if d_y_batch else None,
                                                 _vp(d_ll), _vp(d_grad) if d_grad else None, _ip(gi), P, _vp(d_status),
                                                 _vp(d_alpha) if d_alpha else None), "gpt_ll_batched_dev")

    def predict_batched(self, thetas, Xs, ns, y_batch=None):
        """Predictive mean and variance at every row of ``thetas`` (B, nparams + 1) in one launch.
        Returns (mean (B, Ms), var (B, Ms), ll (B,), status (B,)); raises NotImplementedError when the library cannot
        serve the batch with its persistent kernel (transformation matrix, more than 2048 observations)."""
        self._own_structure()
        thetas = _f64(np.atleast_2d(thetas))
        B = thetas.shape[0]
        if thetas.shape[1] != self.nparams + 1:
            raise ValueError("thetas must be (B, nparams + 1)")
        Xs = _f64(np.atleast_2d(Xs))
        Ms = Xs.shape[0]
        if Xs.shape[1] != self.D:
            raise ValueError("Xs must be (Ms, %d)" % self.D)
        ns = _i32(np.atleast_2d(ns), (Ms, self.D))
        yb = _f64(y_batch, (B, self.M)) if y_batch is not None else None
        mean = np.empty((B, Ms), dtype=np.float64)
        var = np.empty((B, Ms), dtype=np.float64)
        ll = np.empty(B, dtype=np.float64)
        status = np.zeros(B, dtype=np.int32)
        self._check(self._lib.gpt_predict_batched(self._h, B, _dp(thetas), _dp(yb), Ms, _dp(Xs), _ip(ns), _dp(mean),
                                                  _dp(var), _dp(ll), status.ctypes.data_as(_c_int_p)),
                    "gpt_predict_batched")
        return mean, var, ll, status

    # -- prediction -------------------------------------------------------------------------------
    def predict(self, Xs, ns, want_var=True, want_cov=False):
        Xs = _f64(np.atleast_2d(Xs))
        Ms, D = Xs.shape
        ns = _i32(np.atleast_2d(ns), (Ms, D))
        mean = np.empty(Ms, dtype=np.float64)
        var = np.empty(Ms, dtype=np.float64) if (want_var or want_cov) else None
        cov = np.empty((Ms, Ms), dtype=np.float64) if want_cov else None
        self._check(self._lib.gpt_predict(self._h, Ms, _dp(Xs), _ip(ns), _dp(mean), _dp(var), _dp(cov)), "gpt_predict")
        return mean, var, cov

    def predict_dev(self, Xs, ns, d_mean, d_var=0):
        """Host test points, raw DEVICE output pointers (ints); asynchronous on the handle's stream."""
        Xs = _f64(np.atleast_2d(Xs))
        Ms, D = Xs.shape
        ns = _i32(np.atleast_2d(ns), (Ms, D))
        self._check(self._lib.gpt_predict_dev(self._h, Ms, _dp(Xs), _ip(ns), _vp(d_mean), _vp(d_var) if d_var else None),
                    "gpt_predict_dev")

    def predict_from_Kstar(self, Kstar, kss_diag=None, Kss=None, want_var=True, want_cov=False):
        """Host-evaluated kernels: ``Kstar`` is the (N latent x Ms) cross-covariance of gaussian_process.py:966,
        ``kss_diag`` / ``Kss`` the prior variance / covariance of the test points."""
        KstarT = _f64(np.ascontiguousarray(np.asarray(Kstar, dtype=np.float64).T))
        Ms = KstarT.shape[0]
        if KstarT.shape[1] != self.N:
            raise ValueError("Kstar must have one row per latent training point")
        kd = _f64(kss_diag, (Ms,)) if kss_diag is not None else None
        KS = _f64(Kss, (Ms, Ms)) if Kss is not None else None
        mean = np.empty(Ms, dtype=np.float64)
        var = np.empty(Ms, dtype=np.float64) if (want_var or want_cov) else None
        cov = np.empty((Ms, Ms), dtype=np.float64) if want_cov else None
        self._check(self._lib.gpt_predict_from_Kstar(self._h, Ms, _dp(KstarT), _dp(kd), _dp(KS), _dp(mean), _dp(var),
                                                     _dp(cov)), "gpt_predict_from_Kstar")
        return mean, var, cov

    def draw_sample(self, mean, cov, rand_vars, jitter):
        mean = _f64(mean)
        Ms = mean.shape[0]
        cov = _f64(cov, (Ms, Ms))
        rand_vars = _f64(np.atleast_2d(rand_vars))
        if rand_vars.shape[0] != Ms:
            raise ValueError("rand_vars must have one row per test point")
        S = rand_vars.shape[1]
        out = np.empty((Ms, S), dtype=np.float64)
        status = ctypes.c_int(0)
        self._check(self._lib.gpt_draw_sample(self._h, Ms, S, _dp(mean), _dp(cov), _dp(rand_vars), float(jitter),
                                              _dp(out), ctypes.byref(status)), "gpt_draw_sample")
        return out, status.value
