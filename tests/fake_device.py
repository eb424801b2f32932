"""Test double for gptools_b200._lib.Device backed by the numpy oracle.

Lets the CPU test-suite (-m "not gpu") exercise every line of the HOST logic of GaussianProcess -- data
bookkeeping, hyperparameter plumbing, priors, mean functions, batched entry, drivers -- without a GPU.  It is
test infrastructure only (it imports oracle/, which the product never does) and is injected by monkeypatching
``gp._dev_obj``; nothing in gptools_b200 can reach it."""
import numpy as np

from oracle import gp_oracle as orc


class FakeDevice(object):
    def __init__(self):
        self.calls = []
        self._state = None

    def set_data(self, X, n, y, err_y, T=None):
        self.X, self.n, self.y, self.err_y, self.T = (np.array(X, float), np.array(n, int), np.array(y, float),
                                                      np.array(err_y, float), None if T is None else np.array(T, float))
        self.N, self.D = self.X.shape
        self.M = len(self.y)
        self.calls.append("set_data")

    def set_y(self, y):
        self.y = np.array(y, float)
        self.calls.append("set_y")

    def set_kernel(self, kid, nparams, diag_factor):
        self.kernel_id, self.nparams, self.diag_factor = kid, nparams, diag_factor
        self.calls.append("set_kernel")

    def cov_pairs(self, kid, params, Xi, Xj, ni, nj, hyper_deriv=None):
        return orc.kernel_pairs(kid, params, Xi, Xj, ni, nj, hyper_deriv=hyper_deriv)

    def compute_Kij(self, kid, params, Xi, ni, Xj=None, nj=None, hyper_deriv=None):
        return orc.compute_Kij(kid, params, Xi, Xj, ni, nj, hyper_deriv=hyper_deriv)

    def _run(self, params, noise_sigma, grad_idx, y=None):
        gi = None
        noise_slot = None
        if grad_idx is not None:
            gi = [g for g in grad_idx if g < self.nparams]
            if self.nparams in list(grad_idx):
                noise_slot = list(grad_idx).index(self.nparams)
        try:
            r = orc.compute_K_L_alpha_ll(self.kernel_id, params, self.X, self.n, self.y if y is None else y,
                                         self.err_y, T=self.T, noise_sigma=noise_sigma, diag_factor=self.diag_factor,
                                         grad_idx=gi if gi else None)
        except (np.linalg.LinAlgError, ValueError):   # not positive definite / non-finite entries: a potrf status on the device
            return None
        grad = None
        if grad_idx is not None:
            grad = np.zeros(len(grad_idx))
            k = 0
            for q, g in enumerate(grad_idx):
                if g < self.nparams:
                    grad[q] = r["ll_deriv"][k]
                    k += 1
            if noise_slot is not None:
                a = r["alpha"].ravel()
                Kinv = np.linalg.inv(r["K_tot"])
                grad[noise_slot] = noise_sigma * (a.dot(a) - np.trace(Kinv))
        return r, grad

    def ll(self, params, noise_sigma=0.0, grad_idx=None):
        self.calls.append("ll")
        out = self._run(params, noise_sigma, grad_idx)
        if out is None:
            return 0.0, (np.zeros(len(grad_idx)) if grad_idx is not None else None), 1
        self._state, grad = out
        self._params = np.array(params, float)
        return self._state["ll"], grad, 0

    def get_alpha(self):
        return self._state["alpha"].ravel().copy()

    def get_L(self):
        return self._state["L"].copy()

    def get_K(self):
        return self._state["K"].copy()

    def ll_batched(self, thetas, grad_idx=None, y_batch=None, return_alpha=False):
        self.calls.append("ll_batched")
        thetas = np.atleast_2d(thetas)
        B = len(thetas)
        ll = np.zeros(B)
        st = np.zeros(B, dtype=np.int32)
        grad = np.zeros((B, len(grad_idx))) if grad_idx is not None and len(grad_idx) else None
        alpha = np.zeros((B, self.M))
        for b in range(B):
            out = self._run(thetas[b, :-1], thetas[b, -1], grad_idx if grad is not None else None,
                            y=None if y_batch is None else y_batch[b])
            if out is None:
                st[b] = 1
                continue
            ll[b] = out[0]["ll"]
            alpha[b] = out[0]["alpha"].ravel()
            if grad is not None:
                grad[b] = out[1]
        return (ll, grad, st, alpha) if return_alpha else (ll, grad, st)

    def predict_batched(self, thetas, Xs, ns, y_batch=None):
        self.calls.append("predict_batched")
        thetas = np.atleast_2d(thetas)
        B, Ms = len(thetas), len(Xs)
        mean, var = np.zeros((B, Ms)), np.zeros((B, Ms))
        ll, st = np.zeros(B), np.zeros(B, dtype=np.int32)
        for b in range(B):
            out = self._run(thetas[b, :-1], thetas[b, -1], None, y=None if y_batch is None else y_batch[b])
            if out is None:
                st[b] = 1
                continue
            r = out[0]
            m, _, cov = orc.predict(self.kernel_id, thetas[b, :-1], self.X, self.n, r["L"], r["alpha"], Xs, ns, T=self.T)
            mean[b], var[b], ll[b] = m, np.diag(cov), r["ll"]
        return mean, var, ll, st

    def predict(self, Xs, ns, want_var=True, want_cov=False):
        self.calls.append("predict")
        mean, std, cov = orc.predict(self.kernel_id, self._params, self.X, self.n, self._state["L"],
                                     self._state["alpha"], Xs, ns, T=self.T)
        return mean, (np.diag(cov).copy() if (want_var or want_cov) else None), (cov if want_cov else None)

    def draw_sample(self, mean, cov, rand_vars, jitter):
        self.calls.append("draw_sample")
        return orc.draw_sample(mean, cov, rand_vars, diag_factor=jitter / orc.EPS), 0
