"""Noise kernels (host side).  Inside the GP they are never evaluated pairwise: ZeroKernel contributes
nothing and DiagonalNoiseKernel contributes sigma_n^2 on the diagonal of the latent covariance
(gaussian_process.py:1434-1439), which the assembly kernel fuses in.  The pairwise ``__call__`` is kept for
API compatibility (kernel/noise.py:76-152)."""
import numpy as np

from .core import Kernel

__all__ = ["DiagonalNoiseKernel", "ZeroKernel"]


class DiagonalNoiseKernel(Kernel):
    """Homoscedastic, uncorrelated noise on the derivative order ``n`` (kernel/noise.py:27-110)."""

    def __init__(self, num_dim=1, initial_noise=None, fixed_noise=False, noise_bound=None, n=0, hyperprior=None):
        try:
            iter(n)
        except TypeError:
            self.n = n * np.ones(num_dim, dtype=int)
        else:
            if len(n) != num_dim:
                raise ValueError("Length of n must be equal to num_dim!")
            self.n = np.asarray(n, dtype=int)
        super(DiagonalNoiseKernel, self).__init__(
            num_dim=num_dim, num_params=1,
            initial_params=None if initial_noise is None else [initial_noise],
            fixed_params=[True] if fixed_noise else None,
            param_bounds=None if noise_bound is None else [noise_bound],
            hyperprior=hyperprior, param_names=[r'\sigma_n'])

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        Xi = np.atleast_2d(np.asarray(Xi))
        if not symmetric:
            return np.zeros(Xi.shape[0])
        hit = ((Xi == Xj) & (ni == self.n) & (nj == self.n)).all(axis=1)
        val = self.params[0] ** 2 * np.asarray(hit, dtype=float).flatten()
        if hyper_deriv is None:
            return val
        return 2.0 * val / self.params[hyper_deriv]


class ZeroKernel(DiagonalNoiseKernel):
    """Identically zero; the default noise kernel (kernel/noise.py:113-152)."""

    def __init__(self, num_dim=1):
        super(ZeroKernel, self).__init__(num_dim=num_dim, initial_noise=0.0, fixed_noise=True)

    def __call__(self, Xi, Xj, ni, nj, hyper_deriv=None, symmetric=False):
        return np.zeros(np.atleast_2d(np.asarray(Xi)).shape[0], dtype=float)
