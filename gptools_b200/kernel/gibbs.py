"""Gibbs (non-stationary squared-exponential) kernel in 1-D with a tanh length-scale profile."""
import numpy as np

from .core import DeviceKernel

__all__ = ["tanh_warp", "GibbsKernel1dTanh"]


def tanh_warp(x, n, l1, l2, lw, x0):
    r"""l(x) = (l1 + l2)/2 - (l1 - l2)/2 tanh((x - x0)/lw) and its first derivative (kernel/gibbs.py:426-465).
    Host helper for inspecting the length-scale profile; the device evaluates it in ``gibbs_tanh_l``."""
    if n == 0:
        return (l1 + l2) / 2.0 - (l1 - l2) / 2.0 * np.tanh((x - x0) / lw)
    if n == 1:
        return -(l1 - l2) / (2.0 * lw) * (np.cosh((x - x0) / lw)) ** (-2.0)
    raise NotImplementedError("Only derivatives up to order 1 are supported!")


class GibbsKernel1dTanh(DeviceKernel):
    r"""k = sigma_f^2 sqrt(2 l(x) l(x') / (l(x)^2 + l(x')^2)) exp(-(x - x')^2 / (l(x)^2 + l(x')^2));
    params = [sigma_f, l1, l2, lw, x0] (kernel/gibbs.py:244-505).  Derivative orders up to (1, 1).
    ``hyper_deriv`` is supported for all five parameters (the reference raises NotImplementedError,
    kernel/gibbs.py:319): dual-number closed forms on the device (csrc/covfn_hyper.cuh)."""

    kernel_id = 3
    supports_hyper_deriv = True

    def __init__(self, **kwargs):
        if kwargs.get('num_dim', 1) != 1:
            raise ValueError("Gibbs kernel only supports 1d data.")
        kwargs.pop('num_dim', None)
        super(GibbsKernel1dTanh, self).__init__(num_dim=1, num_params=5,
                                                param_names=[r'\sigma_f', 'l_1', 'l_2', 'l_w', 'x_0'], **kwargs)
        self.l_func = tanh_warp

    def _check_orders(self, ni, nj):
        if np.any(ni > 1) or np.any(nj > 1):
            raise NotImplementedError("Derivatives greater than [1, 1] are not supported!")
