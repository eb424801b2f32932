"""Squared-exponential kernel with arbitrary derivative orders and hyperparameter derivatives."""
from .core import DeviceKernel

__all__ = ["SquaredExponentialKernel"]


class SquaredExponentialKernel(DeviceKernel):
    r"""k = sigma_f^2 exp(-1/2 sum_d tau_d^2 / l_d^2); params = [sigma_f, l_1, ..., l_D]
    (kernel/squared_exponential.py:31-174).  Derivative orders of any degree are handled with the
    Hermite recurrence in the device function ``se_cov`` (csrc/covfn.cuh); ``hyper_deriv`` is supported
    for every parameter."""

    kernel_id = 0
    supports_hyper_deriv = True

    def __init__(self, num_dim=1, **kwargs):
        names = [r'\sigma_f'] + ['l_{:d}'.format(i + 1) for i in range(num_dim)]
        super(SquaredExponentialKernel, self).__init__(num_dim=num_dim, num_params=num_dim + 1,
                                                       param_names=names, **kwargs)
