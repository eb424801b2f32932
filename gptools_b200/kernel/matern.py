"""Matern kernels: the fixed nu = 5/2 form and the generic-order form."""
import numpy as np

from .core import DeviceKernel

__all__ = ["Matern52Kernel", "MaternKernel"]


class Matern52Kernel(DeviceKernel):
    r"""Matern nu = 5/2, value and first derivatives only; params = [sigma_f, l_1, ..., l_D]
    (kernel/matern.py:468-555; closed forms of kernel/src/matern.c:61-186 in ``matern52_cov``).

    ``hyper_deriv`` is supported (the reference raises NotImplementedError, kernel/matern.py:543): the closed
    forms are re-evaluated in dual numbers on the device (csrc/covfn_hyper.cuh)."""

    kernel_id = 1
    supports_hyper_deriv = True

    def __init__(self, num_dim=1, **kwargs):
        names = [r'\sigma_f'] + ['l_{:d}'.format(i + 1) for i in range(num_dim)]
        super(Matern52Kernel, self).__init__(num_dim=num_dim, num_params=num_dim + 1, param_names=names, **kwargs)

    def _check_orders(self, ni, nj):
        # kernel/matern.py:545-546 (the C code would exit(1), matern.c:173-176)
        if np.any(np.sum(ni, axis=1) > 1) or np.any(np.sum(nj, axis=1) > 1):
            raise ValueError("Matern52Kernel only supports 0th and 1st order derivatives")


class MaternKernel(DeviceKernel):
    r"""Generic Matern; params = [sigma_f, nu, l_1, ..., l_D] (kernel/matern.py:251-465).

    The device function ``matern_cov`` reproduces the reference's behaviour -- including its one-term
    power series for derivative orders >= 1 when 0 < 2 nu r^2 <= 5e-4 (utils.py:1493-1516), its averaging over
    nu -+ 0.001 there for integer nu (utils.py:1480-1484, 1498-1502) and the origin limits
    (kernel/matern.py:444-457) -- for any nu > 0 (K_nu of real order by Temme's method; the closed form for
    half-integer nu) and total derivative order <= 2 per pair, which covers value + first-derivative observations
    and predictions.  ``hyper_deriv`` is supported for sigma_f and the length scales (dual numbers,
    csrc/covfn_hyper.cuh) and -- by Richardson-extrapolated central differences of the device evaluation -- for nu (the
    reference has no hyper-derivatives at all here)."""

    kernel_id = 2
    supports_hyper_deriv = True

    #: d/dnu has no closed form (it needs dK_nu/dnu): Richardson-extrapolated central differences of the device's own
    #: evaluation (of K in ``__call__``, of ll in ``GaussianProcess.compute_K_L_alpha_ll``), relative error ~1e-8
    fd_hyper_idxs = (1,)

    def __init__(self, num_dim=1, **kwargs):
        names = [r'\sigma_f', r'\nu'] + ['l_{:d}'.format(i + 1) for i in range(num_dim)]
        super(MaternKernel, self).__init__(num_dim=num_dim, num_params=num_dim + 2, param_names=names, **kwargs)

    @property
    def nu(self):
        return self.params[1]

    def _check_params_for_device(self):
        nu = float(self.params[1])
        if not nu > 0:
            raise ValueError("MaternKernel needs nu > 0; got nu = %r" % nu)

    def batch_rows_supported(self, param_rows):
        with np.errstate(invalid="ignore"):
            return np.atleast_2d(param_rows)[:, 1] > 0

    def _check_orders(self, ni, nj):
        ti, tj = np.sum(ni, axis=1), np.sum(nj, axis=1)
        too_high = np.any(ti + tj > 2) if len(ti) == len(tj) else (ti.max() + tj.max() > 2)
        if too_high:
            raise NotImplementedError("MaternKernel on the device supports a total derivative order <= 2 per pair")
