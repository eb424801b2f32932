"""Gibbs (non-stationary squared-exponential) kernel in 1-D with a tanh length-scale profile."""
import numpy as np

from .core import DeviceKernel

__all__ = ["tanh_warp", "double_tanh_warp", "cubic_bucket_warp", "quintic_bucket_warp", "exp_gauss_warp",
           "GibbsKernel1d", "GibbsKernel1dTanh", "GibbsKernel1dDoubleTanh", "GibbsKernel1dCubicBucket",
           "GibbsKernel1dQuinticBucket", "GibbsKernel1dExpGauss"]


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


# ------------------------------------------------------------------------------------------------------------------
# Other length-scale profiles (kernel/gibbs.py:508-902).  They are evaluated on the HOST, once per point: the device
# kernel GPT_GIBBS_AUX takes l(x) and l'(x) as two extra columns of the point array, so assembly, factorisation,
# prediction and sampling run through exactly the same CUDA path as every other kernel.
# ------------------------------------------------------------------------------------------------------------------
def _order(n):
    if n not in (0, 1):
        raise NotImplementedError("Only derivatives up to order 1 are supported!")
    return n


def double_tanh_warp(x, n, lcore, lmid, ledge, la, lb, xa, xb):
    r"""Sum of two tanh steps, l = a tanh((x-xa)/la) + b tanh((x-xb)/lb) + c with the plateaus lcore (x << xa),
    lmid (xa << x << xb) and ledge (x >> xb) (kernel/gibbs.py:508-557)."""
    x = np.asarray(x, dtype=float)
    a, b, c = 0.5 * (lmid - lcore), 0.5 * (ledge - lmid), 0.5 * (lcore + ledge)
    ta, tb = np.tanh((x - xa) / la), np.tanh((x - xb) / lb)
    if _order(n) == 0:
        return a * ta + b * tb + c
    return a / la * (1.0 - ta * ta) + b / lb * (1.0 - tb * tb)


def _bucket_edges(x0, w1, w2, w3):
    x1 = x0 - 0.5 * w2 - 0.5 * w1     # centre of the left transition
    x2 = x0 + 0.5 * w2 + 0.5 * w3     # centre of the right transition
    return x1, x2


def cubic_bucket_warp(x, n, l1, l2, l3, x0, w1, w2, w3):
    r""""Bucket" profile: l1 | cubic transition of width w1 | l2 over width w2 centred on x0 | cubic transition of
    width w3 | l3 (kernel/gibbs.py:603-651).  The transitions are the smoothstep 3 s^2 - 2 s^3."""
    x = np.asarray(x, dtype=float)
    x1, x2 = _bucket_edges(x0, w1, w2, w3)
    s1, s2 = (x - x1) / w1 + 0.5, (x - x2) / w3 + 0.5
    in1 = (x > x1 - 0.5 * w1) & (x < x1 + 0.5 * w1)
    in2 = (x > x2 - 0.5 * w3) & (x < x2 + 0.5 * w3)
    if _order(n) == 0:
        out = np.where(x <= x1 - 0.5 * w1, l1, 0.0)
        out = out + np.where(in1, l1 + (l2 - l1) * (3.0 * s1 ** 2 - 2.0 * s1 ** 3), 0.0)
        out = out + np.where((x >= x1 + 0.5 * w1) & (x <= x2 - 0.5 * w3), l2, 0.0)
        out = out + np.where(in2, l2 + (l3 - l2) * (3.0 * s2 ** 2 - 2.0 * s2 ** 3), 0.0)
        return out + np.where(x >= x2 + 0.5 * w3, l3, 0.0)
    return (np.where(in1, (l2 - l1) * 6.0 * (s1 - s1 ** 2) / w1, 0.0) +
            np.where(in2, (l3 - l2) * 6.0 * (s2 - s2 ** 2) / w3, 0.0))


def quintic_bucket_warp(x, n, l1, l2, l3, x0, w1, w2, w3):
    r"""Bucket profile with quintic transitions (continuous second derivative), kernel/gibbs.py:695-760:
    on a transition, with u = 2 (x - centre) / width in (-1, 1), l = mean + half-step * (3/8 u^5 - 5/4 u^3 + 15/8 u)."""
    x = np.asarray(x, dtype=float)
    x1, x2 = _bucket_edges(x0, w1, w2, w3)
    u1, u2 = 2.0 * (x - x1) / w1, 2.0 * (x - x2) / w3
    in1 = (x > x1 - 0.5 * w1) & (x < x1 + 0.5 * w1)
    in2 = (x > x2 - 0.5 * w3) & (x < x2 + 0.5 * w3)
    poly = lambda u: 0.375 * u ** 5 - 1.25 * u ** 3 + 1.875 * u
    dpoly = lambda u: 1.875 * u ** 4 - 3.75 * u ** 2 + 1.875
    if _order(n) == 0:
        out = np.where(x <= x1 - 0.5 * w1, l1, 0.0)
        out = out + np.where(in1, 0.5 * (l2 - l1) * poly(u1) + 0.5 * (l1 + l2), 0.0)
        out = out + np.where((x >= x1 + 0.5 * w1) & (x <= x2 - 0.5 * w3), l2, 0.0)
        out = out + np.where(in2, 0.5 * (l3 - l2) * poly(u2) + 0.5 * (l2 + l3), 0.0)
        return out + np.where(x >= x2 + 0.5 * w3, l3, 0.0)
    return (np.where(in1, 0.5 * (l2 - l1) * dpoly(u1) / w1, 0.0) +
            np.where(in2, 0.5 * (l3 - l2) * dpoly(u2) / w3, 0.0))


def exp_gauss_warp(X, n, l0, *msb):
    r"""l = l0 exp(sum_i b_i exp(-(x - m_i)^2 / (2 s_i^2))); ``msb`` holds the means, then the standard deviations,
    then the weights (kernel/gibbs.py:804-855)."""
    X = np.asarray(X, dtype=float)
    msb = np.asarray(msb, dtype=float)
    ng = len(msb) // 3
    mm, ss, bb = msb[:ng], msb[ng:2 * ng], msb[2 * ng:]
    terms = [b * np.exp(-(X - m) ** 2 / (2.0 * s * s)) for m, s, b in zip(mm, ss, bb)]
    expo = sum(terms) if terms else np.zeros_like(X)
    if _order(n) == 0:
        return l0 * np.exp(expo)
    slope = sum(t * (X - m) / (s * s) for t, m, s in zip(terms, mm, ss)) if terms else np.zeros_like(X)
    return -l0 * np.exp(expo) * slope


class GibbsKernel1d(DeviceKernel):
    r"""Gibbs kernel in 1-D with an arbitrary length-scale function ``l_func(x, n, *params)`` for n = 0, 1
    (kernel/gibbs.py:244-424); params = [sigma_f, *l_func parameters].  Derivative orders up to (1, 1).

    ``l_func`` is evaluated on the host at every point; the covariance itself is the device closed form
    (``gibbs_cov_l`` in csrc/covfn.cuh, kernel id GPT_GIBBS_AUX) fed with (x, l(x), l'(x)) per point.  Like the
    reference, hyperparameter derivatives are not available (kernel/gibbs.py:319), and because l(x) changes with
    the hyperparameters the many-theta batched kernel does not apply: batches are evaluated theta by theta."""

    kernel_id = 4
    supports_hyper_deriv = False

    def __init__(self, l_func, num_params=None, **kwargs):
        self.l_func = l_func
        if kwargs.get('num_dim', 1) != 1:
            raise ValueError("Gibbs kernel only supports 1d data.")
        kwargs.pop('num_dim', None)
        if num_params is None:
            import inspect
            try:
                num_params = len(inspect.getfullargspec(l_func)[0]) - 2 + 1   # minus (x, n), plus sigma_f
            except TypeError:
                num_params = len(inspect.getfullargspec(l_func.__call__)[0]) - 3 + 1
        super(GibbsKernel1d, self).__init__(num_dim=1, num_params=num_params, **kwargs)

    def _check_orders(self, ni, nj):
        if np.any(np.asarray(ni)[:, :1] > 1) or np.any(np.asarray(nj)[:, :1] > 1):
            raise NotImplementedError("Derivatives greater than [1, 1] are not supported!")

    def device_descriptor(self):
        return (self.kernel_id, np.array(self.params[:1], dtype=float))

    def device_points(self, X, n):
        """(x, l(x), l'(x)) per point and the matching order array (orders apply to column 0 only)."""
        X = np.atleast_2d(np.asarray(X, dtype=float))
        n = np.atleast_2d(np.asarray(n, dtype=int))
        x = X[:, 0]
        lp = list(self.params[1:])
        Xa = np.column_stack([x, np.asarray(self.l_func(x, 0, *lp), dtype=float) * np.ones_like(x),
                              np.asarray(self.l_func(x, 1, *lp), dtype=float) * np.ones_like(x)])
        na = np.column_stack([n[:, 0], np.zeros(len(x), dtype=int), np.zeros(len(x), dtype=int)])
        return Xa, na

    def device_points_key(self):
        return tuple(float(v) for v in self.params[1:])


class GibbsKernel1dDoubleTanh(GibbsKernel1d):
    """params = [sigma_f, lcore, lmid, ledge, la, lb, xa, xb] (kernel/gibbs.py:560-600)."""

    def __init__(self, **kwargs):
        super(GibbsKernel1dDoubleTanh, self).__init__(
            double_tanh_warp, num_params=8,
            param_names=[r'\sigma_f', 'l_c', 'l_m', 'l_e', 'l_a', 'l_b', 'x_a', 'x_b'], **kwargs)


class GibbsKernel1dCubicBucket(GibbsKernel1d):
    """params = [sigma_f, l1, l2, l3, x0, w1, w2, w3] (kernel/gibbs.py:654-692)."""

    def __init__(self, **kwargs):
        super(GibbsKernel1dCubicBucket, self).__init__(
            cubic_bucket_warp, num_params=8,
            param_names=[r'\sigma_f', 'l_1', 'l_2', 'l_3', 'x_0', 'w_1', 'w_2', 'w_3'], **kwargs)


class GibbsKernel1dQuinticBucket(GibbsKernel1d):
    """params = [sigma_f, l1, l2, l3, x0, w1, w2, w3] (kernel/gibbs.py:763-801)."""

    def __init__(self, **kwargs):
        super(GibbsKernel1dQuinticBucket, self).__init__(
            quintic_bucket_warp, num_params=8,
            param_names=[r'\sigma_f', 'l_1', 'l_2', 'l_3', 'x_0', 'w_1', 'w_2', 'w_3'], **kwargs)


class GibbsKernel1dExpGauss(GibbsKernel1d):
    """params = [sigma_f, l0, m_1..m_G, s_1..s_G, b_1..b_G] (kernel/gibbs.py:858-902)."""

    def __init__(self, n_gaussians, **kwargs):
        self.n_gaussians = int(n_gaussians)
        names = ([r'\sigma_f', 'l_0'] + ['\\mu_{:d}'.format(i + 1) for i in range(self.n_gaussians)] +
                 ['\\sigma_{:d}'.format(i + 1) for i in range(self.n_gaussians)] +
                 ['\\beta_{:d}'.format(i + 1) for i in range(self.n_gaussians)])
        super(GibbsKernel1dExpGauss, self).__init__(exp_gauss_warp, num_params=2 + 3 * self.n_gaussians,
                                                    param_names=names, **kwargs)
