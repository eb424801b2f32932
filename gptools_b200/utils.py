"""Hyperpriors and bound/parameter views (host side).

These are the pieces of the reference's ``gptools/utils.py`` that the likelihood path consumes:
``ll`` includes ``hyperprior(params)`` (gaussian_process.py:1469) and its derivative (:1516-1520), and
the optimizer / sampler draw their starting points from ``hyperprior.random_draw``.  They are O(P)
host arithmetic evaluated once per theta and stay in Python; semantics follow utils.py:53-1267
(class and method names, argument meaning, return shapes).
"""
import numpy as np
import scipy.special
import scipy.stats

__all__ = [
    "JointPrior", "CombinedBounds", "MaskedBounds", "ProductJointPrior", "UniformJointPrior",
    "IndependentJointPrior", "NormalJointPrior", "LogNormalJointPrior", "GammaJointPrior",
    "GammaJointPriorAlt", "SortedUniformJointPrior", "unique_rows",
]


class JointPrior(object):
    """Joint prior over hyperparameters (utils.py:53-135).

    Subclasses provide ``__call__(theta, hyper_deriv=None)`` (log-pdf or its derivative with respect
    to ``theta[hyper_deriv]``), ``random_draw(size)``, ``sample_u(q)``, ``elementwise_cdf(p)`` and a
    ``bounds`` attribute.  ``p1 * p2`` concatenates two priors.
    """

    def __init__(self, i=1.0):
        # utils.py:62-71 stores 1.0 regardless of the argument; kept.
        self.i = 1.0

    def __call__(self, theta, hyper_deriv=None):
        raise NotImplementedError("__call__ must be implemented in your own class.")

    def random_draw(self, size=None):
        raise NotImplementedError("random_draw must be implemented in your own class.")

    def sample_u(self, q):
        raise NotImplementedError("ppf must be implemented in your own class.")

    def elementwise_cdf(self, p):
        raise NotImplementedError("cdf must be implemented in your own class.")

    def __mul__(self, other):
        return ProductJointPrior(self, other)

    # batched evaluation used by the many-theta entry points (not in the reference); subclasses vectorise
    def logpdf_batch(self, thetas):
        """log-pdf of every row of ``thetas`` (B, num_params) -> (B,)."""
        return np.array([self(t) for t in np.atleast_2d(thetas)], dtype=float)

    def dlogpdf_batch(self, thetas, hyper_deriv):
        """d log-pdf / d theta[hyper_deriv] for every row -> (B,)."""
        return np.array([self(t, hyper_deriv=hyper_deriv) for t in np.atleast_2d(thetas)], dtype=float)


class CombinedBounds(object):
    """Concatenated, write-through view of two sequences (utils.py:137-191)."""

    def __init__(self, l1, l2):
        self.l1 = l1
        self.l2 = l2

    def __getitem__(self, pos):
        return (list(self.l1) + list(self.l2))[pos]

    def __setitem__(self, pos, value):
        n1 = len(self.l1)
        if pos < n1:
            self.l1[pos] = value
        else:
            self.l2[pos - n1] = value

    def __len__(self):
        return len(self.l1) + len(self.l2)

    def __invert__(self):
        return ~np.asarray(self[:])

    def __iter__(self):
        return iter(self[:])

    def __str__(self):
        return str(self[:])

    def __repr__(self):
        return "%s from CombinedBounds(%s, %s)" % (self, self.l1, self.l2)


class MaskedBounds(object):
    """Write-through view of the elements ``a[m]`` (utils.py:193-230)."""

    def __init__(self, a, m):
        self.a = a
        self.m = m

    def __getitem__(self, pos):
        idx = self.m[pos]
        if np.ndim(idx) == 0:
            return self.a[idx]
        try:
            return self.a[idx]
        except TypeError:  # plain lists (e.g. bounds stored as list of tuples)
            return [self.a[i] for i in idx]

    def __setitem__(self, pos, value):
        idx = self.m[pos]
        if np.ndim(idx) == 0:
            self.a[idx] = value
        else:
            for i, v in zip(idx, value):
                self.a[i] = v

    def __len__(self):
        return len(self.m)

    def __iter__(self):
        return iter(self[:])

    def __str__(self):
        return str(self[:])

    def __repr__(self):
        return "%s from MaskedBounds(%s, %s)" % (self, self.a, self.m)


class ProductJointPrior(JointPrior):
    """Two independent priors side by side (utils.py:232-349)."""

    def __init__(self, p1, p2):
        if not isinstance(p1, JointPrior) or not isinstance(p2, JointPrior):
            raise TypeError("Both arguments to ProductPrior must be instances of JointPrior!")
        self.p1 = p1
        self.p2 = p2

    @property
    def i(self):
        return min(self.p1.i, self.p2.i)

    @i.setter
    def i(self, v):
        self.p1.i = v
        self.p2.i = v

    @property
    def bounds(self):
        return CombinedBounds(self.p1.bounds, self.p2.bounds)

    @bounds.setter
    def bounds(self, v):
        n1 = len(self.p1.bounds)
        self.p1.bounds = v[:n1]
        self.p2.bounds = v[n1:]

    def _split(self, v):
        n1 = len(self.p1.bounds)
        return v[:n1], v[n1:], n1

    def __call__(self, theta, hyper_deriv=None):
        t1, t2, n1 = self._split(theta)
        if hyper_deriv is not None:
            if hyper_deriv < n1:
                return self.p1(t1, hyper_deriv=hyper_deriv)
            return self.p2(t2, hyper_deriv=hyper_deriv - n1)
        return self.p1(t1) + self.p2(t2)

    def logpdf_batch(self, thetas):
        thetas = np.atleast_2d(thetas)
        n1 = len(self.p1.bounds)
        return self.p1.logpdf_batch(thetas[:, :n1]) + self.p2.logpdf_batch(thetas[:, n1:])

    def dlogpdf_batch(self, thetas, hyper_deriv):
        thetas = np.atleast_2d(thetas)
        n1 = len(self.p1.bounds)
        if hyper_deriv < n1:
            return self.p1.dlogpdf_batch(thetas[:, :n1], hyper_deriv)
        return self.p2.dlogpdf_batch(thetas[:, n1:], hyper_deriv - n1)

    def sample_u(self, q):
        q1, q2, _ = self._split(q)
        return np.concatenate((self.p1.sample_u(q1), self.p2.sample_u(q2)))

    def elementwise_cdf(self, p):
        a, b, _ = self._split(p)
        return np.concatenate((self.p1.elementwise_cdf(a), self.p2.elementwise_cdf(b)))

    def random_draw(self, size=None):
        d1 = self.p1.random_draw(size=size)
        d2 = self.p2.random_draw(size=size)
        if d1.ndim == 1:
            return np.hstack((d1, d2))
        return np.vstack((d1, d2))


def _check_unit_vector(q, n, name="q"):
    q = np.atleast_1d(q)
    if len(q) != n:
        raise ValueError("length of %s must equal the number of parameters!" % name)
    if q.ndim != 1:
        raise ValueError("%s must be one-dimensional!" % name)
    return q


class UniformJointPrior(JointPrior):
    """Independent uniform priors on ``bounds`` (utils.py:351-456)."""

    def __init__(self, bounds, ub=None, **kwargs):
        super(UniformJointPrior, self).__init__(**kwargs)
        if ub is not None:
            try:
                bounds = list(zip(bounds, ub))
            except TypeError:
                bounds = [(bounds, ub)]
        self.bounds = list(bounds) if not isinstance(bounds, (CombinedBounds, MaskedBounds)) else bounds

    def __call__(self, theta, hyper_deriv=None):
        if hyper_deriv is not None:
            return 0.0
        ll = 0.0
        for v, b in zip(theta, self.bounds):
            if b[0] <= v and v <= b[1]:
                ll += -np.log(b[1] - b[0])
            else:
                return -np.inf
        return ll

    def logpdf_batch(self, thetas):
        thetas = np.atleast_2d(thetas)
        nb = min(thetas.shape[1], len(self.bounds))
        if nb == 0:
            return np.zeros(thetas.shape[0])
        lo = [float(self.bounds[i][0]) for i in range(nb)]
        hi = [float(self.bounds[i][1]) for i in range(nb)]
        # column by column: a broadcast over a last axis of length <= 10 costs several times as much in numpy
        inside = None
        ll = 0.0
        for i in range(nb):
            col = thetas[:, i]
            ok = (col >= lo[i]) & (col <= hi[i])
            inside = ok if inside is None else (inside & ok)
            ll += -np.log(hi[i] - lo[i])          # same accumulation order as the scalar call
        out = np.full(thetas.shape[0], ll)
        out[~inside] = -np.inf
        return out

    def dlogpdf_batch(self, thetas, hyper_deriv):
        return np.zeros(np.atleast_2d(thetas).shape[0])

    def sample_u(self, q):
        q = _check_unit_vector(q, len(self.bounds))
        if (q < 0).any() or (q > 1).any():
            raise ValueError("q must be within [0, 1]!")
        return np.asarray([(b[1] - b[0]) * v + b[0] for v, b in zip(q, self.bounds)])

    def elementwise_cdf(self, p):
        p = _check_unit_vector(p, len(self.bounds), "p")
        lo = np.array([b[0] for b in self.bounds], dtype=float)
        hi = np.array([b[1] for b in self.bounds], dtype=float)
        return np.clip((p - lo) / (hi - lo), 0.0, 1.0)

    def random_draw(self, size=None):
        return np.asarray([np.random.uniform(low=b[0], high=b[1], size=size) for b in self.bounds])


class IndependentJointPrior(JointPrior):
    """Independent univariate priors given as scipy.stats frozen distributions or callables
    (utils.py:664-764)."""

    def __init__(self, univariate_priors, **kwargs):
        super(IndependentJointPrior, self).__init__(**kwargs)
        self.univariate_priors = univariate_priors

    def __call__(self, theta, hyper_deriv=None):
        if hyper_deriv is not None:
            raise NotImplementedError("Hyperparameter derivatives not supported for IndependentJointPrior!")
        ll = 0
        for v, p in zip(theta, self.univariate_priors):
            try:
                ll += p(theta)
            except TypeError:
                ll += p.logpdf(v)
        return ll

    @property
    def bounds(self):
        return [p.interval(self.i) for p in self.univariate_priors]

    def sample_u(self, q):
        q = _check_unit_vector(q, len(self.univariate_priors))
        if (q < 0).any() or (q > 1).any():
            raise ValueError("q must be within [0, 1]!")
        return np.asarray([p.ppf(v) for v, p in zip(q, self.univariate_priors)])

    def elementwise_cdf(self, p):
        p = _check_unit_vector(p, len(self.univariate_priors), "p")
        return np.asarray([pr.cdf(v) for v, pr in zip(p, self.univariate_priors)])

    def random_draw(self, size=None):
        return np.asarray([p.rvs(size=size) for p in self.univariate_priors])


class _StatsPrior(JointPrior):
    """Shared machinery for priors that are products of one scipy.stats family."""

    def _dists(self):
        raise NotImplementedError

    def _nvar(self):
        return len(self._dists())

    def __call__(self, theta, hyper_deriv=None):
        if hyper_deriv is not None:
            return self._dlogpdf(theta, hyper_deriv)
        ll = 0
        for v, d in zip(theta, self._dists()):
            ll += d.logpdf(v)
        return ll

    def logpdf_batch(self, thetas):
        thetas = np.atleast_2d(thetas)
        ll = 0
        for j, d in enumerate(self._dists()):
            ll = ll + d.logpdf(thetas[:, j])
        return np.asarray(ll, dtype=float) * np.ones(thetas.shape[0])

    @property
    def bounds(self):
        return [d.interval(self.i) for d in self._dists()]

    def sample_u(self, q):
        q = _check_unit_vector(q, self._nvar())
        if (q < 0).any() or (q > 1).any():
            raise ValueError("q must be within [0, 1]!")
        return np.asarray([d.ppf(v) for v, d in zip(q, self._dists())])

    def elementwise_cdf(self, p):
        p = _check_unit_vector(p, self._nvar(), "p")
        return np.asarray([d.cdf(v) for v, d in zip(p, self._dists())])

    def random_draw(self, size=None):
        return np.asarray([d.rvs(size=size) for d in self._dists()])


def _pair_1d(a, b, na, nb):
    a = np.atleast_1d(np.asarray(a, dtype=float))
    b = np.atleast_1d(np.asarray(b, dtype=float))
    if a.shape != b.shape:
        raise ValueError("%s and %s must have the same shape!" % (na, nb))
    if a.ndim != 1:
        raise ValueError("%s and %s must both be one dimensional!" % (na, nb))
    return a, b


class NormalJointPrior(_StatsPrior):
    """Independent normal priors (utils.py:766-867)."""

    def __init__(self, mu, sigma, **kwargs):
        super(NormalJointPrior, self).__init__(**kwargs)
        self.sigma, self.mu = _pair_1d(sigma, mu, "sigma", "mu")

    def _dists(self):
        return [scipy.stats.norm(loc=m, scale=s) for s, m in zip(self.sigma, self.mu)]

    def _dlogpdf(self, theta, k):
        return (self.mu[k] - theta[k]) / self.sigma[k] ** 2.0


class LogNormalJointPrior(_StatsPrior):
    """Independent log-normal priors; ``mu``, ``sigma`` are the parameters of log(theta)
    (utils.py:869-973)."""

    def __init__(self, mu, sigma, **kwargs):
        super(LogNormalJointPrior, self).__init__(**kwargs)
        self.sigma, mu = _pair_1d(sigma, mu, "sigma", "mu")
        self.emu = np.exp(mu)

    def _dists(self):
        return [scipy.stats.lognorm(s, loc=0, scale=em) for s, em in zip(self.sigma, self.emu)]

    def _dlogpdf(self, theta, k):
        return -1.0 / theta[k] * (1.0 + np.log(theta[k] / self.emu[k]) / self.sigma[k] ** 2.0)


class GammaJointPrior(_StatsPrior):
    """Independent gamma priors with shape ``a`` and rate ``b`` (utils.py:975-1079)."""

    def __init__(self, a, b, **kwargs):
        super(GammaJointPrior, self).__init__(**kwargs)
        self.a, self.b = _pair_1d(a, b, "a", "b")

    def _dists(self):
        return [scipy.stats.gamma(a, loc=0, scale=1.0 / b) for a, b in zip(self.a, self.b)]

    def _dlogpdf(self, theta, k):
        if self.a[k] == 1.0 and theta[k] == 0.0:
            return -self.b[k]
        return (self.a[k] - 1.0) / theta[k] - self.b[k]


class GammaJointPriorAlt(GammaJointPrior):
    """Gamma priors parameterised by mode ``m`` and standard deviation ``s`` (utils.py:1081-1111)."""

    def __init__(self, m, s, i=1.0):
        self.i = i
        self.m, self.s = _pair_1d(m, s, "mu", "s")

    @property
    def a(self):
        return 1.0 + self.b * self.m

    @property
    def b(self):
        return (self.m + np.sqrt(self.m ** 2 + 4.0 * self.s ** 2)) / (2.0 * self.s ** 2)


class SortedUniformJointPrior(JointPrior):
    """Uniform prior on sorted variables in [lb, ub] (utils.py:1113-1267)."""

    def __init__(self, num_var, lb, ub, **kwargs):
        super(SortedUniformJointPrior, self).__init__(**kwargs)
        self.num_var = num_var
        self.lb = lb
        self.ub = ub

    def __call__(self, theta, hyper_deriv=None):
        if hyper_deriv is not None:
            return 0.0
        theta = np.asarray(theta)
        # utils.py:1146: rejects only if EVERY element differs from its sorted position (kept as is)
        if (np.sort(theta) != theta).all() or (theta < self.lb).any() or (theta > self.ub).any():
            return -np.inf
        return np.log(scipy.special.factorial(self.num_var)) - self.num_var * np.log(self.ub - self.lb)

    @property
    def bounds(self):
        return [(self.lb, self.ub)] * self.num_var

    def sample_u(self, q):
        q = _check_unit_vector(q, self.num_var)
        if (q < 0).any() or (q > 1).any():
            raise ValueError("q must be within [0, 1]!")
        out = np.zeros_like(q, dtype=float)
        out[0] = self.lb
        for d in range(len(out)):
            prev = out[max(d - 1, 0)]
            out[d] = (1.0 - (1.0 - q[d]) ** (1.0 / (self.num_var - d))) * (self.ub - prev) + prev
        return out

    def elementwise_cdf(self, p):
        p = _check_unit_vector(p, self.num_var, "p")
        c = np.zeros(self.num_var)
        for d in range(self.num_var):
            prev = p[d - 1] if d > 0 else self.lb
            if p[d] <= prev:
                c[d] = 0.0
            elif p[d] >= self.ub:
                c[d] = 1.0
            else:
                c[d] = 1.0 - (1.0 - (p[d] - prev) / (self.ub - prev)) ** (self.num_var - d)
        return c

    def random_draw(self, size=None):
        single = size is None
        if single:
            size = 1
        shape = [self.num_var]
        try:
            shape.extend(size)
        except TypeError:
            shape.append(size)
        out = np.sort(np.random.uniform(low=self.lb, high=self.ub, size=shape), axis=0)
        return out.ravel() if single else out


def unique_rows(arr, return_index=False, return_inverse=False):
    """Unique rows of a 2-D array (utils.py:1666-1721).  Row order of the result follows numpy's
    lexicographic sort, like the reference's void-view trick."""
    arr = np.ascontiguousarray(arr)
    out = np.unique(arr, axis=0, return_index=return_index, return_inverse=return_inverse)
    if return_inverse:
        out = list(out)
        out[-1] = np.asarray(out[-1]).ravel()
        out = tuple(out)
    return out
