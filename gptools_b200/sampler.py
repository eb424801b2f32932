"""Affine-invariant ensemble sampler (Goodman & Weare 2010 stretch move) driven by a BATCHED log-probability.

The reference delegates hyperparameter MCMC to emcee.EnsembleSampler (gaussian_process.py:1757-1787), which
calls the log-posterior once per walker (optionally from a pool of worker processes).  emcee's parallel
stretch move updates one half of the ensemble at a time using the other half as the complementary set;
all proposals of a half are independent, so here they are evaluated as ONE device launch
(GaussianProcess.update_hyperparameters_batch).  Attribute names follow emcee's sampler object so the
reference's post-processing idioms (``sampler.chain[:, burn::thin, :]``, ``flatchain``,
``acceptance_fraction``, continuing with ``sampler=``) keep working.
"""
import numpy as np

__all__ = ["EnsembleSampler"]


class EnsembleSampler(object):
    def __init__(self, nwalkers, dim, lnprob_batch, a=2.0, random_state=None):
        if nwalkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")
        if nwalkers < 2 * dim:
            raise ValueError("The number of walkers needs to be at least twice the dimension of the problem.")
        self.k = int(nwalkers)
        self.dim = int(dim)
        self.a = float(a)
        self.lnprob_batch = lnprob_batch
        self._random = random_state if random_state is not None else np.random.RandomState()
        self.reset()

    def reset(self):
        self._chain = np.empty((self.k, 0, self.dim))
        self._lnprob = np.empty((self.k, 0))
        self.naccepted = np.zeros(self.k)
        self.iterations = 0

    @property
    def chain(self):
        return self._chain

    @property
    def lnprobability(self):
        return self._lnprob

    @property
    def flatchain(self):
        s = self._chain.shape
        return self._chain.reshape(s[0] * s[1], s[2])

    @property
    def flatlnprobability(self):
        return self._lnprob.flatten()

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iterations, 1)

    def _lnprob_checked(self, p):
        lp = np.asarray(self.lnprob_batch(p), dtype=float)
        lp[np.isnan(lp)] = -np.inf
        return lp

    def run_mcmc(self, pos0, N, lnprob0=None):
        """Advance every walker N steps from ``pos0`` (nwalkers, dim).  Returns (pos, lnprob)."""
        p = np.array(pos0, dtype=float)
        if p.shape != (self.k, self.dim):
            raise ValueError("pos0 must have shape (nwalkers, dim)")
        lnprob = self._lnprob_checked(p) if lnprob0 is None else np.array(lnprob0, dtype=float)
        if not np.isfinite(lnprob).any():
            raise ValueError("The initial log-probability is -inf for every walker.")
        chain = np.empty((self.k, N, self.dim))
        lnp_hist = np.empty((self.k, N))
        half = self.k // 2
        first, second = slice(half), slice(half, self.k)
        for it in range(N):
            for S0, S1 in ((first, second), (second, first)):
                s = p[S0]
                c = p[S1]
                ns, nc = len(s), len(c)
                zz = ((self.a - 1.0) * self._random.rand(ns) + 1.0) ** 2.0 / self.a
                partner = self._random.randint(nc, size=ns)
                q = c[partner] - zz[:, None] * (c[partner] - s)
                newlnprob = self._lnprob_checked(q)
                with np.errstate(invalid="ignore"):
                    lnpdiff = (self.dim - 1.0) * np.log(zz) + newlnprob - lnprob[S0]
                accept = lnpdiff > np.log(self._random.rand(ns))
                accept &= np.isfinite(newlnprob)
                idx = np.arange(self.k)[S0][accept]
                p[idx] = q[accept]
                lnprob[idx] = newlnprob[accept]
                self.naccepted[idx] += 1
            chain[:, it, :] = p
            lnp_hist[:, it] = lnprob
            self.iterations += 1
        self._chain = np.concatenate((self._chain, chain), axis=1)
        self._lnprob = np.concatenate((self._lnprob, lnp_hist), axis=1)
        return p, lnprob
