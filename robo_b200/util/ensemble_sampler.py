"""Affine-invariant ensemble sampler (Goodman & Weare 2010 stretch move), the algorithm behind
``emcee.EnsembleSampler``, which the reference drives at robo/models/gaussian_process_mcmc.py:114-142.

emcee is a third-party dependency of the reference that is not vendored and not installable
here, so the subset of its 2.x API that RoBO uses is provided:

    sampler = EnsembleSampler(nwalkers, dim, lnpostfn[, batch_lnpostfn=...])
    pos, lnprob, state = sampler.run_mcmc(p0, N, rstate0=rng)
    sampler.chain                       # (nwalkers, steps, dim)

As in emcee, one step updates the two halves of the ensemble in turn; each half needs the
log-posterior of nwalkers/2 proposals, which are INDEPENDENT — ``batch_lnpostfn(thetas)`` lets the
caller evaluate them together (robo_b200 runs one GPU factorisation per proposal on concurrent
streams).  The random stream is numpy's, not emcee's: MCMC parity with the reference is
statistical, not bitwise.
"""
import numpy as np


class EnsembleSampler(object):

    def __init__(self, nwalkers, dim, lnpostfn, a=2.0, batch_lnpostfn=None):
        if nwalkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")
        if nwalkers < 2 * dim:
            raise ValueError("The number of walkers needs to be at least twice the dimension of the problem.")
        self.k = int(nwalkers)
        self.dim = int(dim)
        self.a = float(a)
        self.lnpostfn = lnpostfn
        self.batch_lnpostfn = batch_lnpostfn
        self.random_state = None
        self._chain = np.empty((self.k, 0, self.dim))
        self._lnprob = np.empty((self.k, 0))
        self.naccepted = np.zeros(self.k)
        self.n_lnprob_calls = 0

    @property
    def chain(self):
        return self._chain

    @property
    def lnprobability(self):
        return self._lnprob

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(1, self._chain.shape[1])

    def _lnprob_many(self, thetas):
        self.n_lnprob_calls += len(thetas)
        if self.batch_lnpostfn is not None:
            out = np.asarray(self.batch_lnpostfn(thetas), dtype=np.float64)
        else:
            out = np.array([self.lnpostfn(t) for t in thetas], dtype=np.float64)
        out[np.isnan(out)] = -np.inf
        return out

    def run_mcmc(self, p0, N, rstate0=None, lnprob0=None):
        rng = rstate0 if isinstance(rstate0, np.random.RandomState) else np.random.RandomState()
        if rstate0 is not None and not isinstance(rstate0, np.random.RandomState):
            rng.set_state(rstate0)
        p = np.array(p0, dtype=np.float64, copy=True)
        if p.shape != (self.k, self.dim):
            raise ValueError("p0 must have shape (nwalkers, dim)")
        lnprob = self._lnprob_many(p) if lnprob0 is None else np.array(lnprob0, dtype=np.float64)
        if np.any(np.isnan(lnprob)):
            raise ValueError("The initial lnprob was NaN.")
        chain = np.empty((self.k, N, self.dim))
        lnp_hist = np.empty((self.k, N))
        half = self.k // 2
        first, second = slice(half), slice(half, self.k)
        for step in range(N):
            for S0, S1 in ((first, second), (second, first)):
                s, c = p[S0], p[S1]
                Ns, Nc = len(s), len(c)
                # stretch move: z ~ g(z) ∝ 1/sqrt(z) on [1/a, a]
                zz = ((self.a - 1.0) * rng.rand(Ns) + 1.0) ** 2.0 / self.a
                rint = rng.randint(Nc, size=(Ns,))
                q = c[rint] - zz[:, None] * (c[rint] - s)
                newlnprob = self._lnprob_many(q)
                lnpdiff = (self.dim - 1.0) * np.log(zz) + newlnprob - lnprob[S0]
                accept = lnpdiff > np.log(rng.rand(Ns))
                idx = np.arange(self.k)[S0][accept]
                p[idx] = q[accept]
                lnprob[idx] = newlnprob[accept]
                self.naccepted[idx] += 1
            chain[:, step] = p
            lnp_hist[:, step] = lnprob
        self._chain = np.concatenate((self._chain, chain), axis=1)
        self._lnprob = np.concatenate((self._lnprob, lnp_hist), axis=1)
        self.random_state = rng.get_state()
        return p, lnprob, self.random_state
