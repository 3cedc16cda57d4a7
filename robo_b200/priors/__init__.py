"""Hyper-priors used by the fmin facade: O(H) scalar host math added to the GPU log-likelihood
(out of the hot path, stays Python).  Semantics — including the quirks —
follow robo/priors/base_prior.py and robo/priors/default_priors.py."""
import numpy as np
import scipy.stats as sps


class TophatPrior(object):
    """base_prior.py:75-156."""

    def __init__(self, l_bound, u_bound, rng=None):
        self.rng = np.random.RandomState(np.random.randint(0, 10000)) if rng is None else rng
        self.min, self.max = l_bound, u_bound
        if not (self.max > self.min):
            raise Exception("Upper bound of Tophat prior must be greater than the lower bound!")

    def lnprob(self, theta):
        if np.any(theta < self.min) or np.any(theta > self.max):
            return -np.inf
        return 0

    def sample_from_prior(self, n_samples):
        p0 = self.min + self.rng.rand(n_samples) * (self.max - self.min)
        return p0[:, np.newaxis]

    def gradient(self, theta):
        return np.zeros([theta.shape[0]])


class HorseshoePrior(object):
    """base_prior.py:158-237 (returns +inf at theta == 0 like the reference, :194-196)."""

    def __init__(self, scale=0.1, rng=None):
        self.rng = np.random.RandomState(np.random.randint(0, 10000)) if rng is None else rng
        self.scale = scale

    def lnprob(self, theta):
        if np.any(theta == 0.0):
            return np.inf
        return np.log(np.log(1 + 3.0 * (self.scale / np.exp(theta)) ** 2))

    def sample_from_prior(self, n_samples):
        lamda = np.abs(self.rng.standard_cauchy(size=n_samples))
        p0 = np.log(np.abs(self.rng.randn() * lamda * self.scale))
        return p0[:, np.newaxis]

    def gradient(self, theta):
        a = -(6 * self.scale ** 2)
        b = (3 * self.scale ** 2 + np.exp(2 * theta))
        b *= np.log(3 * self.scale ** 2 * np.exp(- 2 * theta) + 1)
        return a / b


class LognormalPrior(object):
    """base_prior.py:239-316 (``mean`` is passed as scipy's ``loc``, :278, like the reference)."""

    def __init__(self, sigma, mean=0, rng=None):
        self.rng = np.random.RandomState(np.random.randint(0, 10000)) if rng is None else rng
        self.sigma, self.mean = sigma, mean

    def lnprob(self, theta):
        return sps.lognorm.logpdf(theta, self.sigma, loc=self.mean)

    def sample_from_prior(self, n_samples):
        p0 = self.rng.lognormal(mean=self.mean, sigma=self.sigma, size=n_samples)
        return p0[:, np.newaxis]

    def gradient(self, theta):
        return None


class DefaultPrior(object):
    """default_priors.py:8-53: lognormal on the amplitude, tophat(-10, 2) on the length scales,
    horseshoe(0.1) on the noise; gradient identically zero (:51-53)."""

    def __init__(self, n_dims, rng=None):
        self.rng = np.random.RandomState(np.random.randint(0, 10000)) if rng is None else rng
        self.n_dims = n_dims
        self.tophat = TophatPrior(-10, 2, rng=self.rng)
        self.ln_prior = LognormalPrior(mean=0.0, sigma=1.0, rng=self.rng)
        self.horseshoe = HorseshoePrior(scale=0.1, rng=self.rng)

    def lnprob(self, theta):
        lp = 0
        lp += self.ln_prior.lnprob(theta[0])
        lp += self.tophat.lnprob(theta[1:-1])
        lp += self.horseshoe.lnprob(theta[-1])
        return lp

    def sample_from_prior(self, n_samples):
        p0 = np.zeros([n_samples, self.n_dims])
        p0[:, 0] = self.ln_prior.sample_from_prior(n_samples)[:, 0]
        ls_sample = np.array([self.tophat.sample_from_prior(n_samples)[:, 0]
                              for _ in range(1, (self.n_dims - 1))]).T
        p0[:, 1:(self.n_dims - 1)] = ls_sample
        p0[:, -1] = self.horseshoe.sample_from_prior(n_samples)[:, 0]
        return p0

    def gradient(self, theta):
        return np.zeros([theta.shape[0]])
