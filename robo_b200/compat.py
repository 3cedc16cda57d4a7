"""Run the UNMODIFIED reference on the B200 path (INTEGRATION.md option C).

RoBO's own ``robo.models.gaussian_process`` does ``import george``; george is a third-party
dependency that is not installable in this environment.  ``install()`` registers

* ``george``          -> a module whose ``GP`` is :class:`GeorgeGP` (george.GP verbs on a gpk handle) and
                         whose ``kernels`` is :mod:`robo_b200.kernels`
* ``emcee``           -> :mod:`robo_b200.util.ensemble_sampler` (``EnsembleSampler``), only if emcee is missing
* ``pybnn.*``, ``pyrfr.regression`` -> inert stubs, only if missing (``robo/fmin/__init__.py`` imports every
                         facade, and ``robo/models/random_forest.py:7-11`` raises at import without pyrfr)

after which ``robo.fmin.bayesian_optimization(..., model_type="gp" | "gp_mcmc")``, the reference solver,
every maximizer and MarginalizationGPMCMC run as they are, with all GP arithmetic in libgpk.so.
"""
import sys
import types

import numpy as np

from robo_b200 import kernels as _kernels
from robo_b200.device_gp import DeviceGP


class GeorgeGP(DeviceGP):
    """``george.GP(kernel, mean=m)`` duck type.  george takes the targets only in ``log_likelihood(y)`` /
    ``predict(y, t)``; the device factorisation wants them up front (the forward solve is fused into it), so
    ``compute`` factorises with the targets of the previous call when there are any (zeros otherwise — the
    factor itself, and therefore the LinAlgError behaviour, does not depend on y) and ``log_likelihood`` /
    ``predict`` refit only if they are handed different targets."""

    def __init__(self, kernel, fit_kernel=True, mean=None, fit_mean=None, white_noise=None,
                 fit_white_noise=None, solver=None, **kwargs):
        super(GeorgeGP, self).__init__(kernel, mean=0.0 if mean is None else float(mean),
                                       device=kwargs.get("device", 0), white_noise=white_noise)

    def compute(self, x, yerr=0.0, **kwargs):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None]
        if self._y is None or len(self._y) != len(x):
            self.set_data(x, np.zeros(len(x)))
            self._y_is_dummy = True
        elif self._x is None or x.shape != self._x.shape or not np.array_equal(x, self._x):
            self.set_data(x, self._y)
        super(GeorgeGP, self).compute(None, yerr=yerr)

    def _with_targets(self, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        if getattr(self, "_y_is_dummy", False) or not np.array_equal(y, self._y):
            self.set_data(self._x, y)
            self._y_is_dummy = False
            super(GeorgeGP, self).compute(None, yerr=self._yerr)

    def log_likelihood(self, y, quiet=False):
        try:
            self._with_targets(y)
        except np.linalg.LinAlgError:
            if quiet:
                return -np.inf
            raise
        return super(GeorgeGP, self).log_likelihood(None, quiet=quiet)

    lnlikelihood = log_likelihood

    def predict(self, y, t, return_cov=True, return_var=False):
        self._with_targets(y)
        t = np.ascontiguousarray(t, dtype=np.float64)
        if t.ndim == 1:
            t = t[:, None]
        if return_var or not return_cov:
            mu, var = self.predict_moments(t)
            return (mu, var) if return_var else mu
        return self.posterior_cov(t)          # george returns the raw covariance; the reference clips it itself

    def sample_conditional(self, y, t, size=1):
        mu, cov = self.predict(y, t)
        return np.random.multivariate_normal(mu, cov, size=size) if size > 1 \
            else np.random.multivariate_normal(mu, cov)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install(force_emcee=False):
    """Register the shims in sys.modules (idempotent).  Returns the list of module names it provided."""
    provided = []
    kmod = _module("george.kernels", **{k: getattr(_kernels, k) for k in _kernels.__all__})
    gmod = _module("george", GP=GeorgeGP, kernels=kmod, __version__="0.3-robo_b200-shim")
    sys.modules["george"] = gmod
    sys.modules["george.kernels"] = kmod
    provided += ["george", "george.kernels"]

    def missing(name):
        if name in sys.modules:
            return False
        try:
            __import__(name)
            return False
        except Exception:
            return True

    if force_emcee or missing("emcee"):
        from robo_b200.util import ensemble_sampler
        sys.modules["emcee"] = _module("emcee", EnsembleSampler=ensemble_sampler.EnsembleSampler,
                                       __version__="2-robo_b200-shim")
        provided.append("emcee")

    class _Unavailable(object):
        def __init__(self, *a, **k):
            raise ImportError("this optional RoBO dependency is not installed (robo_b200.compat stub)")

    if missing("pybnn"):
        sys.modules["pybnn"] = _module("pybnn")
        sys.modules["pybnn.dngo"] = _module("pybnn.dngo", DNGO=_Unavailable)
        sys.modules["pybnn.bohamiann"] = _module("pybnn.bohamiann", Bohamiann=_Unavailable)
        sys.modules["pybnn.multi_task_bohamiann"] = _module("pybnn.multi_task_bohamiann", MultiTaskBohamiann=_Unavailable)
        sys.modules["pybnn.util"] = _module("pybnn.util")
        sys.modules["pybnn.util.layers"] = _module("pybnn.util.layers", AppendLayer=_Unavailable)
        provided.append("pybnn")
    if missing("pyrfr"):
        sys.modules["pyrfr"] = _module("pyrfr")
        sys.modules["pyrfr.regression"] = _module("pyrfr.regression")
        provided.append("pyrfr")
    if not hasattr(np, "Infinity"):          # log_ei.py:89,96,118 use np.Infinity (removed in numpy 2)
        np.Infinity = np.inf
    return provided
