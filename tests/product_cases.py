"""Builds robo_b200 kernels / models for the golden cases (mirrors tests/golden_cases.py,
which builds the oracle's)."""
import numpy as np

from robo_b200 import kernels as K
from robo_b200.models.gaussian_process import GaussianProcess


def product_kernel(family, theta, D):
    theta = np.asarray(theta, dtype=np.float64)
    if family == "matern52_noamp":
        return K.Matern52Kernel(np.exp(theta), ndim=D)
    if family == "matern52":
        return K.Product(K.ConstantKernel(theta[0], ndim=D), K.Matern52Kernel(np.exp(theta[1:]), ndim=D))
    if family == "rbf":
        return K.Product(K.ConstantKernel(theta[0], ndim=D), K.ExpSquaredKernel(np.exp(theta[1:]), ndim=D))
    if family == "prod1d_matern52":
        k = K.ConstantKernel(theta[0], ndim=D)
        for d in range(D):
            k = K.Product(k, K.Matern52Kernel(np.exp(theta[1 + d:2 + d]), ndim=D, axes=d))
        return k
    raise KeyError(family)


def product_model(d, family, theta, prior=None):
    """Un-trained robo_b200 GaussianProcess configured like golden case dict ``d``."""
    D = d["X"].shape[1]
    return GaussianProcess(product_kernel(family, theta, D), prior=prior, noise=float(d["noise"]),
                           normalize_input=bool(d["normalize_input"]),
                           normalize_output=bool(d["normalize_output"]),
                           lower=d["lower_"], upper=d["upper_"], rng=np.random.RandomState(0))


# tolerances (north_star): 1e-10 relative on the posterior mean / variance, 1e-8 on EI.
# Denominators (SURVEY.md section 8c asks to state them):
#   mean     : max(|mu|, std(y))            (mu crosses zero; its error scales with |alpha|)
#   variance : max(var, 1e-6 * k(x,x))      (var = k** - ||L^-1 k*||^2 cancels near data)
def assert_mean_close(mu, ref, y, tol=1e-10):
    scale = np.maximum(np.abs(ref), np.std(y))
    err = np.max(np.abs(mu - ref) / scale)
    assert err <= tol, "posterior mean: scaled error %.3g > %.1g" % (err, tol)


def assert_var_close(var, ref, kss, tol=1e-10):
    scale = np.maximum(np.abs(ref), 1e-6 * kss)
    err = np.max(np.abs(var - ref) / scale)
    assert err <= tol, "posterior variance: scaled error %.3g > %.1g" % (err, tol)


def assert_acq_close(a, ref, rtol=1e-8, atol=1e-13):
    a, ref = np.asarray(a), np.asarray(ref)
    assert a.shape == ref.shape
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(a), fin), "finite pattern differs"
    assert np.array_equal(a[~fin], ref[~fin], equal_nan=True) or np.all(a[~fin] == ref[~fin])
    err = np.abs(a[fin] - ref[fin]) - (atol + rtol * np.abs(ref[fin]))
    assert np.all(err <= 0), "acquisition: max excess error %.3g" % err.max()
