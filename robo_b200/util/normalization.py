"""Training-data scaling on the host (the four helpers of robo/util/normalization.py:4-32, same names and
return conventions because RoBO's callers import them by name).

Only the *training* set passes through here, once per ``train`` (O(N D)).  Test inputs are scaled and
predictive moments un-scaled inside the CUDA kernels (``gpk_set_input_bounds`` /
``gpk_set_output_transform``), with the same arithmetic: a true division by ``upper - lower`` and
``x * std + mean``.
"""
import numpy as np


def _column_range(X, lower, upper):
    lo = X.min(axis=0) if lower is None else lower
    hi = X.max(axis=0) if upper is None else upper
    return lo, hi


def zero_one_normalization(X, lower=None, upper=None):
    """(X - lower) / (upper - lower); missing bounds default to the column extrema of X and are returned so
    that test points can be mapped with the same box."""
    lo, hi = _column_range(X, lower, upper)
    return np.true_divide(X - lo, hi - lo), lo, hi


def zero_one_unnormalization(X_normalized, lower, upper):
    span = upper - lower
    return lower + span * X_normalized


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    """Standardise with the population standard deviation (ddof = 0)."""
    mu = X.mean(axis=0) if mean is None else mean
    sd = X.std(axis=0) if std is None else std
    return (X - mu) / sd, mu, sd


def zero_mean_unit_var_unnormalization(X_normalized, mean, std):
    return X_normalized * std + mean
