"""Host-side input/output scaling (robo/util/normalization.py:4-32).  Only the *training*
data passes through these (O(N D) once per train); test inputs and predictive moments are
scaled inside the CUDA kernels (gpk_set_input_bounds / gpk_set_output_transform)."""
import numpy as np


def zero_one_normalization(X, lower=None, upper=None):
    if lower is None:
        lower = np.min(X, axis=0)
    if upper is None:
        upper = np.max(X, axis=0)
    return np.true_divide((X - lower), (upper - lower)), lower, upper


def zero_one_unnormalization(X_normalized, lower, upper):
    return lower + (upper - lower) * X_normalized


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    if mean is None:
        mean = np.mean(X, axis=0)
    if std is None:
        std = np.std(X, axis=0)
    return (X - mean) / std, mean, std


def zero_mean_unit_var_unnormalization(X_normalized, mean, std):
    return X_normalized * std + mean
