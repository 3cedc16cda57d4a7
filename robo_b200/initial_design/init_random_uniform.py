"""Uniform initial design (robo/initial_design/init_random_uniform.py:5-30); host RNG so the
stream of random numbers is the reference's for the same seed."""
import numpy as np


def init_random_uniform(lower, upper, n_points, rng=None):
    if rng is None:
        rng = np.random.RandomState(np.random.randint(0, 10000))
    n_dims = lower.shape[0]
    return np.array([rng.uniform(lower, upper, n_dims) for _ in range(n_points)])
