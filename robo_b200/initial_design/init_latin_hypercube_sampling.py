"""Latin hypercube initial design (robo/initial_design/init_latin_hypercube_sampling.py:5-44)."""
import numpy as np


def init_latin_hypercube_sampling(lower, upper, n_points, rng=None):
    if rng is None:
        rng = np.random.RandomState(np.random.randint(0, 10000))
    n_dims = lower.shape[0]
    s_bounds = np.array([np.linspace(lower[i], upper[i], n_points + 1) for i in range(n_dims)])
    s_lower, s_upper = s_bounds[:, :-1], s_bounds[:, 1:]
    samples = s_lower + rng.uniform(0, 1, s_lower.shape) * (s_upper - s_lower)
    for i in range(n_dims):
        rng.shuffle(samples[i, :])
    return samples.T
