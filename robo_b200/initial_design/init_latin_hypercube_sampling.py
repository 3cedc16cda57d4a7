"""Latin-hypercube initial design with the sampling scheme and random-number consumption of
robo/initial_design/init_latin_hypercube_sampling.py:5-44 (one uniform draw per stratum and dimension, then
an independent shuffle of every dimension), so a seeded ``rng`` yields the reference's design."""
import numpy as np


def init_latin_hypercube_sampling(lower, upper, n_points, rng=None):
    rng = np.random.RandomState(np.random.randint(0, 10000)) if rng is None else rng
    n_dims = lower.shape[0]
    edges = np.stack([np.linspace(lower[d], upper[d], n_points + 1) for d in range(n_dims)])   # (D, n+1)
    left, width = edges[:, :-1], np.diff(edges, axis=1)
    design = left + rng.uniform(0, 1, left.shape) * (edges[:, 1:] - left)                       # one point per stratum
    del width
    for d in range(n_dims):
        rng.shuffle(design[d, :])
    return design.T
