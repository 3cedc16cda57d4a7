"""RandomSampling maximizer (robo/maximizers/random_sampling.py:7-52): the natural batched
caller of the hot path.  Candidate generation is the reference's (70 % uniform, 30 % Gaussian
around the incumbent, concatenated uniform-first); the scoring is ONE fused GPU call that
returns the arg-max index, so the acquisition values never cross PCIe."""
import numpy as np

from robo_b200.initial_design import init_random_uniform
from robo_b200.maximizers.base_maximizer import BaseMaximizer


class RandomSampling(BaseMaximizer):

    def __init__(self, objective_function, lower, upper, n_samples=500, rng=None):
        self.n_samples = n_samples
        super(RandomSampling, self).__init__(objective_function, lower, upper, rng)

    def candidates(self):
        rand = init_random_uniform(self.lower, self.upper, int(self.n_samples * .7))
        loc = self.objective_func.model.get_incumbent()[0],
        scale = np.ones([self.lower.shape[0]]) * 0.1
        n_inc = int(self.n_samples * 0.3)
        if n_inc > 0:
            rand_incs = np.array([np.clip(np.random.normal(loc, scale), self.lower, self.upper)[0]
                                  for _ in range(n_inc)])
            return np.concatenate((rand, rand_incs), axis=0)
        return rand

    def maximize(self):
        X = self.candidates()
        if hasattr(self.objective_func, "argmax"):
            return X[self.objective_func.argmax(X)]
        y = self.objective_func(X)
        return X[y.argmax()]
