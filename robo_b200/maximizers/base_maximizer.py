"""robo/maximizers/base_maximizer.py:7-31."""
import numpy as np


class BaseMaximizer(object):

    def __init__(self, objective_function, lower, upper, rng=None):
        self.lower = lower
        self.upper = upper
        self.objective_func = objective_function
        if rng is None:
            self.rng = np.random.RandomState(np.random.randint(0, 10000))
        else:
            self.rng = rng

    def maximize(self):
        raise NotImplementedError
