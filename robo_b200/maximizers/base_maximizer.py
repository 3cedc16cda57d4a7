"""Maximizer interface (constructor arguments and attributes of robo/maximizers/base_maximizer.py:7-31:
``objective_func``, ``lower``, ``upper``, ``rng``; subclasses implement ``maximize``)."""
import numpy as np


class BaseMaximizer(object):
    """Holds the acquisition function to maximise and the box it is maximised over."""

    def __init__(self, objective_function, lower, upper, rng=None):
        self.objective_func = objective_function
        self.lower, self.upper = lower, upper
        self.rng = rng if rng is not None else np.random.RandomState(np.random.randint(0, 10000))

    def maximize(self):
        raise NotImplementedError("subclasses return the point with the highest acquisition value")
