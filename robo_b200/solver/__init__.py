from .bayesian_optimization import BayesianOptimization  # noqa: F401
