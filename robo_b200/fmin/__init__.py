from .bayesian_optimization import bayesian_optimization  # noqa: F401
