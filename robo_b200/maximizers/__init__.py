from .random_sampling import RandomSampling  # noqa: F401
