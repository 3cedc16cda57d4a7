from .gaussian_process import GaussianProcess  # noqa: F401
