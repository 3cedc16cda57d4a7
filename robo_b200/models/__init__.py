from .gaussian_process import GaussianProcess  # noqa: F401
from .gaussian_process_mcmc import GaussianProcessMCMC  # noqa: F401
