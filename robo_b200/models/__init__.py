from .gaussian_process import GaussianProcess  # noqa: F401
from .gaussian_process_mcmc import GaussianProcessMCMC  # noqa: F401
from .fabolas_gp import FabolasGP, FabolasGPMCMC  # noqa: F401
