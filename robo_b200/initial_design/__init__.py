from .init_random_uniform import init_random_uniform  # noqa: F401
from .init_latin_hypercube_sampling import init_latin_hypercube_sampling  # noqa: F401
