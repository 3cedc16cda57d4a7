from .init_random_uniform import init_random_uniform  # noqa: F401
