from .random_sampling import RandomSampling  # noqa: F401
from .device_random_sampling import DeviceRandomSampling  # noqa: F401
