"""robo_b200 — B200-native (sm_100a) GP posterior + acquisition hot path behind RoBO's API.

Public surface mirrors the reference's modules for this path only:
    robo_b200.models.gaussian_process.GaussianProcess      (robo/models/gaussian_process.py)
    robo_b200.acquisition_functions.{EI, LogEI, PI, LCB}    (robo/acquisition_functions/*.py)
    robo_b200.maximizers.random_sampling.RandomSampling     (robo/maximizers/random_sampling.py)
    robo_b200.kernels                                       (george.kernels duck-type)
All arithmetic runs in robo_b200/libgpk.so (CUDA, C ABI in include/gpk.h).
"""
__version__ = "0.1.0"
