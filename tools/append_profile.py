#!/usr/bin/env python
"""One gpk_fit_append between cudaProfilerStart/Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from robo_b200 import _lib                     # noqa: E402
from robo_b200 import kernels as K             # noqa: E402

N, D = 4096, 16
rng = np.random.RandomState(1234)
X = rng.rand(N, D)
y = np.sinc(X * 10 - 5).sum(axis=1) + 0.01 * rng.randn(N)
theta = np.concatenate(([0.0], np.full(D, np.log(D / 4.0))))
f = K.Product(K.ConstantKernel(theta[0], ndim=D), K.Matern52Kernel(np.exp(theta[1:]), ndim=D)).flatten()
da = float(np.sqrt(np.float64(np.sqrt(1e-3)) ** 2 + 1.25e-12) ** 2)
for rep in range(2):
    h = _lib.Handle(0)
    h.set_data(X[:N - 8], y[:N - 8])
    h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
    h.fit(da, float(np.mean(y[:N - 8])))
    h.predict(X[:128])
    if rep == 1:
        torch.cuda.cudart().cudaProfilerStart()
    print(h.fit_append(X, y, da, float(np.mean(y))), h.timings()["fit_ms"])
    if rep == 1:
        torch.cuda.cudart().cudaProfilerStop()
    h.close()
