#!/usr/bin/env python
"""Fit-time sweep over the Cholesky implementation options (diag kernel x look-ahead) at N=4096, D=16,
checking that every variant produces the same log-likelihood."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from robo_b200 import _lib                     # noqa: E402
from robo_b200 import kernels as K             # noqa: E402

N, D = int(os.environ.get("N", 4096)), 16
rng = np.random.RandomState(1234)
X = rng.rand(N, D)
y = np.sinc(X * 10 - 5).sum(axis=1) + 0.01 * rng.randn(N)
theta = np.concatenate(([0.0], np.full(D, np.log(D / 4.0))))
f = K.Product(K.ConstantKernel(theta[0], ndim=D), K.Matern52Kernel(np.exp(theta[1:]), ndim=D)).flatten()
da = float(np.sqrt(np.float64(np.sqrt(1e-3)) ** 2 + 1.25e-12) ** 2)
ref = None
for diag in (0, 2, 3):
    for la in (0, 1):
        h = _lib.Handle(0)
        h.set_option("diag", diag)
        h.set_option("lookahead", la)
        h.set_data(X, y)
        h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
        ts = []
        for _ in range(6):
            logdet, ll = h.fit(da, float(np.mean(y)))
            ts.append(h.timings()["fit_ms"])
        ref = ll if ref is None else ref
        print("diag=%d lookahead=%d  fit_ms median %.3f min %.3f  ll=%.12f  rel.diff vs first %.2e"
              % (diag, la, np.median(ts[1:]), min(ts), ll, abs(ll - ref) / abs(ref)))
        h.close()
