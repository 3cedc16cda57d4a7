#!/usr/bin/env python
"""Incremental refit (gpk_fit_append) at the benchmark size: N0 -> N0 + k rows inside the last
128-row block, timed against what a BO iteration with frozen hyper-parameters costs otherwise (full factorisation +
L^-1 build), with the agreement of the two paths.  N=4096 D=16 by default (N0 = N - 8)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from robo_b200 import _lib                     # noqa: E402
from robo_b200 import kernels as K             # noqa: E402

N, D, KADD = int(os.environ.get("N", 4096)), 16, int(os.environ.get("KADD", 8))
rng = np.random.RandomState(1234)
X = rng.rand(N, D)
y = np.sinc(X * 10 - 5).sum(axis=1) + 0.01 * rng.randn(N)
Xs = np.random.RandomState(4321).rand(4096, D)
theta = np.concatenate(([0.0], np.full(D, np.log(D / 4.0))))
f = K.Product(K.ConstantKernel(theta[0], ndim=D), K.Matern52Kernel(np.exp(theta[1:]), ndim=D)).flatten()
da = float(np.sqrt(np.float64(np.sqrt(1e-3)) ** 2 + 1.25e-12) ** 2)
N0 = N - KADD


def fresh(n):
    h = _lib.Handle(0)
    h.set_data(X[:n], y[:n])
    h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
    return h


# full path: set_data + fit + first prediction (builds L^-1)
hf = fresh(N)
full_ms = []
for _ in range(5):
    t0 = time.perf_counter()
    hf.set_data(X, y)
    hf.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
    ld_f, ll_f = hf.fit(da, float(np.mean(y)))
    mu_f, var_f = hf.predict(Xs[:128])
    full_ms.append((time.perf_counter() - t0) * 1e3)
tf = hf.timings()

# incremental path: the same end state from a model fitted on N0 rows
app_ms, app_dev_ms = [], []
for _ in range(5):
    h = fresh(N0)
    h.fit(da, float(np.mean(y[:N0])))
    h.predict(Xs[:128])
    t0 = time.perf_counter()
    res = h.fit_append(X, y, da, float(np.mean(y)))
    mu_a, var_a = h.predict(Xs[:128])
    app_ms.append((time.perf_counter() - t0) * 1e3)
    app_dev_ms.append(h.timings()["fit_ms"])
    assert res is not None
    h.close()
ld_a, ll_a = res
out = {
    "n": N, "appended_rows": KADD, "d": D,
    "full_refit_wall_ms": float(np.median(full_ms)), "full_fit_device_ms": tf["fit_ms"], "linv_device_ms": tf["linv_ms"],
    "append_wall_ms": float(np.median(app_ms)), "append_device_ms": float(np.median(app_dev_ms)),
    "loglik_rel_diff": abs(ll_a - ll_f) / abs(ll_f), "logdet_rel_diff": abs(ld_a - ld_f) / abs(ld_f),
    "mean_max_scaled_diff": float(np.max(np.abs(mu_a - mu_f) / np.maximum(np.abs(mu_f), np.std(y)))),
    "var_max_rel_diff": float(np.max(np.abs(var_a - var_f) / np.maximum(var_f, 1e-6))),
}
print(json.dumps(out))
