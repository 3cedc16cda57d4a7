#!/usr/bin/env python
"""Workload for the ncu captures of round 2 (tools/profile_r2.sh): C2-sized fit, two scoring chunks, likelihood gradient.
    python tools/profile_driver.py [N] [D] [M]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from robo_b200 import _lib                                   # noqa: E402
from robo_b200 import kernels as K                           # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
rng = np.random.RandomState(1234)
X = rng.rand(N, D)
y = np.sinc(X * 10 - 5).sum(axis=1) + 0.01 * rng.randn(N)
Xs = np.random.RandomState(4321).rand(M, D)
theta = np.concatenate(([0.0], np.full(D, np.log(D / 4.0))))
h = _lib.Handle(0)
for k, v in (("GPK_OPT_CHAINSPLIT", "chainsplit"), ("GPK_OPT_COV", "cov")):
    if os.environ.get(k):
        h.set_option(v, int(os.environ[k]))
h.set_data(X, y)
f = K.Product(K.ConstantKernel(theta[0], ndim=D), K.Matern52Kernel(np.exp(theta[1:]), ndim=D)).flatten()
h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
dadd = float(np.sqrt(np.float64(np.sqrt(1e-3)) ** 2 + 1.25e-12) ** 2)
for _ in range(3):
    logdet, ll = h.fit(dadd, float(np.mean(y)))
t = h.timings()
print("fit_ms %.3f kbuild %.3f potrf %.3f" % (t["fit_ms"], t["kbuild_ms"], t["potrf_ms"]))
for _ in range(2):
    r = h.acq(Xs, _lib.ACQ_EI, float(np.min(y)), 0.0, want_values=False)
t = h.timings()
print("score_ms %.3f kstar %.3f vargemm %.3f linv %.3f best %d" % (t["score_ms"], t["kstar_ms"], t["vargemm_ms"], t["linv_ms"], r["best_idx"]))
g = h.nll_grad(1e-3, D)
print("grad[0:3]", g[:3])
