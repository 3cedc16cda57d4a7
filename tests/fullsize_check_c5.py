#!/usr/bin/env python
"""TEST INFRASTRUCTURE (lives under tests/ because it calls the oracle; not collected by pytest: run by hand on a GPU box).
One-off full-size check of config 5 (N=8192, D=32): GPU log-likelihood against the CPU oracle (about a
minute of numpy/LAPACK), GPU gradient against central differences of the GPU nll."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))     # repo root (tests/..)
sys.path.insert(0, ROOT)
from robo_b200 import _lib
from robo_b200 import kernels as K
from oracle import robo_oracle as O

N, D = 8192, 32
X, y, _, theta, noise = O.synthetic_problem(N, D, 1)
theta = theta + 0.05 * np.random.RandomState(5).randn(D + 1)
h = _lib.Handle(0)
h.set_data(X, y)
mean = float(np.mean(y))


def gpu_nll(th, nz):
    f = K.Product(K.ConstantKernel(th[0], ndim=D), K.Matern52Kernel(np.exp(th[1:]), ndim=D)).flatten()
    h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
    return -h.fit(float(np.sqrt(np.float64(np.sqrt(nz)) ** 2 + 1.25e-12) ** 2), mean)[1]


nll = gpu_nll(theta, noise)
g = h.nll_grad(noise, D)
t0 = time.time()
st = O.gp_fit(O.make_kernel("matern52", D, theta), X, y, noise=noise, normalize_input=False)
ll_ref, logdet_ref = O.gp_loglik_terms(st)
print("oracle fit %.1f s  ll_ref %.10f  gpu %.10f  rel.err %.2e" % (time.time() - t0, ll_ref, -nll, abs(-nll - ll_ref) / abs(ll_ref)))
for p in (0, 1, 17, D):
    hh = 1e-5
    tp, tm = theta.copy(), theta.copy()
    tp[p] += hh; tm[p] -= hh
    fd = (gpu_nll(tp, noise) - gpu_nll(tm, noise)) / (2 * hh)
    print("grad[%d] analytic %.8f  central diff %.8f  rel %.2e" % (p, g[p], fd, abs(g[p] - fd) / max(1, abs(fd))))
lp, lm = np.log(noise) + 1e-5, np.log(noise) - 1e-5
fd = (gpu_nll(theta, np.exp(lp)) - gpu_nll(theta, np.exp(lm))) / 2e-5
print("grad[noise] analytic %.8f  central diff %.8f  rel %.2e" % (g[-1], fd, abs(g[-1] - fd) / max(1, abs(fd))))
