"""Loads tests/golden/*.npz and rebuilds the kernel each case was generated with
(kernel constructors mirror oracle/make_golden.py)."""
import os

import numpy as np

from oracle import george_oracle as G
from oracle import robo_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GP_CASES = ["gp_unit", "gp_branin_ny0", "gp_branin_ny1", "gp_autobounds",
            "gp_rbf_d8", "gp_prod1d", "gp_mid_d16"]


def kernel_spec(name):
    """-> (family, theta) in the neutral form both the oracle and the product
    kernel constructors understand: family in {'matern52','rbf','prod1d_matern52'},
    theta = george parameter vector of the kernel."""
    if name == "gp_unit":
        return "matern52_noamp", np.zeros(2)
    if name.startswith("gp_branin"):
        return "matern52", np.array([np.log(1.7), np.log(0.15), np.log(0.4)])
    if name == "gp_autobounds":
        return "matern52", np.concatenate(([np.log(1.0 / 3)], np.log([0.3, 0.5, 0.7])))
    if name == "gp_rbf_d8":
        return "rbf", O.synthetic_problem(2, 8, 1)[3]
    if name == "gp_prod1d":
        return "prod1d_matern52", np.concatenate(([np.log(1.0 / 3)], np.log([0.2, 0.05, 0.6])))
    if name == "gp_mid_d16":
        return "matern52", O.synthetic_problem(2, 16, 1)[3]
    raise KeyError(name)


def oracle_kernel(family, theta, D):
    if family == "matern52_noamp":
        return G.Matern52Kernel(np.exp(theta), ndim=D)
    if family in ("matern52", "rbf"):
        return O.make_kernel(family, D, theta)
    if family == "prod1d_matern52":
        k = G.ConstantKernel(theta[0], ndim=D)
        for d in range(D):
            k = G.Product(k, G.Matern52Kernel(np.exp(theta[1 + d:2 + d]), ndim=D, axes=d))
        return k
    raise KeyError(family)


def load_case(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    d["lower_"] = d["lower"] if d["lower"].size else None
    d["upper_"] = d["upper"] if d["upper"].size else None
    family, theta = kernel_spec(name)
    D = d["X"].shape[1]
    return d, (lambda: oracle_kernel(family, theta, D))
