"""The restatement oracle/robo_oracle.py must reproduce the committed golden
vectors, which were produced by the reference's own classes (oracle/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import george_oracle as G
from oracle import robo_oracle as O
from tests.golden_cases import load_case, GP_CASES


@pytest.mark.parametrize("name", GP_CASES)
def test_gp_case(name, golden_dir):
    d, kernel_fn = load_case(name)
    st = O.gp_fit(kernel_fn(), d["X"], d["y"], noise=float(d["noise"]),
                  normalize_input=bool(d["normalize_input"]),
                  normalize_output=bool(d["normalize_output"]),
                  lower=d["lower_"], upper=d["upper_"])
    mu, var = O.gp_predict(st, d["Xs"])
    np.testing.assert_allclose(mu, d["mu"], rtol=1e-13)
    np.testing.assert_allclose(var, d["var"], rtol=1e-13)
    m = int(d["full_cov_m"])
    _, cov = O.gp_predict(st, d["Xs"][:m], full_cov=True)
    np.testing.assert_allclose(cov, d["cov"], rtol=1e-13)
    ll, logdet = O.gp_loglik_terms(st)
    np.testing.assert_allclose([ll, logdet], [d["ll"], d["logdet"]], rtol=1e-13)
    for kind in ("ei", "log_ei", "pi", "lcb"):
        np.testing.assert_allclose(O.acquisition(st, d["Xs"], kind), d["acq_" + kind], rtol=1e-13)
    inc_x, inc_y = O.gp_get_incumbent(st)
    np.testing.assert_allclose(inc_x, d["inc_x"], rtol=1e-15)
    assert inc_y == d["inc_y"]


def test_acq_moments(golden_dir):
    d = np.load(os.path.join(golden_dir, "acq_moments.npz"))
    m, v, eta = d["m"], d["v"], float(d["eta"])
    with np.errstate(all="ignore"):
        for par in (0.0, 0.3):
            np.testing.assert_allclose(O.acq_log_ei(m, v, eta, par), d["log_ei_par%g" % par], rtol=1e-14)
            np.testing.assert_allclose(O.acq_pi(m, v, eta, par), d["pi_par%g" % par], rtol=1e-14, equal_nan=True)
            np.testing.assert_allclose(O.acq_lcb(m, v, 1 + par), d["lcb_par%g" % (1 + par)], rtol=1e-14)
            pos = v > 0
            np.testing.assert_allclose(O.acq_ei(m[pos], v[pos], eta, par), d["ei_pos_par%g" % par], rtol=1e-14)
        assert O.acq_ei(m, v, eta).shape == (1, 1)        # ei.py:72-74 whole-batch zero


def test_grad_nll_correct_vs_finite_differences():
    X, y, _, theta, noise = O.synthetic_problem(40, 3, 1)
    st = O.gp_fit(O.make_kernel("matern52", 3, theta), X, y, noise=noise, normalize_input=False)
    th = np.append(theta, np.log(noise)) + 0.1
    g = O.gp_grad_nll_correct(st, th)
    h = 1e-6
    for p in range(len(th)):
        tp, tm = th.copy(), th.copy()
        tp[p] += h
        tm[p] -= h
        fd = (O.gp_nll(st, tp) - O.gp_nll(st, tm)) / (2 * h)
        assert abs(g[p] - fd) < 1e-5 * max(1.0, abs(fd))


def test_var_only_matches_full_cov_path():
    X, y, Xs, theta, noise = O.synthetic_problem(128, 8, 64)
    st = O.gp_fit(O.make_kernel("matern52", 8, theta), X, y, noise=noise, normalize_input=False)
    m1, v1 = O.gp_predict(st, Xs)
    m2, v2 = O.gp_predict_var_only(st, Xs)
    np.testing.assert_allclose(m1, m2, rtol=1e-12)
    np.testing.assert_allclose(v1, v2, rtol=1e-9)


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10: pins the checker of the device candidate generator."""
    r = O.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(x[0]) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    r = O.philox4x32_10([f], [f], [f], [f], f, f)
    assert [int(x[0]) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = O.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(x[0]) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    c = O.generate_candidates(5, 0, 1000, 700, np.array([-5.0, 0, 1]), np.array([10.0, 15, 2]), np.array([0.0, 1, 1.5]), 0.1)
    assert c.shape == (1000, 3) and np.all(c >= [-5, 0, 1]) and np.all(c <= [10, 15, 2])
    assert abs(c[700:, 0].std() - 0.1) < 0.02


# ----------------------------------------------------------------------------------------------
# threaded C restatement (oracle/kmat.c) used for the BASELINE-size checks: pinned to the numpy oracle
# ----------------------------------------------------------------------------------------------
def _prod1d(theta, D, cls):
    k = G.ConstantKernel(theta[0], ndim=D)
    for d in range(D):
        k = G.Product(k, cls(np.exp(theta[1 + d:2 + d]), ndim=D, axes=d))
    return k


@pytest.mark.parametrize("case", ["ard_matern52", "ard_rbf", "ard_matern32", "prod1d_matern52", "iso_matern52"])
def test_c_kernel_matrix_matches_numpy_oracle(case):
    rng = np.random.RandomState(7)
    D = 5
    X1, X2 = rng.rand(37, D) * 3 - 1, rng.rand(53, D) * 3 - 1
    theta = np.concatenate(([0.3], rng.uniform(-2, 1, D)))
    if case == "ard_matern52":
        k = O.make_kernel("matern52", D, theta)
    elif case == "ard_rbf":
        k = O.make_kernel("rbf", D, theta)
    elif case == "ard_matern32":
        k = G.Product(G.ConstantKernel(theta[0], ndim=D), G.Matern32Kernel(np.exp(theta[1:]), ndim=D))
    elif case == "prod1d_matern52":
        k = _prod1d(theta, D, G.Matern52Kernel)
    else:
        k = 2.0 * G.Matern52Kernel(0.7, ndim=D)
    ref = k.get_value(X1, X2)
    got = O.kmat_fast(k, X1, X2)
    assert np.max(np.abs(got - ref) / ref) <= 8 * np.finfo(float).eps      # libm exp vs numpy exp, a few ulp
    f = O.flatten_kernel(k)
    assert len(f["axis"]) == D and f["last"][-1] == 1


def test_fast_oracle_paths_match_reference_faithful_oracle():
    """gp_predict_var_only_fast (BASELINE-size candidate batches) and gp_grad_nll_terms_fast (config 5) against the
    reference-faithful restatements they accelerate."""
    X, y, Xs, theta, noise = O.synthetic_problem(200, 4, 700)
    st = O.gp_fit(O.make_kernel("matern52", 4, theta), X, y, noise=noise, normalize_input=True, normalize_output=True,
                  lower=np.zeros(4) - 0.1, upper=np.ones(4) + 0.2)
    mu, var = O.gp_predict(st, Xs)
    mu2, var2 = O.gp_predict_var_only_fast(st, Xs, chunk=256)
    np.testing.assert_allclose(mu2, mu, rtol=0, atol=1e-11 * np.abs(mu).max())
    np.testing.assert_allclose(var2, var, rtol=1e-9, atol=1e-13)
    th = np.append(theta, np.log(noise)) + 0.05
    g = O.gp_grad_nll_correct(st, th)
    g2 = O.gp_grad_nll_terms_fast(st, th)
    np.testing.assert_allclose(g2, g, rtol=1e-10, atol=1e-10)
    # product of 1-D kernels: one entry per term, george parameter order
    t3 = np.array([0.2, -0.5, 0.1, -1.0])
    k = _prod1d(t3, 3, G.Matern52Kernel)
    X3 = np.random.RandomState(1).rand(150, 3)
    y3 = np.sin(X3.sum(axis=1))
    st3 = O.gp_fit(k, X3, y3, noise=1e-3, normalize_input=False)
    th3 = np.append(t3, np.log(1e-3))
    np.testing.assert_allclose(O.gp_grad_nll_terms_fast(st3, th3), O.gp_grad_nll_correct(st3, th3), rtol=1e-10, atol=1e-10)
