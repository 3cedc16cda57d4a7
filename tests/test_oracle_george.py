"""Pins the george restatement (oracle/george_oracle.py) to independent
implementations: sklearn kernels, mpmath 50-digit GP algebra, finite differences,
and the reference's only known-answer formula (test_gaussian_process.py:44-49)."""
import numpy as np
import pytest
import scipy.linalg as spla

from oracle import george_oracle as G
from oracle import robo_oracle as O


def test_matern52_matches_sklearn():
    from sklearn.gaussian_process.kernels import Matern
    rng = np.random.RandomState(0)
    X1, X2 = rng.rand(30, 4), rng.rand(20, 4)
    metric = np.array([0.3, 1.0, 2.5, 0.07])
    k = G.Matern52Kernel(metric, ndim=4)
    ref = Matern(length_scale=np.sqrt(metric), nu=2.5)(X1, X2)
    np.testing.assert_allclose(k.get_value(X1, X2), ref, rtol=2e-15, atol=1e-16)


def test_expsquared_matches_sklearn():
    from sklearn.gaussian_process.kernels import RBF
    rng = np.random.RandomState(1)
    X1 = rng.rand(25, 3)
    metric = np.array([0.5, 0.2, 1.5])
    k = G.ExpSquaredKernel(metric, ndim=3)
    np.testing.assert_allclose(k.get_value(X1), RBF(length_scale=np.sqrt(metric))(X1),
                               rtol=2e-15, atol=1e-16)


def test_scalar_times_kernel_layout():
    # fmin/bayesian_optimization.py:79-85: len(kernel) == D + 1, theta[0] = log amplitude
    k = 2 * G.Matern52Kernel(np.ones(3), ndim=3)
    assert len(k) == 4
    np.testing.assert_allclose(k.get_parameter_vector(), [np.log(2.0 / 3), 0, 0, 0])
    k.set_parameter_vector([0.3, -1, -2, -3])
    np.testing.assert_allclose(k[:], [0.3, -1, -2, -3])
    X = np.random.RandomState(0).rand(5, 3)
    base = G.Matern52Kernel(np.exp([-1, -2, -3]), ndim=3).get_value(X)
    np.testing.assert_allclose(k.get_value(X), np.exp(0.3) * base, rtol=1e-15)


def test_axes_product_equals_separable():
    X = np.random.RandomState(3).rand(12, 3)
    k = 1
    for d, m in enumerate([0.2, 0.05, 0.6]):
        k *= G.Matern52Kernel(np.ones(1) * m, ndim=3, axes=d)
    expect = np.ones((12, 12)) / 3.0
    for d, m in enumerate([0.2, 0.05, 0.6]):
        expect *= G.Matern52Kernel(m, ndim=1).get_value(X[:, d:d + 1])
    np.testing.assert_allclose(k.get_value(X), expect, rtol=1e-15)


@pytest.mark.parametrize("kind", ["matern52", "rbf"])
def test_kernel_gradient_finite_difference(kind):
    rng = np.random.RandomState(4)
    X = rng.rand(9, 3)
    theta = np.array([0.4, -0.7, 0.2, -1.3])
    k = O.make_kernel(kind, 3, theta)
    g = k.gradient(X)
    assert g.shape == (9, 9, 4)
    h = 1e-6
    for p in range(4):
        tp, tm = theta.copy(), theta.copy()
        tp[p] += h
        tm[p] -= h
        fd = (O.make_kernel(kind, 3, tp).get_value(X) - O.make_kernel(kind, 3, tm).get_value(X)) / (2 * h)
        np.testing.assert_allclose(g[:, :, p], fd, rtol=1e-7, atol=1e-9)


def test_gp_against_mpmath_50_digits():
    """compute / log_likelihood / predict against 50-digit arithmetic."""
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.RandomState(7)
    N, D, M = 12, 2, 5
    X, Xs = rng.rand(N, D), rng.rand(M, D)
    y = np.sinc(X * 10 - 5).sum(axis=1)
    metric = [0.4, 0.9]
    amp, noise, mean = 1.3, 1e-3, float(np.mean(y))

    def kmp(a, b):
        r2 = sum((mp.mpf(a[d]) - mp.mpf(b[d])) ** 2 / mp.mpf(metric[d]) for d in range(D))
        r = mp.sqrt(5 * r2)
        return mp.mpf(amp) * (1 + r + 5 * r2 / 3) * mp.exp(-r)

    yerr2 = mp.mpf(float(np.sqrt(noise))) ** 2 + mp.mpf(G.TINY)
    K = mp.matrix(N, N)
    for i in range(N):
        for j in range(N):
            K[i, j] = kmp(X[i], X[j]) + (yerr2 if i == j else 0)
    r = mp.matrix([mp.mpf(v) - mp.mpf(mean) for v in y])
    Kinv_r = mp.lu_solve(K, r)
    ll = -(r.T * Kinv_r)[0] / 2 - mp.log(mp.det(K)) / 2 - N * mp.log(2 * mp.pi) / 2
    Ks = mp.matrix(M, N)
    for i in range(M):
        for j in range(N):
            Ks[i, j] = kmp(Xs[i], X[j])
    mu = Ks * Kinv_r
    var = [kmp(Xs[i], Xs[i]) - (Ks[i, :] * mp.lu_solve(K, Ks[i, :].T))[0] for i in range(M)]

    k = G.Product(G.ConstantKernel(np.log(amp), ndim=D), G.Matern52Kernel(metric, ndim=D))
    gp = G.GP(k, mean=mean)
    gp.compute(X, yerr=np.sqrt(noise))
    assert abs(gp.log_likelihood(y) - float(ll)) < 1e-10 * abs(float(ll))
    m_o, c_o = gp.predict(y, Xs)
    np.testing.assert_allclose(m_o, [float(v) + mean for v in mu], rtol=1e-11)
    np.testing.assert_allclose(np.diag(c_o), [float(v) for v in var], rtol=1e-9)


def test_reference_known_answer_formula():
    """test/test_models/test_gaussian_process.py:44-49 (noise on the training
    diagonal only, not on the predictive covariance)."""
    rng = np.random.RandomState(11)
    X = rng.rand(10, 2)
    y = np.sinc(X * 10 - 5).sum(axis=1)
    Xs = rng.rand(10, 2)
    kernel = G.Matern52Kernel(np.ones(2), ndim=2)
    st = O.gp_fit(kernel, X, y, noise=1e-3, normalize_input=False)
    _, v = O.gp_predict(st, Xs, full_cov=True)
    K_zz = kernel.get_value(Xs)
    K_zx = kernel.get_value(Xs, X)
    K_nz = kernel.get_value(X) + st["noise"] * np.eye(10)
    K_zz_x = K_zz - np.dot(K_zx, np.inner(spla.inv(K_nz), K_zx))
    assert np.mean((K_zz_x - v) ** 2) < 10e-5


def test_not_positive_definite_raises_linalgerror():
    X = np.zeros((4, 2))            # four identical points, no noise -> singular
    gp = G.GP(G.Matern52Kernel(np.ones(2), ndim=2), white_noise=-np.inf)
    with pytest.raises(np.linalg.LinAlgError):
        gp.compute(X, yerr=0.0)
