"""Host-side logic of the product (robo_b200/*.py) on the oracle-backed FakeHandle (tests/fake_gpk.py):
everything above the C ABI that the GPU suite also runs, so that the CPU test tier covers it too."""
import copy

import numpy as np
import pytest

from oracle import robo_oracle as O


@pytest.fixture
def fake(monkeypatch):
    from tests import fake_gpk
    return fake_gpk.install(monkeypatch)


def branin(x):
    x1, x2 = x[0], x[1]
    return (x2 - 5.1 / (4 * np.pi ** 2) * x1 ** 2 + 5 / np.pi * x1 - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(x1) + 10


def test_fmin_facade_gp_and_gp_mcmc(fake):
    from robo_b200.fmin import bayesian_optimization
    lower, upper = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    res = bayesian_optimization(branin, lower, upper, num_iterations=12, maximizer="random", acquisition_func="ei",
                                model_type="gp", n_init=3, rng=np.random.RandomState(0))
    assert len(res["y"]) == 12 and res["f_opt"] == min(res["y"]) and np.all(np.diff(res["incumbent_values"]) <= 0)
    assert np.all(np.array(res["X"]) >= lower) and np.all(np.array(res["X"]) <= upper)
    res = bayesian_optimization(branin, lower, upper, num_iterations=5, n_init=3, chain_length=6, burnin_steps=4,
                                rng=np.random.RandomState(1))          # default: gp_mcmc + log_ei
    assert len(res["y"]) == 5
    with pytest.raises(ValueError):
        bayesian_optimization(branin, lower, upper, num_iterations=4, model_type="rf")
    with pytest.raises(AssertionError):
        bayesian_optimization(branin, upper, lower, num_iterations=4)


def test_gp_mcmc_model_and_marginalisation(fake):
    """test/test_models/test_gaussian_process_mcmc.py + test_marginalization.py shape contracts, batched ==
    sequential log-likelihood, mixture moments."""
    from robo_b200 import kernels as K
    from robo_b200.acquisition_functions import EI, LCB, PI, LogEI, MarginalizationGPMCMC
    from robo_b200.models import GaussianProcessMCMC
    from robo_b200.priors import DefaultPrior
    rng = np.random.RandomState(0)
    X = rng.rand(10, 2)
    y = np.sinc(X * 10 - 5).sum(axis=1)
    kernel = 2 * K.Matern52Kernel(np.ones(2), ndim=2)
    model = GaussianProcessMCMC(kernel, prior=DefaultPrior(len(kernel) + 1, rng=np.random.RandomState(1)), n_hypers=8,
                                chain_length=5, burnin_steps=5, normalize_input=False, rng=np.random.RandomState(2))
    model.train(X, y, do_optimize=True)
    assert len(model.models) == 8 and np.asarray(model.hypers).shape == (8, 4) and model.burned
    p0 = model.p0.copy()
    model.train(X, y, do_optimize=True)                     # second call continues from the stored walkers
    assert model.p0.shape == p0.shape
    Xt = rng.rand(7, 2)
    m, v = model.predict(Xt)
    mus = np.array([s.predict(Xt)[0] for s in model.models])
    vs = np.array([s.predict(Xt)[1] for s in model.models])
    mr, vr = O.mcmc_mixture_moments(mus, vs)
    np.testing.assert_allclose(m, mr)
    np.testing.assert_allclose(v, vr)
    thetas = model.p0[:4]
    from robo_b200.models.gaussian_process_mcmc import _LikelihoodPool
    model._pool = _LikelihoodPool(kernel, model.X, model.y, model.mean, 4)
    np.testing.assert_allclose(model.loglikelihood_batch(thetas), [model.loglikelihood(t) for t in thetas])
    assert model.loglikelihood(np.array([25.0, 0, 0, 0])) == -np.inf
    for cls in (EI, LogEI, PI, LCB):
        acq = MarginalizationGPMCMC(cls(model))
        acq.update(model)
        assert acq._fused_spec() is not None          # device sub-models: ONE multi-model call (gpk_acq_multi)
        a = acq.compute(Xt)
        assert a.shape == (7,)
        np.testing.assert_allclose(a, np.mean([cls(s).compute(Xt) for s in model.models], axis=0))
        assert acq.argmax(Xt) == int(np.argmax(a))
        # a cost model, or estimators with different parameters, fall back to the reference's loop over estimators
        acq.estimators[0].par = 0.25
        assert acq._fused_spec() is None
        b = acq.compute(Xt)
        assert b.shape == (7,)
        acq.estimators[0].par = acq.estimators[1].par
    clone = copy.deepcopy(model)
    np.testing.assert_allclose(clone.predict(Xt)[0], m)
    model.train(X, y, do_optimize=False)
    assert len(model.models) == 1


def test_model_semantics_on_fake_handle(fake):
    """reference behaviours kept by the host layer: untrained predict, shape asserts, y_std == 0, noise x10
    retry, nll guards, update(), EI quirks, deepcopy."""
    from robo_b200 import kernels as K
    from robo_b200.acquisition_functions import EI, LogEI
    from robo_b200.device_gp import DeviceGP
    from robo_b200.models import GaussianProcess
    rng = np.random.RandomState(3)
    X, y = rng.rand(12, 2), rng.rand(12)
    model = GaussianProcess(K.Matern52Kernel(np.ones(2), ndim=2), normalize_output=True, normalize_input=False)
    with pytest.raises(Exception, match="trained first"):
        model.predict(X)
    with pytest.raises(ValueError, match="same value"):
        model.train(X, np.ones(12), do_optimize=False)
    model.train(X, y, do_optimize=False)
    assert model.nll(np.array([21.0, 0.0, 0.0])) == 1e25
    assert np.isfinite(model.nll(np.array([0.1, 0.2, -3.0])))
    mu, var = model.predict(X)
    assert mu.shape == (12,) and var.shape == (12,) and np.all(var >= np.finfo(float).eps)
    assert model.predict(X, full_cov=True)[1].shape == (12, 12)
    assert model.predict_variance(X[:1], X[1:5]).shape == (4, 1)
    assert model.sample_functions(X[:6], n_funcs=3).shape == (3, 6)
    inc, inc_val = model.get_incumbent()
    assert inc_val == pytest.approx(y.min())
    model.update(X[:2], model.y[:2])
    assert model.X.shape[0] == 14
    clone = copy.deepcopy(model)
    np.testing.assert_allclose(clone.predict(X)[0], model.predict(X)[0])
    # LinAlgError -> noise x 10 -> retry (gaussian_process.py:118-122)
    calls = []
    real = DeviceGP.compute

    def flaky(self, x=None, yerr=0.0, **kw):
        calls.append(yerr)
        if len(calls) == 1:
            raise np.linalg.LinAlgError("not positive definite")
        return real(self, x, yerr=yerr)
    import unittest.mock as um
    with um.patch.object(DeviceGP, "compute", flaky):
        m2 = GaussianProcess(K.Matern52Kernel(np.ones(2), ndim=2), noise=1e-3, normalize_input=False)
        m2.train(X, y, do_optimize=False)
    assert m2.noise == pytest.approx(1e-2) and len(calls) == 2
    # optimise
    m3 = GaussianProcess(2 * K.Matern52Kernel(np.ones(2), ndim=2), normalize_input=False)
    m3.train(X, y, do_optimize=True)
    assert m3.hypers.shape == (4,) and m3.noise == pytest.approx(np.exp(m3.hypers[-1]))
    # acquisition: derivative of LogEI returns None like the reference
    assert LogEI(model).compute(X, derivative=True) is None
    assert EI(model).compute(X).shape == (12,)


def test_device_random_sampling_host_logic(fake):
    from robo_b200 import kernels as K
    from robo_b200.acquisition_functions import EI
    from robo_b200.maximizers import DeviceRandomSampling, RandomSampling
    from robo_b200.models import GaussianProcess
    rng = np.random.RandomState(0)
    lower, upper = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    X = lower + (upper - lower) * rng.rand(15, 2)
    y = np.array([branin(x) for x in X])
    model = GaussianProcess(2 * K.Matern52Kernel(np.ones(2), ndim=2), lower=lower, upper=upper)
    model.train(X, y, do_optimize=False)
    acq = EI(model)
    mx = DeviceRandomSampling(acq, lower, upper, n_samples=400, rng=np.random.RandomState(1))
    x = mx.maximize()
    cand = O.generate_candidates(mx.last["seed"], 0, 400, 280, lower, upper, model.get_incumbent()[0], 0.1)
    assert np.array_equal(x, cand[int(np.argmax(acq.compute(cand)))])
    x2 = mx.maximize()
    assert mx.calls == 2 and not np.array_equal(x, x2)
    # candidate count of the reference: int(0.7 n) + int(0.3 n)  (n = 5 -> 3 + 1), random_sampling.py:38-47
    mx5 = DeviceRandomSampling(acq, lower, upper, n_samples=5, rng=np.random.RandomState(2))
    x5 = mx5.maximize()
    cand5 = O.generate_candidates(mx5.last["seed"], 0, 4, 3, lower, upper, model.get_incumbent()[0], 0.1)
    assert np.array_equal(x5, cand5[int(np.argmax(acq.compute(cand5)))])
    rs = RandomSampling(acq, lower, upper, n_samples=100, rng=np.random.RandomState(0))
    xr = rs.maximize()
    assert xr.shape == (2,) and np.all(xr >= lower) and np.all(xr <= upper)


def test_incremental_refit_host_logic(fake):
    """train(do_optimize=False) with rows appended to a factorised training set goes through
    Handle.fit_append when (and only when) the library's preconditions hold; results equal a fresh model's."""
    from robo_b200 import kernels as K
    from robo_b200.models import GaussianProcess
    rng = np.random.RandomState(3)
    X, y = rng.rand(400, 3), rng.rand(400)
    lower, upper = np.zeros(3), np.ones(3)
    Xt = rng.rand(9, 3)

    def fresh(n, noise=1e-3):
        m = GaussianProcess(2.0 * K.Matern52Kernel(np.ones(3), ndim=3), noise=noise, lower=lower, upper=upper)
        m.train(X[:n], y[:n], do_optimize=False)
        return m

    model = fresh(300)
    h = model.gp.handle
    assert (h.n_fits, h.n_appends) == (1, 0)
    # 1. no scoring call since the fit -> L^-1 not built -> full refit
    model.train(X[:303], y[:303], do_optimize=False)
    assert (h.n_fits, h.n_appends, model.gp.n_appends) == (2, 0, 0)
    # 2. after a predict the appended rows take the shortcut; BaseModel.update goes the same way
    model.predict(Xt)
    model.train(X[:310], y[:310], do_optimize=False)
    assert (h.n_fits, h.n_appends, model.gp.n_appends) == (2, 1, 1)
    mu, var = model.predict(Xt)
    mu_f, var_f = fresh(310).predict(Xt)
    np.testing.assert_allclose(mu, mu_f, rtol=1e-12)
    np.testing.assert_allclose(var, var_f, rtol=1e-12)
    assert model.gp.log_likelihood(model.y) == fresh(310).gp.log_likelihood(y[:310])
    # 3. crossing a 128-row block boundary, changed noise, changed earlier rows, fewer rows: full refits
    model.train(X[:390], y[:390], do_optimize=False)
    assert (h.n_fits, h.n_appends) == (3, 1)
    model.predict(Xt)
    model.noise = 2e-3
    model.train(X[:392], y[:392], do_optimize=False)
    assert (h.n_fits, h.n_appends) == (4, 1)
    model.predict(Xt)
    X2 = X[:394].copy()
    X2[0, 0] += 1e-3
    model.train(X2, y[:394], do_optimize=False)
    assert (h.n_fits, h.n_appends) == (5, 1)
    model.predict(Xt)
    model.train(X[:391], y[:391], do_optimize=False)
    assert (h.n_fits, h.n_appends) == (6, 1)
    # 4. the switch
    model.predict(Xt)
    model.gp.incremental = False
    model.train(X[:393], y[:393], do_optimize=False)
    assert (h.n_fits, h.n_appends) == (7, 1)
    # 5. a deep copy (MarginalizationGPMCMC copies models) never appends onto a handle it does not own
    m2 = copy.deepcopy(fresh(300))
    m2.predict(Xt)
    m2.train(X[:305], y[:305], do_optimize=False)
    np.testing.assert_allclose(m2.predict(Xt)[0], fresh(305).predict(Xt)[0], rtol=1e-12)
