#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference classes.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

What runs verbatim from /root/reference (read-only, via sys.path):
  robo/models/gaussian_process.py   (GaussianProcess.train/predict/nll/optimize/...)
  robo/models/base_model.py, robo/util/normalization.py
  robo/acquisition_functions/{ei,log_ei,pi,lcb}.py
  robo/priors/{base_prior,default_priors}.py
What is restated: ``george`` (absent, un-vendored; see oracle/george_oracle.py),
registered as sys.modules['george'] before the reference is imported.
One numpy-2 compatibility shim: ``np.Infinity`` (used at log_ei.py:89,96,118,
removed in numpy 2.0) is aliased to ``np.inf``.

Every case also asserts that oracle.robo_oracle (the restatement that travels
to the GPU box) reproduces the reference run; so the committed vectors pin the
restatement to the reference's own code.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("ROBO_REFERENCE", "/root/reference")

from oracle import george_oracle as G          # noqa: E402
from oracle import robo_oracle as O            # noqa: E402

G.install_as_george()
if not hasattr(np, "Infinity"):
    np.Infinity = np.inf
sys.path.insert(0, REF)

from robo.models.gaussian_process import GaussianProcess       # noqa: E402
from robo.acquisition_functions.ei import EI                   # noqa: E402
from robo.acquisition_functions.log_ei import LogEI            # noqa: E402
from robo.acquisition_functions.pi import PI                   # noqa: E402
from robo.acquisition_functions.lcb import LCB                 # noqa: E402
from robo.priors.default_priors import DefaultPrior, TophatPrior   # noqa: E402
from robo.models.base_model import BaseModel                   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def same(a, b, what, rtol=1e-13, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ok = np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    if not ok:
        raise AssertionError("restatement != reference for %s: max abs diff %g"
                             % (what, np.nanmax(np.abs(a - b))))


def branin(x):
    x1, x2 = x[:, 0], x[:, 1]
    return (x2 - 5.1 / (4 * np.pi ** 2) * x1 ** 2 + 5 / np.pi * x1 - 6) ** 2 \
        + 10 * (1 - 1 / (8 * np.pi)) * np.cos(x1) + 10


def run_gp_case(name, kernel_fn, X, y, Xs, noise, normalize_input, normalize_output,
                lower, upper, prior_fn=None, thetas=None, full_cov_m=16):
    """Fit the reference GaussianProcess (do_optimize=False) and record everything."""
    kernel = kernel_fn()
    prior = prior_fn(len(kernel) + 1) if prior_fn else None
    model = GaussianProcess(kernel, prior=prior, noise=noise,
                            normalize_input=normalize_input,
                            normalize_output=normalize_output,
                            lower=lower, upper=upper,
                            rng=np.random.RandomState(0))
    model.train(X, y, do_optimize=False)
    mu, var = model.predict(Xs)
    mu_c, cov = model.predict(Xs[:full_cov_m], full_cov=True)
    inc_x, inc_y = model.get_incumbent()
    pv = model.predict_variance(Xs[:1], Xs[1:9])
    ll = model.gp.log_likelihood(model.y, quiet=True)
    logdet = model.gp.solver.log_determinant
    acq = dict(ei=EI(model).compute(Xs),
               log_ei=LogEI(model).compute(Xs),
               pi=PI(model).compute(Xs),
               lcb=LCB(model).compute(Xs),
               ei_par=EI(model, par=0.1).compute(Xs),
               lcb_par=LCB(model, par=2.5).compute(Xs),
               ei_eta=EI(model).compute(Xs, eta=float(np.median(y))))
    hypers = np.array(model.hypers, dtype=np.float64)

    # restatement must agree with the reference run
    st = O.gp_fit(kernel_fn(), X, y, noise=noise, normalize_input=normalize_input,
                  normalize_output=normalize_output, lower=lower, upper=upper)
    o_mu, o_var = O.gp_predict(st, Xs)
    same(o_mu, mu, name + ".mu")
    same(o_var, var, name + ".var")
    o_mu_c, o_cov = O.gp_predict(st, Xs[:full_cov_m], full_cov=True)
    same(o_cov, cov, name + ".cov")
    same(O.gp_get_incumbent(st)[1], inc_y, name + ".inc_y")
    same(O.gp_get_incumbent(st)[0], inc_x, name + ".inc_x")
    same(O.gp_predict_variance(st, Xs[:1], Xs[1:9]), pv, name + ".predict_variance")
    o_ll, o_logdet = O.gp_loglik_terms(st)
    same(o_ll, ll, name + ".ll")
    same(o_logdet, logdet, name + ".logdet")
    same(O.acquisition(st, Xs, "ei"), acq["ei"], name + ".ei")
    same(O.acquisition(st, Xs, "log_ei"), acq["log_ei"], name + ".log_ei")
    same(O.acquisition(st, Xs, "pi"), acq["pi"], name + ".pi")
    same(O.acquisition(st, Xs, "lcb"), acq["lcb"], name + ".lcb")
    same(O.acquisition(st, Xs, "ei", par=0.1), acq["ei_par"], name + ".ei_par")
    same(O.acquisition(st, Xs, "lcb", par=2.5), acq["lcb_par"], name + ".lcb_par")
    same(O.acquisition(st, Xs, "ei", eta=float(np.median(y))), acq["ei_eta"], name + ".ei_eta")

    nll_thetas = np.zeros((0, len(hypers)))
    nll_vals = np.zeros(0)
    if thetas is not None:
        nll_thetas = np.asarray(thetas, dtype=np.float64)
        nll_vals = np.array([model.nll(t) for t in nll_thetas])
        st2 = O.gp_fit(kernel_fn(), X, y, noise=noise, normalize_input=normalize_input,
                       normalize_output=normalize_output, lower=lower, upper=upper)
        o_nll = np.array([O.gp_nll(st2, t, prior) for t in nll_thetas])
        same(o_nll, nll_vals, name + ".nll")

    np.savez(os.path.join(OUT, name + ".npz"),
             X=X, y=y, Xs=Xs, noise=noise,
             normalize_input=normalize_input, normalize_output=normalize_output,
             lower=np.array([]) if lower is None else lower,
             upper=np.array([]) if upper is None else upper,
             hypers=hypers, mu=mu, var=var, cov=cov, full_cov_m=full_cov_m,
             inc_x=inc_x, inc_y=inc_y, predict_variance=pv, ll=ll, logdet=logdet,
             nll_thetas=nll_thetas, nll_vals=nll_vals,
             **{"acq_" + k: v for k, v in acq.items()})
    print("wrote %-18s N=%d D=%d M=%d  ll=%.12g" % (name, X.shape[0], X.shape[1], len(Xs), ll))


def case_reference_unit_test():
    """test/test_models/test_gaussian_process.py:13-49 with a fixed seed."""
    rng = np.random.RandomState(11)
    X = rng.rand(10, 2)
    y = np.sinc(X * 10 - 5).sum(axis=1)
    Xs = rng.rand(10, 2)
    run_gp_case("gp_unit", lambda: G.Matern52Kernel(np.ones(2), ndim=2), X, y, Xs,
                noise=1e-3, normalize_input=False, normalize_output=False,
                lower=None, upper=None, prior_fn=lambda n: TophatPrior(-2, 2),
                thetas=[[0.2, 0.2, 0.001], [-1.0, 0.5, -3.0], [25.0, 0.0, 0.0]],
                full_cov_m=10)
    # the reference's only known-answer on the posterior (:44-49)
    d = np.load(os.path.join(OUT, "gp_unit.npz"))
    k = G.Matern52Kernel(np.ones(2), ndim=2)
    import scipy.linalg as spla
    K_zz = k.get_value(Xs)
    K_zx = k.get_value(Xs, X)
    K_nz = k.get_value(X) + 1e-3 * np.eye(10)
    K_zz_x = K_zz - np.dot(K_zx, np.inner(spla.inv(K_nz), K_zx))
    # clip like predict does before comparing
    assert np.mean((np.clip(K_zz_x, O.EPS, np.inf) - d["cov"]) ** 2) < 10e-5


def case_branin(normalize_output):
    rng = np.random.RandomState(5)
    lower = np.array([-5.0, 0.0])
    upper = np.array([10.0, 15.0])
    X = lower + (upper - lower) * rng.rand(40, 2)
    y = branin(X)
    Xs = lower + (upper - lower) * rng.rand(300, 2)

    def kern():
        k = 2 * G.Matern52Kernel(np.ones(2), ndim=2)     # fmin/bayesian_optimization.py:75-81
        k.set_parameter_vector(np.array([np.log(1.7), np.log(0.15), np.log(0.4)]))
        return k
    thetas = [[0.5, -2.0, -1.0, -6.0], [0.0, 0.0, 0.0, -3.0], [2.0, -4.0, 1.5, -9.0]]
    run_gp_case("gp_branin_ny%d" % int(normalize_output), kern, X, y, Xs, noise=1e-3 if normalize_output else 1e-2,
                normalize_input=True, normalize_output=normalize_output,
                lower=lower, upper=upper, prior_fn=DefaultPrior, thetas=thetas)


def case_default_bounds():
    """lower/upper None -> column min/max of the training X (normalization.py:6-9)."""
    rng = np.random.RandomState(9)
    X = rng.randn(30, 3) * 2 + 1
    y = np.sin(X).sum(axis=1)
    Xs = rng.randn(50, 3) * 2 + 1
    run_gp_case("gp_autobounds", lambda: 1.0 * G.Matern52Kernel(np.array([0.3, 0.5, 0.7]), ndim=3),
                X, y, Xs, noise=1e-4, normalize_input=True, normalize_output=True,
                lower=None, upper=None)


def case_rbf():
    X, y, Xs, theta, noise = O.synthetic_problem(200, 8, 128, 21, 22)
    run_gp_case("gp_rbf_d8", lambda: O.make_kernel("rbf", 8, theta), X, y, Xs, noise=noise,
                normalize_input=False, normalize_output=False, lower=None, upper=None,
                thetas=[np.append(theta, np.log(noise)).tolist()])


def case_prod1d():
    """Product of 1-D Matern-5/2 kernels with amplitude (fabolas.py:100-110, config part)."""
    rng = np.random.RandomState(3)
    X = rng.rand(150, 3)
    y = np.cos(3 * X).prod(axis=1)
    Xs = rng.rand(100, 3)

    def kern():
        k = 1
        for d in range(3):
            k *= G.Matern52Kernel(np.ones([1]) * [0.2, 0.05, 0.6][d], ndim=3, axes=d)
        return k
    run_gp_case("gp_prod1d", kern, X, y, Xs, noise=1e-3, normalize_input=False,
                normalize_output=False, lower=None, upper=None,
                thetas=[[-0.5, -1.0, -2.0, 0.3, -5.0]])


def case_mid():
    X, y, Xs, theta, noise = O.synthetic_problem(512, 16, 256)
    run_gp_case("gp_mid_d16", lambda: O.make_kernel("matern52", 16, theta), X, y, Xs, noise=noise,
                normalize_input=True, normalize_output=False,
                lower=np.zeros(16), upper=np.ones(16),
                thetas=[np.append(theta, np.log(noise)).tolist()])


class _MomentsModel(BaseModel):
    """Stub returning fixed (m, v): drives the reference acquisition classes
    through every branch of log_ei.py:85-120 (cf. test/dummy_model.py)."""

    def __init__(self, m, v, eta):
        self.m, self.v, self.eta = m, v, eta

    def train(self, X, y):
        pass

    def predict(self, X):
        return self.m.copy(), self.v.copy()

    def get_incumbent(self):
        return None, self.eta


def case_acq_moments():
    rng = np.random.RandomState(17)
    eta = 0.25
    m = np.concatenate((rng.randn(200) * 2,            # generic
                        [eta, eta, eta - 1.0, eta + 1.0, eta + 1e-300, eta - 1e-9],
                        eta + np.array([30.0, 10.0, 3.0, -3.0, -10.0, -30.0]) * 0.1,
                        rng.randn(50) * 1e-3 + eta))
    v = np.concatenate((rng.rand(200) * 3 + 1e-6,
                        [0.5, 0.0, 0.0, 0.0, 1e-4, 1e-20],
                        np.full(6, 0.01),
                        rng.rand(50) * 1e-8 + O.EPS))
    X = np.zeros((m.size, 1))
    out = {}
    for par in (0.0, 0.3):
        model = _MomentsModel(m, v, eta)
        le = LogEI(model, par=par).compute(X)
        same(O.acq_log_ei(m, v, eta, par), le, "acq_moments.log_ei")
        out["log_ei_par%g" % par] = le
        pi = PI(model, par=par).compute(X)
        same(O.acq_pi(np.where(v > 0, m, m), np.where(v > 0, v, v), eta, par), pi, "acq_moments.pi")
        out["pi_par%g" % par] = pi
        lcb = LCB(model, par=1.0 + par).compute(X)
        same(O.acq_lcb(m, v, 1.0 + par), lcb, "acq_moments.lcb")
        out["lcb_par%g" % (1.0 + par)] = lcb
        # EI: the reference returns [[0]] for the whole batch if any s == 0 (ei.py:72-74)
        ei_q = EI(model, par=par).compute(X)
        assert ei_q.shape == (1, 1) and ei_q[0, 0] == 0
        pos = v > 0
        model_pos = _MomentsModel(m[pos], v[pos], eta)
        ei = EI(model_pos, par=par).compute(X[pos])
        same(O.acq_ei(m[pos], v[pos], eta, par), ei, "acq_moments.ei")
        out["ei_pos_par%g" % par] = ei
    np.savez(os.path.join(OUT, "acq_moments.npz"), m=m, v=v, eta=eta, **out)
    print("wrote acq_moments       n=%d" % m.size)


def case_optimize(cov_amp, name):
    """train(do_optimize=True): L-BFGS-B on nll (gaussian_process.py:193-219).

    cov_amp=2 is the fmin default (fmin/bayesian_optimization.py:75): with ndim=2 george's
    ``c * kernel`` gives log_constant = log(2/2) = 0, where LognormalPrior.lnprob is -inf
    (base_prior.py:278), so the reference starts L-BFGS-B at nll = 1e25 (degenerate case,
    kept because it is what the default facade does).  cov_amp=3 starts at a finite point."""
    rng = np.random.RandomState(2)
    lower = np.array([-5.0, 0.0])
    upper = np.array([10.0, 15.0])
    X = lower + (upper - lower) * rng.rand(25, 2)
    y = branin(X)
    kernel = cov_amp * G.Matern52Kernel(np.ones(2), ndim=2)
    prior = DefaultPrior(len(kernel) + 1, rng=np.random.RandomState(0))
    model = GaussianProcess(kernel, prior=prior, normalize_input=True, lower=lower, upper=upper,
                            rng=np.random.RandomState(0))
    model.train(X, y, do_optimize=True)
    theta = np.array(model.hypers)
    nll_opt = model.nll(theta)
    p0 = np.append((cov_amp * G.Matern52Kernel(np.ones(2), ndim=2)).get_parameter_vector(), np.log(1e-3))
    nll_p0 = model.nll(p0)
    model.gp.kernel.set_parameter_vector(theta[:-1])
    model.gp.compute(model.X, yerr=np.sqrt(model.noise))
    Xs = lower + (upper - lower) * rng.rand(64, 2)
    mu, var = model.predict(Xs)
    np.savez(os.path.join(OUT, name + ".npz"), cov_amp=cov_amp, X=X, y=y, Xs=Xs, lower=lower, upper=upper,
             theta_opt=theta, nll_opt=nll_opt, p0=p0, nll_p0=nll_p0, mu=mu, var=var,
             noise=model.noise)
    print("wrote " + name + "  theta*=%s nll*=%.10g nll(p0)=%.10g" % (np.round(theta, 4), nll_opt, nll_p0))


if __name__ == "__main__":
    case_reference_unit_test()
    case_branin(False)
    case_branin(True)
    case_default_bounds()
    case_rbf()
    case_prod1d()
    case_mid()
    case_acq_moments()
    case_optimize(2, 'gp_optimize_default')
    case_optimize(3, 'gp_optimize')
    print("golden vectors written to", OUT)
