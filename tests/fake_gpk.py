"""Oracle-backed stand-in for robo_b200._lib.Handle — TEST INFRASTRUCTURE ONLY.

Lets the CPU test-suite (no GPU in the build container) drive the *host-side* product code
(robo_b200 models / acquisition functions / maximizers / compat shims) and the UNMODIFIED reference
(solver, fmin facade, maximizers, MarginalizationGPMCMC from /root/reference) through the exact
method surface of ``_lib.Handle``, with the arithmetic supplied by oracle/.  The GPU tests exercise
the same surface against libgpk.so.  Never imported by the product."""
import numpy as np
import scipy.linalg as spla

from oracle import george_oracle as G
from oracle import robo_oracle as O

FAMILIES = {0: G.Matern52Kernel, 1: G.ExpSquaredKernel, 2: G.Matern32Kernel}
ACQ_NAME = {1: "ei", 2: "log_ei", 3: "pi", 4: "lcb"}


class FakeHandle(object):
    def __init__(self, device=0):
        self.device = device
        self.X = self.y = self.kernel = None
        self.bounds = None
        self.out = (False, 0.0, 1.0)
        self.fitted = False
        self.n_fits = 0
        self.n_appends = 0
        self.linv_built = False

    def close(self):
        pass

    def set_option(self, key, value):
        pass

    def set_stream(self, s):
        pass

    def synchronize(self):
        pass

    def set_data(self, X, y):
        self.X, self.y = np.array(X, dtype=np.float64), np.array(y, dtype=np.float64)
        self.fitted = False

    def set_input_bounds(self, lower, upper):
        self.bounds = None if lower is None else (np.array(lower, float), np.array(upper, float))

    def set_output_transform(self, enabled, y_mean=0.0, y_std=1.0):
        self.out = (bool(enabled), float(y_mean), float(y_std))

    def set_kernel(self, family, log_amp, axis, group, log_metric):
        D = self.X.shape[1] if self.X is not None else int(max(axis)) + 1
        k = G.ConstantKernel(log_amp, ndim=D)
        axis, group, lm = np.asarray(axis), np.asarray(group), np.asarray(log_metric, dtype=float)
        for g in range(int(group.max()) + 1):
            sel = group == g
            k = G.Product(k, FAMILIES[int(family)](np.exp(lm[sel]), ndim=D, axes=axis[sel]))
        self.kernel, self.amp, self.fitted = k, float(np.exp(log_amp)), False
        self.linv_built = False
        self.spec = (int(family), float(log_amp), axis.copy(), group.copy(), lm.copy())

    def fit(self, diag_add, mean):
        self.fitted = False
        self.linv_built = False
        self.diag_add = diag_add
        self.n_fits += 1
        K = self.kernel.get_value(self.X)
        K[np.diag_indices_from(K)] += diag_add
        if not np.all(np.isfinite(K)):
            raise np.linalg.LinAlgError("not positive definite")
        self.L = spla.cholesky(K, lower=True)           # raises numpy.linalg.LinAlgError
        self.mean = float(mean)
        self.z = spla.solve_triangular(self.L, self.y - mean, lower=True)
        self.alpha = spla.solve_triangular(self.L, self.z, lower=True, trans="T")
        logdet = 2.0 * np.sum(np.log(np.diag(self.L)))
        ll = -0.5 * self.z @ self.z - 0.5 * logdet - 0.5 * len(self.y) * np.log(2 * np.pi)
        self.fitted = True
        return logdet, ll

    def fit_append(self, X, y, diag_add, mean):
        """Same preconditions as gpk_fit_append (fitted, L^-1 built by an earlier scoring call, same diagonal term,
        new rows inside the last 128-row block); the arithmetic is a plain refit."""
        X, y = np.array(X, dtype=np.float64), np.array(y, dtype=np.float64)
        n_old, n = len(self.y), len(y)
        NP = -(-n_old // 128) * 128
        if (not self.fitted or not self.linv_built or NP < 256 or -(-n // 128) * 128 != NP or n <= n_old
                or n_old <= NP - 128 or diag_add != self.diag_add):
            return None
        assert np.array_equal(X[:n_old], self.X)
        kernel, amp, spec = self.kernel, self.amp, self.spec
        self.X, self.y = X, y
        out = self.fit(diag_add, mean)
        self.n_fits -= 1
        self.n_appends += 1
        self.linv_built = True
        return out

    def fit_begin(self, diag_add, mean):
        try:
            self._pending = self.fit(diag_add, mean)
        except np.linalg.LinAlgError as e:
            self._pending = e

    def fit_end(self):
        p, self._pending = self._pending, None
        if isinstance(p, Exception):
            raise p
        return p

    def _norm(self, Xs):
        Xs = np.asarray(Xs, dtype=np.float64)
        return Xs if self.bounds is None else (Xs - self.bounds[0]) / (self.bounds[1] - self.bounds[0])

    def _moments(self, Xs, full=False, clip=True):
        if not self.fitted:
            raise RuntimeError("model not fitted")
        self.linv_built = True
        Xn = self._norm(Xs)
        Ks = self.kernel.get_value(Xn, self.X)
        mu = Ks @ self.alpha + self.mean
        V = spla.solve_triangular(self.L, Ks.T, lower=True)
        var = self.kernel.get_value(Xn) - V.T @ V if full else self.amp - np.einsum("ij,ij->j", V, V)
        on, ym, ys = self.out
        if on:
            mu, var = mu * ys + ym, var * ys ** 2
        return mu, (np.clip(var, O.EPS, np.inf) if clip else var)

    def posterior_cov(self, Xs):
        return self._moments(Xs, full=True, clip=False)

    def predict(self, Xs):
        return self._moments(Xs)

    def predict_cov(self, Xs):
        return self._moments(Xs, full=True)

    def acq(self, Xs, kind, eta=0.0, par=0.0, want_values=True, want_moments=False):
        mu, var = self._moments(Xs)
        vals, nneg = None, 0
        if kind != 0:
            name = ACQ_NAME[int(kind)]
            with np.errstate(all="ignore"):
                if name == "lcb":
                    vals = O.acq_lcb(mu, var, par)
                elif name == "ei":
                    s = np.sqrt(var)
                    z = (eta - mu - par) / s
                    from scipy.special import ndtr
                    vals = s * (z * ndtr(z) + np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi))
                    nneg = int((vals < 0).sum())
                else:
                    vals = O.ACQ[name](mu, var, eta, par)
        bi = int(np.argmax(vals)) if vals is not None else -1
        return dict(values=vals if want_values else None, mu=mu if want_moments else None,
                    var=var if want_moments else None, best_val=float(vals[bi]) if vals is not None else 0.0,
                    best_idx=bi, n_negative=nneg)

    def acq_moments(self, mu, var, kind, eta=0.0, par=0.0):
        name = ACQ_NAME[int(kind)]
        mu, var = np.asarray(mu, float).ravel(), np.asarray(var, float).ravel()
        with np.errstate(all="ignore"):
            vals = O.acq_lcb(mu, var, par) if name == "lcb" else O.ACQ[name](mu, var, eta, par)
        return np.asarray(vals, dtype=float), int((np.asarray(vals) < 0).sum()) if name == "ei" else 0

    def reduce_models(self, A, B=None):
        A = np.asarray(A, float)
        if B is None:
            return A.mean(axis=0)
        return O.mcmc_mixture_moments(A, np.asarray(B, float))

    def kernel_matrix(self, X1, X2):
        return self.kernel.get_value(np.asarray(X1, float), np.asarray(X2, float))

    def nll_grad(self, noise_var, n_terms):
        Kinv = spla.cho_solve((self.L, True), np.eye(len(self.y)))
        A = np.outer(self.alpha, self.alpha) - Kinv
        family, log_amp, axis, group, lm = self.spec
        g = np.zeros(n_terms + 2)
        Kf = self.kernel.get_value(self.X)
        g[0] = -0.5 * np.sum(A * Kf)
        Kg = self.kernel.gradient(self.X)              # [const, per-group metric params...] in flatten order
        for t in range(n_terms):
            g[1 + t] = -0.5 * np.sum(A * Kg[:, :, 1 + t])
        g[-1] = -0.5 * noise_var * np.trace(A)
        return g

    def generate_candidates(self, seed, first, count, n_uniform, lower, upper, incumbent, scale):
        return O.generate_candidates(seed, first, count, n_uniform, lower, upper, incumbent, scale)

    def maximize_random(self, seed, first, count, n_uniform, lower, upper, incumbent, scale, kind, eta=0.0, par=0.0):
        C = self.generate_candidates(seed, first, count, n_uniform, lower, upper, incumbent, scale)
        r = self.acq(C, kind, eta, par)
        return C[r["best_idx"]], r["best_val"], first + r["best_idx"]

    def comm_info(self):
        return dict(rank=0, world=1, nccl_version=0)

    def timings(self):
        return dict(fit_ms=0.0, score_ms=0.0, launches_total=0)


def install(monkeypatch):
    """Route robo_b200 through FakeHandle for the duration of a test."""
    from robo_b200 import _lib
    pool = {}

    def moments_handle(device=0):
        return pool.setdefault(device, FakeHandle(device))
    def acq_multi(handles, Xs, mode, kind=0, eta=None, par=0.0, want_argmax=False):
        if mode == 1:
            mom = [h.predict(Xs) for h in handles]
            m, v = O.mcmc_mixture_moments(np.array([a for a, _ in mom]), np.array([b for _, b in mom]))
            return dict(mean=m, var=v)
        etas = np.broadcast_to(np.zeros(1) if eta is None else np.asarray(eta, float), (len(handles),))
        rs = [h.acq(Xs, kind, float(e), par) for h, e in zip(handles, etas)]
        vals = np.mean([r["values"] for r in rs], axis=0)
        return dict(values=vals, n_negative=sum(r["n_negative"] for r in rs), best_val=float(vals.max()),
                    best_idx=int(np.argmax(vals)))
    monkeypatch.setattr(_lib, "acq_multi", acq_multi)
    monkeypatch.setattr(_lib, "Handle", FakeHandle)
    monkeypatch.setattr(_lib, "moments_handle", moments_handle)
    return FakeHandle
