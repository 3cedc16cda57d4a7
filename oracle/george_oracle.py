"""numpy/scipy restatement of the subset of *george* 0.3.x that RoBO calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

george is the third-party C++/Cython GP library that holds all arithmetic of
the reference's hot path (reference ``requirements.txt:8`` pins it to the
moving branch ``automl/george@development``; ``setup.py:7`` lists bare
``george``).  Its source is not under /root/reference and it cannot be built
here (no network, no Eigen), so its published algorithm is restated below.
Call sites in the reference that this module serves:

* ``george.kernels.Matern52Kernel(metric, ndim[, axes])``, ``scalar * kernel``,
  ``kernel *= kernel``              robo/fmin/bayesian_optimization.py:79-81,
                                    robo/fmin/fabolas.py:105-117
* ``len(kernel)``, ``kernel[:]``, ``get/set_parameter_vector``
                                    robo/models/gaussian_process.py:110,113,151,204,
                                    robo/models/gaussian_process_mcmc.py:145
* ``kernel.get_value(X1[, X2])``    test/test_models/test_gaussian_process.py:44-46
* ``kernel.gradient(X)``            robo/models/gaussian_process.py:181
* ``george.GP(kernel, mean=m)``     robo/models/gaussian_process.py:106
* ``gp.compute(X, yerr=)``          robo/models/gaussian_process.py:119,122,155,173
* ``gp.log_likelihood(y, quiet=True)``  robo/models/gaussian_process.py:159
* ``gp.predict(y, X*)``             robo/models/gaussian_process.py:280
* ``gp.sample_conditional``         robo/models/gaussian_process.py:324
* ``gp._compute_alpha``, ``gp.solver.apply_inverse``, ``gp._alpha``, ``gp._x``
                                    robo/models/gaussian_process.py:175-185

Semantics restated:

* metric = *squared* length scale; parameters are log(metric_d);
  r2 = sum_d (x_d - x'_d)^2 / metric_d over the kernel's axes.
* Matern-5/2: k = (1 + sqrt(5 r2) + 5 r2 / 3) exp(-sqrt(5 r2)).
* Matern-3/2: k = (1 + sqrt(3 r2)) exp(-sqrt(3 r2)).
* ExpSquared: k = exp(-r2 / 2).
* ``c * kernel`` = Product(ConstantKernel(log(c / ndim)), kernel)   [george 0.3
  ``Kernel.__mul__/__rmul__``; affects only the *initial* amplitude].
* ``GP.compute``: K = k(X, X) + diag(yerr^2 + TINY), TINY = 1.25e-12, factor =
  scipy.linalg.cholesky(K, lower=False) (raises numpy.linalg.LinAlgError when
  not positive definite); log|K| = 2 sum log diag.
* ``log_likelihood``: -1/2 r^T K^-1 r - 1/2 log|K| - N/2 log 2 pi, r = y - mean.
* ``predict``: mu = K* alpha + mean, cov = K** - K* K^-1 K*^T (no noise on K**).

Parity status: UNPINNED at the george boundary in the strict sense (no numeric
george output exists in the reference); pinned to independent implementations
in tests/test_oracle_george.py (sklearn Matern/RBF, mpmath 50-digit GP).
"""
import copy

import numpy as np
import scipy.linalg as spla

__version__ = "0.3.1-oracle"

TINY = 1.25e-12


# --------------------------------------------------------------------------- #
# kernels
# --------------------------------------------------------------------------- #
class Kernel(object):
    is_kernel = True
    kernel_type = -1

    def __init__(self, ndim=1, axes=None):
        self.ndim = int(ndim)
        if axes is None:
            self.axes = np.arange(self.ndim)
        else:
            self.axes = np.atleast_1d(np.asarray(axes, dtype=int))
            if np.any(self.axes >= self.ndim) or np.any(self.axes < 0):
                raise ValueError("invalid axis for {0} dims".format(self.ndim))

    # -- parameter protocol (george.modeling.Model) -------------------------
    def __len__(self):
        return len(self.get_parameter_vector())

    def get_parameter_vector(self, include_frozen=False):
        raise NotImplementedError

    def set_parameter_vector(self, vector, include_frozen=False):
        raise NotImplementedError

    def get_parameter_names(self, include_frozen=False):
        raise NotImplementedError

    def __getitem__(self, idx):
        return self.get_parameter_vector()[idx]

    def __setitem__(self, idx, value):
        v = self.get_parameter_vector()
        v[idx] = value
        self.set_parameter_vector(v)

    @property
    def vector(self):          # george-0.2 spelling read by mtbo_gp.py:53
        return self.get_parameter_vector()

    # -- algebra -------------------------------------------------------------
    def _as_kernel(self, b):
        if hasattr(b, "is_kernel"):
            return b
        return ConstantKernel(log_constant=np.log(float(b) / self.ndim),
                              ndim=self.ndim)

    def __add__(self, b):
        if not hasattr(b, "is_kernel"):
            return Sum(self._as_kernel(b), self)
        return Sum(self, b)

    def __radd__(self, b):
        return self.__add__(b)

    def __mul__(self, b):
        if not hasattr(b, "is_kernel"):
            return Product(self._as_kernel(b), self)
        return Product(self, b)

    def __rmul__(self, b):
        return self.__mul__(b)

    # -- evaluation ----------------------------------------------------------
    def _parse(self, x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        if x.ndim == 1:
            x = np.atleast_2d(x).T
        if x.ndim != 2 or x.shape[1] != self.ndim:
            raise ValueError("Dimension mismatch")
        return x

    def get_value(self, x1, x2=None, diag=False):
        x1 = self._parse(x1)
        if x2 is None:
            if diag:
                return np.array([self._value(x1[i:i + 1], x1[i:i + 1])[0, 0]
                                 for i in range(len(x1))])
            return self._value(x1, x1)
        x2 = self._parse(x2)
        return self._value(x1, x2)

    def gradient(self, x1, x2=None):
        """d k / d theta_p, shape (n1, n2, len(self))."""
        x1 = self._parse(x1)
        x2 = x1 if x2 is None else self._parse(x2)
        return self._gradient(x1, x2)

    def _value(self, x1, x2):
        raise NotImplementedError

    def _gradient(self, x1, x2):
        raise NotImplementedError


class ConstantKernel(Kernel):
    kernel_type = 0

    def __init__(self, log_constant, ndim=1, axes=None):
        super(ConstantKernel, self).__init__(ndim, axes)
        self.log_constant = float(log_constant)

    def get_parameter_vector(self, include_frozen=False):
        return np.array([self.log_constant])

    def set_parameter_vector(self, vector, include_frozen=False):
        vector = np.atleast_1d(vector)
        assert len(vector) == 1
        self.log_constant = float(vector[0])

    def get_parameter_names(self, include_frozen=False):
        return ("log_constant",)

    def _value(self, x1, x2):
        return np.full((len(x1), len(x2)), np.exp(self.log_constant))

    def _gradient(self, x1, x2):
        return np.full((len(x1), len(x2), 1), np.exp(self.log_constant))


class _RadialKernel(Kernel):
    """Stationary kernel k = f(r2), r2 = sum_d (x_d - x'_d)^2 / metric_d."""

    def __init__(self, metric, ndim=1, axes=None):
        super(_RadialKernel, self).__init__(ndim, axes)
        metric = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if metric.ndim != 1:
            raise NotImplementedError("general (matrix) metrics are not used by RoBO")
        if len(metric) == 1:
            self.isotropic = True
            self.log_metric = np.log(metric).copy()
        else:
            if len(metric) != len(self.axes):
                raise ValueError("Dimension mismatch")
            self.isotropic = False
            self.log_metric = np.log(metric).copy()

    def get_parameter_vector(self, include_frozen=False):
        return self.log_metric.copy()

    def set_parameter_vector(self, vector, include_frozen=False):
        vector = np.atleast_1d(np.asarray(vector, dtype=np.float64))
        assert len(vector) == len(self.log_metric)
        self.log_metric = vector.copy()

    def get_parameter_names(self, include_frozen=False):
        if self.isotropic:
            return ("metric:log_M_0_0",)
        return tuple("metric:log_M_{0}_{0}".format(i) for i in range(len(self.log_metric)))

    def _axis_metric(self):
        m = np.exp(self.log_metric)
        if self.isotropic:
            m = np.full(len(self.axes), m[0])
        return m

    def _sqdiff(self, x1, x2):
        """(n1, n2, n_axes) array of (x_d - x'_d)^2 / metric_d."""
        m = self._axis_metric()
        d = x1[:, None, self.axes] - x2[None, :, self.axes]
        return d * d / m

    def _r2(self, x1, x2):
        m = self._axis_metric()
        r2 = np.zeros((len(x1), len(x2)))
        for a, md in zip(self.axes, m):     # per-dimension loop keeps memory O(n1 n2)
            d = x1[:, a][:, None] - x2[:, a][None, :]
            r2 += d * d / md
        return r2

    def _f(self, r2):
        raise NotImplementedError

    def _dfdr2(self, r2):
        raise NotImplementedError

    def _value(self, x1, x2):
        return self._f(self._r2(x1, x2))

    def _gradient(self, x1, x2):
        # d r2 / d log metric_d = -(x_d - x'_d)^2 / metric_d
        sq = self._sqdiff(x1, x2)
        dfdr2 = self._dfdr2(sq.sum(axis=2))
        g = -dfdr2[:, :, None] * sq
        if self.isotropic:
            g = g.sum(axis=2, keepdims=True)
        return g


class Matern52Kernel(_RadialKernel):
    kernel_type = 6

    def _f(self, r2):
        r = np.sqrt(5.0 * r2)
        return (1.0 + r + 5.0 * r2 / 3.0) * np.exp(-r)

    def _dfdr2(self, r2):
        r = np.sqrt(5.0 * r2)
        return -(5.0 / 6.0) * (1.0 + r) * np.exp(-r)


class Matern32Kernel(_RadialKernel):
    kernel_type = 5

    def _f(self, r2):
        r = np.sqrt(3.0 * r2)
        return (1.0 + r) * np.exp(-r)

    def _dfdr2(self, r2):
        r = np.sqrt(3.0 * r2)
        return -1.5 * np.exp(-r)


class ExpSquaredKernel(_RadialKernel):
    kernel_type = 3

    def _f(self, r2):
        return np.exp(-0.5 * r2)

    def _dfdr2(self, r2):
        return -0.5 * np.exp(-0.5 * r2)


class _Operator(Kernel):
    def __init__(self, k1, k2):
        if k1.ndim != k2.ndim:
            raise ValueError("Dimension mismatch")
        self.k1 = k1
        self.k2 = k2
        self.ndim = k1.ndim
        self.axes = np.arange(self.ndim)

    def get_parameter_vector(self, include_frozen=False):
        return np.concatenate((self.k1.get_parameter_vector(),
                               self.k2.get_parameter_vector()))

    def set_parameter_vector(self, vector, include_frozen=False):
        vector = np.atleast_1d(np.asarray(vector, dtype=np.float64))
        n1 = len(self.k1)
        assert len(vector) == n1 + len(self.k2)
        self.k1.set_parameter_vector(vector[:n1])
        self.k2.set_parameter_vector(vector[n1:])

    def get_parameter_names(self, include_frozen=False):
        return tuple("k1:" + n for n in self.k1.get_parameter_names()) + \
            tuple("k2:" + n for n in self.k2.get_parameter_names())


class Sum(_Operator):
    def _value(self, x1, x2):
        return self.k1._value(x1, x2) + self.k2._value(x1, x2)

    def _gradient(self, x1, x2):
        return np.concatenate((self.k1._gradient(x1, x2),
                               self.k2._gradient(x1, x2)), axis=2)


class Product(_Operator):
    def _value(self, x1, x2):
        return self.k1._value(x1, x2) * self.k2._value(x1, x2)

    def _gradient(self, x1, x2):
        v1 = self.k1._value(x1, x2)
        v2 = self.k2._value(x1, x2)
        return np.concatenate((self.k1._gradient(x1, x2) * v2[:, :, None],
                               self.k2._gradient(x1, x2) * v1[:, :, None]), axis=2)


class _KernelsNamespace(object):
    """so that ``george.kernels.Matern52Kernel`` resolves on this module."""
    Kernel = Kernel
    ConstantKernel = ConstantKernel
    Matern52Kernel = Matern52Kernel
    Matern32Kernel = Matern32Kernel
    ExpSquaredKernel = ExpSquaredKernel
    Sum = Sum
    Product = Product


kernels = _KernelsNamespace()


# --------------------------------------------------------------------------- #
# solver + GP
# --------------------------------------------------------------------------- #
class BasicSolver(object):
    """george.solvers.BasicSolver: dense Cholesky through scipy (LAPACK potrf/potrs)."""

    def __init__(self, kernel):
        self.kernel = kernel
        self._computed = False
        self._log_det = None

    @property
    def computed(self):
        return self._computed

    @property
    def log_determinant(self):
        return self._log_det

    def compute(self, x, yerr):
        K = self.kernel.get_value(x)
        K[np.diag_indices_from(K)] += yerr ** 2
        self._factor = (spla.cholesky(K, overwrite_a=True, lower=False), False)
        self._log_det = 2.0 * np.sum(np.log(np.diag(self._factor[0])))
        self._computed = True

    def apply_inverse(self, y, in_place=False):
        return spla.cho_solve(self._factor, y, overwrite_b=in_place)

    def dot_solve(self, y):
        return np.dot(y.T, spla.cho_solve(self._factor, y))

    def get_inverse(self):
        return self.apply_inverse(np.eye(self._factor[0].shape[0]), in_place=True)


class GP(object):
    def __init__(self, kernel, fit_kernel=True, mean=None, fit_mean=None,
                 white_noise=None, fit_white_noise=None, solver=None, **kwargs):
        self.kernel = kernel
        self.mean = 0.0 if mean is None else mean
        self.white_noise = np.log(TINY) if white_noise is None else white_noise
        self.solver_type = BasicSolver if solver is None else solver
        self.solver = None
        self._computed = False
        self._alpha = None
        self._x = None
        self._y = None

    @property
    def computed(self):
        return self._computed and self.solver is not None and self.solver.computed

    def _call_mean(self, x):
        if callable(self.mean):
            return np.asarray(self.mean(x), dtype=np.float64)
        return float(self.mean) + np.zeros(len(x))

    def parse_samples(self, t):
        return self.kernel._parse(t)

    def _check_dimensions(self, y):
        y = np.atleast_1d(np.asarray(y, dtype=np.float64))
        if self._x is None or len(self._x) != y.shape[0]:
            raise ValueError("Dimension mismatch")
        return y

    def compute(self, x, yerr=0.0, **kwargs):
        self._x = self.parse_samples(x)
        self._x = np.ascontiguousarray(self._x, dtype=np.float64)
        try:
            self._yerr2 = float(yerr) ** 2 * np.ones(len(self._x))
        except TypeError:
            self._yerr2 = np.asarray(yerr, dtype=np.float64) ** 2
        self.solver = self.solver_type(self.kernel, **kwargs)
        yerr_tot = np.sqrt(self._yerr2 + np.exp(self.white_noise))
        self.solver.compute(self._x, yerr_tot)
        self._const = -0.5 * (len(self._x) * np.log(2 * np.pi) + self.solver.log_determinant)
        self._computed = True
        self._alpha = None

    def _compute_alpha(self, y, cache=True):
        r = np.ascontiguousarray(self._check_dimensions(y) - self._call_mean(self._x),
                                 dtype=np.float64)
        alpha = self.solver.apply_inverse(r, in_place=True).flatten()
        if cache:
            self._alpha = alpha
        return alpha

    def log_likelihood(self, y, quiet=False):
        try:
            r = self._check_dimensions(y) - self._call_mean(self._x)
            ll = self._const - 0.5 * np.dot(r, self.solver.apply_inverse(r))
        except (ValueError, np.linalg.LinAlgError):
            if quiet:
                return -np.inf
            raise
        return ll if np.isfinite(ll) else -np.inf

    lnlikelihood = log_likelihood

    def predict(self, y, t, return_cov=True, return_var=False):
        alpha = self._compute_alpha(y)
        xs = self.parse_samples(t)
        Kxs = self.kernel.get_value(xs, self._x)
        mu = np.dot(Kxs, alpha) + self._call_mean(xs)
        if not (return_var or return_cov):
            return mu
        KinvKxs = self.solver.apply_inverse(Kxs.T)
        if return_var:
            var = self.kernel.get_value(xs, diag=True)
            var -= np.sum(Kxs.T * KinvKxs, axis=0)
            return mu, var
        cov = self.kernel.get_value(xs)
        cov -= np.dot(Kxs, KinvKxs)
        return mu, cov

    def sample_conditional(self, y, t, size=1):
        mu, cov = self.predict(y, t)
        return np.random.multivariate_normal(mu, cov, size=size) if size > 1 \
            else np.random.multivariate_normal(mu, cov)

    def sample(self, t=None, size=1):
        x = self._x if t is None else self.parse_samples(t)
        cov = self.kernel.get_value(x)
        cov[np.diag_indices_from(cov)] += TINY
        return np.random.multivariate_normal(self._call_mean(x), cov, size=size)


def install_as_george():
    """Register this module as ``george`` (+ ``george.kernels``) in sys.modules so
    the unmodified reference (``import george``) runs on top of the restatement.
    Used by oracle/make_golden.py only."""
    import sys
    mod = sys.modules[__name__]
    sys.modules["george"] = mod
    sys.modules["george.kernels"] = kernels
    return mod
