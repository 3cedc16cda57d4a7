"""george-compatible kernel objects for the B200 path (host side: parameters only).

RoBO's callers build the covariance with george's kernel algebra
(robo/fmin/bayesian_optimization.py:79-81, robo/fmin/fabolas.py:104-110) and then only use
    len(kernel), kernel.get_parameter_vector(), kernel.set_parameter_vector(v), kernel[:],
    kernel.get_value(X1[, X2]), copy.deepcopy(kernel)
(gaussian_process.py:110,113,151,204; gaussian_process_mcmc.py:145,153;
test/test_models/test_gaussian_process.py:44-46).  These classes keep that surface and
*flatten* to the C ABI's kernel description (include/gpk.h: gpk_set_kernel):

    k(x, x') = exp(log_amp) * prod_g f( sum_{t in g} (x[axis_t] - x'[axis_t])^2 / exp(log_metric_t) )

Every value is computed on the GPU (gpk_kernel_matrix); nothing here evaluates a kernel on
the CPU.  Supported algebra: products of ConstantKernel and radial kernels of ONE family
(Matern-5/2, Matern-3/2 or ExpSquared).  Sums are not representable on the device path and
raise NotImplementedError when flattened.
"""
import numpy as np

from . import _lib

__all__ = ["Kernel", "ConstantKernel", "Matern52Kernel", "Matern32Kernel", "ExpSquaredKernel",
           "Product", "Sum"]


class Kernel(object):
    is_kernel = True

    def __init__(self, ndim=1, axes=None):
        self.ndim = int(ndim)
        if axes is None:
            self.axes = np.arange(self.ndim)
        else:
            self.axes = np.atleast_1d(np.asarray(axes, dtype=int))
            if np.any(self.axes < 0) or np.any(self.axes >= self.ndim):
                raise ValueError("invalid axis for {0} dims".format(self.ndim))

    # ---- parameter protocol ----------------------------------------------------
    def __len__(self):
        return len(self.get_parameter_vector())

    def __getitem__(self, idx):
        return self.get_parameter_vector()[idx]

    def __setitem__(self, idx, value):
        v = self.get_parameter_vector()
        v[idx] = value
        self.set_parameter_vector(v)

    @property
    def vector(self):
        return self.get_parameter_vector()

    @property
    def pars(self):
        return np.exp(self.get_parameter_vector())

    # ---- algebra ------------------------------------------------------------------
    def _coerce(self, b):
        if hasattr(b, "is_kernel"):
            return b
        # george 0.3: a scalar c becomes ConstantKernel(log(c / ndim))
        return ConstantKernel(log_constant=np.log(float(b) / self.ndim), ndim=self.ndim)

    def __mul__(self, b):
        if not hasattr(b, "is_kernel"):
            return Product(self._coerce(b), self)
        return Product(self, b)

    __rmul__ = __mul__

    def __add__(self, b):
        if not hasattr(b, "is_kernel"):
            return Sum(self._coerce(b), self)
        return Sum(self, b)

    __radd__ = __add__

    # ---- device description --------------------------------------------------------
    def _collect(self, acc):
        raise NotImplementedError

    def flatten(self):
        """-> dict(family, log_amp, axis, group, log_metric, slots) for gpk_set_kernel.
        slots[i] = ('amp', None) or ('metric', [term indices]) for parameter i, used to map
        gradients back onto the george parameter vector."""
        acc = dict(family=None, log_amp=0.0, axis=[], group=[], log_metric=[], slots=[], ngroups=0)
        self._collect(acc)
        if acc["family"] is None:
            raise NotImplementedError("the device path needs at least one radial kernel factor")
        return acc

    # ---- values (GPU) ------------------------------------------------------------------
    def _parse(self, x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        if x.ndim == 1:
            x = np.atleast_2d(x).T
        if x.ndim != 2 or x.shape[1] != self.ndim:
            raise ValueError("Dimension mismatch")
        return x

    def get_value(self, x1, x2=None, device=0):
        x1 = self._parse(x1)
        x2 = x1 if x2 is None else self._parse(x2)
        f = self.flatten()
        h = _lib.moments_handle(device)
        h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
        return h.kernel_matrix(x1, x2)


class ConstantKernel(Kernel):
    def __init__(self, log_constant, ndim=1, axes=None):
        super(ConstantKernel, self).__init__(ndim, axes)
        self.log_constant = float(log_constant)

    def get_parameter_vector(self, include_frozen=False):
        return np.array([self.log_constant])

    def set_parameter_vector(self, vector, include_frozen=False):
        vector = np.atleast_1d(vector)
        if len(vector) != 1:
            raise ValueError("dimension mismatch")
        self.log_constant = float(vector[0])

    def get_parameter_names(self, include_frozen=False):
        return ("log_constant",)

    def _collect(self, acc):
        acc["log_amp"] += self.log_constant
        acc["slots"].append(("amp", None))


class _Radial(Kernel):
    family = None

    def __init__(self, metric, ndim=1, axes=None):
        super(_Radial, self).__init__(ndim, axes)
        metric = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if metric.ndim != 1:
            raise NotImplementedError("general (matrix) metrics are not supported")
        if len(metric) != 1 and len(metric) != len(self.axes):
            raise ValueError("Dimension mismatch")
        self.isotropic = len(metric) == 1
        self.log_metric = np.log(metric)

    def get_parameter_vector(self, include_frozen=False):
        return self.log_metric.copy()

    def set_parameter_vector(self, vector, include_frozen=False):
        vector = np.atleast_1d(np.asarray(vector, dtype=np.float64))
        if len(vector) != len(self.log_metric):
            raise ValueError("dimension mismatch")
        self.log_metric = vector.copy()

    def get_parameter_names(self, include_frozen=False):
        return tuple("metric:log_M_{0}_{0}".format(i) for i in range(len(self.log_metric)))

    def _collect(self, acc):
        if acc["family"] is not None and acc["family"] != self.family:
            raise NotImplementedError("products of different radial families are not supported on the device")
        acc["family"] = self.family
        g = acc["ngroups"]
        acc["ngroups"] += 1
        first = len(acc["axis"])
        for i, a in enumerate(self.axes):
            acc["axis"].append(int(a))
            acc["group"].append(g)
            acc["log_metric"].append(float(self.log_metric[0 if self.isotropic else i]))
        terms = list(range(first, len(acc["axis"])))
        if self.isotropic:
            acc["slots"].append(("metric", terms))
        else:
            for t in terms:
                acc["slots"].append(("metric", [t]))


class Matern52Kernel(_Radial):
    family = _lib.MATERN52


class Matern32Kernel(_Radial):
    family = _lib.MATERN32


class ExpSquaredKernel(_Radial):
    family = _lib.EXPSQUARED


class _Operator(Kernel):
    def __init__(self, k1, k2):
        if k1.ndim != k2.ndim:
            raise ValueError("Dimension mismatch")
        self.k1, self.k2 = k1, k2
        self.ndim = k1.ndim
        self.axes = np.arange(self.ndim)

    def get_parameter_vector(self, include_frozen=False):
        return np.concatenate((self.k1.get_parameter_vector(), self.k2.get_parameter_vector()))

    def set_parameter_vector(self, vector, include_frozen=False):
        vector = np.atleast_1d(np.asarray(vector, dtype=np.float64))
        n1 = len(self.k1)
        if len(vector) != n1 + len(self.k2):
            raise ValueError("dimension mismatch")
        self.k1.set_parameter_vector(vector[:n1])
        self.k2.set_parameter_vector(vector[n1:])

    def get_parameter_names(self, include_frozen=False):
        return tuple("k1:" + n for n in self.k1.get_parameter_names()) + \
            tuple("k2:" + n for n in self.k2.get_parameter_names())


class Product(_Operator):
    def _collect(self, acc):
        self.k1._collect(acc)
        self.k2._collect(acc)


class Sum(_Operator):
    def _collect(self, acc):
        raise NotImplementedError("sums of kernels are not supported on the device path")
