"""MarginalizationGPMCMC (robo/acquisition_functions/marginalization.py): averages an acquisition
function over the hyper-parameter samples of a GaussianProcessMCMC model.  One estimator
(deep copy of the acquisition function) per sub-model exactly like the reference (:33-47,
:64-78); every estimator's compute() is the fused GPU path of its sub-model, and the mean over
models (:121) is reduced on the GPU as well."""
from copy import deepcopy

import numpy as np

from robo_b200 import _lib
from robo_b200.acquisition_functions.base_acquisition import BaseAcquisitionFunction


class MarginalizationGPMCMC(BaseAcquisitionFunction):

    def __init__(self, acquisition_func):
        self.acquisition_func = acquisition_func
        self.model = acquisition_func.model
        self.cost_model = acquisition_func.cost_model if hasattr(acquisition_func, "cost_model") else None
        self.estimators = []
        self._make_estimators()

    def _make_estimators(self):
        for i in range(len(self.model.models)):
            estimator = deepcopy(self.acquisition_func)
            estimator.model = self.model.models[i]
            if self.cost_model is not None and len(self.cost_model.models) > 0:
                estimator.cost_model = self.cost_model.models[i]
            self.estimators.append(estimator)

    def update(self, model, cost_model=None, **kwargs):
        if len(self.estimators) == 0:
            self._make_estimators()
        self.model = model
        if cost_model is not None:
            self.cost_model = cost_model
        if len(self.estimators) != len(self.model.models):
            self.estimators = []
            self._make_estimators()
        for i in range(len(self.model.models)):
            if cost_model is not None:
                self.estimators[i].update(self.model.models[i], self.cost_model.models[i], **kwargs)
            else:
                self.estimators[i].update(self.model.models[i], **kwargs)

    def compute(self, X_test, derivative=False):
        n = len(self.model.models)
        fused = self._fused_spec() if not derivative else None
        if fused is not None:
            # marginalization.py:115-121 as ONE call: the batch goes to the device once, the n sub-models score it
            # concurrently and the mean over models is taken there (gpk_acq_multi mode 0); M doubles come back
            kind, etas, par, handles = fused
            X_dev = self.model.models[0].device_inputs(np.asarray(X_test, dtype=np.float64))
            r = _lib.acq_multi(handles, X_dev, 0, _lib.ACQ_KIND[kind], etas, par)
            if kind == "ei" and r["n_negative"] > 0:
                raise ValueError("Expected Improvement is smaller than 0!")      # ei.py:86-88
            return r["values"]
        acquisition_values = np.zeros([n, X_test.shape[0]])
        for i in range(n):
            acquisition_values[i] = self.estimators[i].compute(X_test, derivative=derivative)
        return _lib.moments_handle().reduce_models(acquisition_values)

    def argmax(self, X_test):
        """numpy.argmax of compute(X_test) taken on the device when the fused path applies."""
        fused = self._fused_spec()
        if fused is None:
            return int(np.argmax(self.compute(X_test)))
        kind, etas, par, handles = fused
        X_dev = self.model.models[0].device_inputs(np.asarray(X_test, dtype=np.float64))
        r = _lib.acq_multi(handles, X_dev, 0, _lib.ACQ_KIND[kind], etas, par, want_argmax=True)
        if kind == "ei" and r["n_negative"] > 0:
            raise ValueError("Expected Improvement is smaller than 0!")
        return int(r["best_idx"])

    def _fused_spec(self):
        """(kind, eta per model, par, handles) when every estimator is a closed-form acquisition on a device GP."""
        if self.cost_model is not None or len(self.estimators) == 0 or not hasattr(self.model, "sub_model_handles"):
            return None
        kinds = set(getattr(e, "kind", None) for e in self.estimators)
        pars = set(float(getattr(e, "par", 0.0)) for e in self.estimators)
        if len(kinds) != 1 or len(pars) != 1 or list(kinds)[0] not in ("ei", "log_ei", "pi", "lcb"):
            return None
        if any(e.model is not m or not hasattr(m, "device_inputs") for e, m in zip(self.estimators, self.model.models)):
            return None
        handles = self.model.sub_model_handles()
        if handles is None or len(handles) != len(self.estimators):
            return None
        kind = list(kinds)[0]
        etas = [0.0 if kind == "lcb" else float(e.model.get_incumbent()[1]) for e in self.estimators]
        return kind, etas, list(pars)[0], handles
