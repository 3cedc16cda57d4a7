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
        acquisition_values = np.zeros([n, X_test.shape[0]])
        for i in range(n):
            acquisition_values[i] = self.estimators[i].compute(X_test, derivative=derivative)
        return _lib.moments_handle().reduce_models(acquisition_values)
