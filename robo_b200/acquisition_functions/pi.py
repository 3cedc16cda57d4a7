"""Probability of Improvement (robo/acquisition_functions/pi.py)."""
from robo_b200.acquisition_functions.base_acquisition import BaseAcquisitionFunction


class PI(BaseAcquisitionFunction):
    kind = "pi"

    def __init__(self, model, par=0.0):
        super(PI, self).__init__(model)
        self.par = par

    def compute(self, X_test, derivative=False, **kwargs):
        """Phi((eta - m - par) / s), eta always the incumbent (pi.py:58-63)."""
        if derivative:
            if not hasattr(self.model, "score_with_gradient"):
                raise NotImplementedError("derivative=True needs a model with predictive gradients")
            import numpy as np
            return self.model.score_with_gradient(np.asarray(X_test, dtype=np.float64), "pi", par=self.par)
        return self._values(X_test, None, self.par)[0]
