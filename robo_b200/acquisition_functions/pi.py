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
            raise NotImplementedError("derivative=True needs model.predictive_gradients")
        return self._values(X_test, None, self.par)[0]
