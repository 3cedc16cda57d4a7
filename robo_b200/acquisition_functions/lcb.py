"""Lower Confidence Bound (robo/acquisition_functions/lcb.py)."""
from robo_b200.acquisition_functions.base_acquisition import BaseAcquisitionFunction


class LCB(BaseAcquisitionFunction):
    kind = "lcb"

    def __init__(self, model, par=1.0):
        self.par = par
        super(LCB, self).__init__(model)

    def compute(self, X, derivative=False, **kwargs):
        """-(m - par sqrt(v)) (lcb.py:62-65); RoBO maximises, so the bound is negated."""
        if derivative:
            if not hasattr(self.model, "score_with_gradient"):
                raise NotImplementedError("derivative=True needs a model with predictive gradients")
            import numpy as np
            return self.model.score_with_gradient(np.asarray(X, dtype=np.float64), "lcb", par=self.par)
        return self._values(X, None, self.par)[0]
