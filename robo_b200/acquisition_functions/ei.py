"""Expected Improvement (robo/acquisition_functions/ei.py)."""
import logging

import numpy as np

from robo_b200.acquisition_functions.base_acquisition import BaseAcquisitionFunction

logger = logging.getLogger(__name__)


class EI(BaseAcquisitionFunction):
    kind = "ei"

    def __init__(self, model, par=0.0):
        super(EI, self).__init__(model)
        self.par = par

    def compute(self, X, derivative=False, eta=None, **kwargs):
        """EI(X) = s (z Phi(z) + phi(z)), z = (eta - m - par) / s   (ei.py:65-78).
        Keeps the reference's quirks: a zero predictive std anywhere zeroes the whole batch
        (ei.py:72-74); any negative value raises ValueError (ei.py:86-88)."""
        if derivative:
            # ei.py:80-85: df = -dm Phi(z) + ds phi(z); the reference needs model.predictive_gradients,
            # which none of its models implements — the robo_b200 GaussianProcess does (on the device)
            if not hasattr(self.model, "score_with_gradient"):
                raise NotImplementedError("derivative=True needs a model with predictive gradients")
            return self.model.score_with_gradient(np.asarray(X, dtype=np.float64), "ei", eta=eta, par=self.par)
        if not hasattr(self.model, "score"):
            m, v = self.model.predict(X)
            if (np.sqrt(v) == 0).any():
                return np.array([[0]])
        f, n_negative, _ = self._values(X, eta, self.par)
        if n_negative > 0:
            logger.error("Expected Improvement is smaller than 0!")
            raise ValueError
        return f
