"""Log Expected Improvement (robo/acquisition_functions/log_ei.py)."""
import logging

from robo_b200.acquisition_functions.base_acquisition import BaseAcquisitionFunction

logger = logging.getLogger(__name__)


class LogEI(BaseAcquisitionFunction):
    kind = "log_ei"

    def __init__(self, model, par=0.0, **kwargs):
        super(LogEI, self).__init__(model)
        self.par = par

    def compute(self, X, derivative=False, eta=None, **kwargs):
        """All three guarded branches of log_ei.py:79-120, evaluated per candidate on the GPU."""
        if derivative:
            logger.error("LogEI does not support derivative calculation until now")
            return
        return self._values(X, eta, self.par)[0]
