"""Acquisition interface (robo/acquisition_functions/base_acquisition.py:4-69) plus the two
dispatch helpers every concrete acquisition shares on the B200 path:

* model exposes ``score`` (robo_b200 GaussianProcess): ONE fused launch sequence does
  predict -> closed form -> arg-max on the GPU;
* any other BaseModel (e.g. the reference's test/dummy_model.py): the model predicts, the
  closed form still runs on the GPU through gpk_acq_moments.  There is no CPU formula here.
"""
import numpy as np

from robo_b200 import _lib


class BaseAcquisitionFunction(object):
    kind = None       # 'ei' | 'log_ei' | 'pi' | 'lcb'

    def __init__(self, model):
        self.model = model

    def update(self, model):
        self.model = model

    def compute(self, x, derivative=False):
        raise NotImplementedError

    def __call__(self, x, **kwargs):
        return self.compute(x, **kwargs)

    def get_json_data(self):
        return {"type": __name__}

    # ---- shared device dispatch ---------------------------------------------------------
    def _values(self, X, eta, par):
        """-> (values (M,), n_negative, best_idx)"""
        X = np.asarray(X, dtype=np.float64)
        if hasattr(self.model, "score"):
            r = self.model.score(X, self.kind, eta=eta, par=par)
            return r["values"], r["n_negative"], r["best_idx"]
        m, v = self.model.predict(X)
        if eta is None and self.kind != "lcb":
            _, eta = self.model.get_incumbent()
        vals, nneg = _lib.moments_handle().acq_moments(m, v, _lib.ACQ_KIND[self.kind],
                                                       0.0 if eta is None else eta, par)
        return vals, nneg, None

    def argmax(self, X, eta=None):
        """Index of the best candidate (numpy.argmax semantics) without materialising the
        acquisition values on the host when the model supports the fused path."""
        X = np.asarray(X, dtype=np.float64)
        if hasattr(self.model, "score"):
            r = self.model.score(X, self.kind, eta=eta, par=self.par, want_values=False)
            return int(r["best_idx"])
        return int(np.argmax(self.compute(X)))
