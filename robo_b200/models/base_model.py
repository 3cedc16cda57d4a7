"""Model interface of the reference (robo/models/base_model.py:5-106), kept so solvers,
maximizers and acquisition functions written against RoBO's BaseModel drop in unchanged."""
import numpy as np


class BaseModel(object):

    def __init__(self):
        self.X = None
        self.y = None

    def train(self, X, y):
        raise NotImplementedError

    def predict(self, X_test):
        raise NotImplementedError

    def update(self, X, y):
        """Append the new observations and retrain (base_model.py:30-45)."""
        X = np.append(self.X, X, axis=0)
        y = np.append(self.y, y, axis=0)
        self.train(X, y)

    def _check_shapes_train(func):
        def func_wrapper(self, X, y, *args, **kwargs):
            assert X.shape[0] == y.shape[0]
            assert len(X.shape) == 2
            assert len(y.shape) == 1
            return func(self, X, y, *args, **kwargs)
        return func_wrapper

    def _check_shapes_predict(func):
        def func_wrapper(self, X, *args, **kwargs):
            assert len(X.shape) == 2
            return func(self, X, *args, **kwargs)
        return func_wrapper

    def get_json_data(self):
        return {'X': self.X if self.X is None else self.X.tolist(),
                'y': self.y if self.y is None else self.y.tolist(),
                'hyperparameters': ""}

    def get_incumbent(self):
        best_idx = np.argmin(self.y)
        return self.X[best_idx], self.y[best_idx]
