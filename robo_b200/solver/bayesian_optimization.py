"""Minimal BO loop with the control flow of robo/solver/bayesian_optimization.py:85-249
(initial design -> [train -> acquisition.update -> maximize -> evaluate]*), so that
BASELINE.json configs[0] (fmin on Branin) can run where the reference tree is absent.  The
reference's own solver drives the robo_b200 model/acquisition/maximizer objects unchanged as
well; this file is orchestration only and contains no arithmetic of the path."""
import logging
import time

import numpy as np

from robo_b200.initial_design import init_random_uniform

logger = logging.getLogger(__name__)


class BayesianOptimization(object):

    def __init__(self, objective_func, lower, upper, acquisition_func, model, maximize_func,
                 initial_design=init_random_uniform, initial_points=3, output_path=None,
                 train_interval=1, n_restarts=1, rng=None):
        if output_path is not None:
            # the reference writes one JSON file per iteration (solver/bayesian_optimization.py:143-144, :251-261); this
            # minimal loop does not: refuse loudly instead of silently dropping the logs (use the reference's own
            # solver with the robo_b200 objects when they are wanted)
            raise NotImplementedError("output_path is not supported by robo_b200.solver; run robo.solver.bayesian_optimization."
                                      "BayesianOptimization with the robo_b200 model / acquisition / maximizer objects")
        self.output_path = output_path
        self.n_restarts = n_restarts
        self.rng = np.random.RandomState(np.random.randint(100000)) if rng is None else rng
        self.model = model
        self.acquisition_func = acquisition_func
        self.maximize_func = maximize_func
        self.lower, self.upper = lower, upper
        self.objective_func = objective_func
        self.initial_design = initial_design
        self.init_points = initial_points
        self.train_interval = train_interval
        self.X = None
        self.y = None
        self.incumbents, self.incumbents_values = [], []
        self.time_func_evals, self.time_overhead, self.runtime = [], [], []
        self.time_train, self.time_maximize = [], []

    def _record(self):
        best = np.argmin(self.y)
        self.incumbents.append(np.asarray(self.X[best]).tolist())
        self.incumbents_values.append(self.y[best])
        self.runtime.append(time.time() - self.start_time)

    def run(self, num_iterations=10, X=None, y=None):
        self.start_time = time.time()
        if X is None and y is None:
            Xl, yl = [], []
            t0 = time.time()
            init = self.initial_design(self.lower, self.upper, self.init_points, rng=self.rng)
            t_over = (time.time() - t0) / self.init_points
            for x in init:
                t0 = time.time()
                yl.append(self.objective_func(x))
                Xl.append(x)
                self.time_func_evals.append(time.time() - t0)
                self.time_overhead.append(t_over)
                self.X, self.y = np.array(Xl), np.array(yl)
                self._record()
        else:
            self.X, self.y = X, y
        for it in range(self.init_points, num_iterations):
            t0 = time.time()
            new_x = self.choose_next(self.X, self.y, it % self.train_interval == 0)
            self.time_overhead.append(time.time() - t0)
            t0 = time.time()
            new_y = self.objective_func(new_x)
            self.time_func_evals.append(time.time() - t0)
            self.X = np.append(self.X, new_x[None, :], axis=0)
            self.y = np.append(self.y, new_y)
            self._record()
        return self.incumbents[-1], self.incumbents_values[-1]

    def choose_next(self, X=None, y=None, do_optimize=True):
        if X is None and y is None:
            return self.initial_design(self.lower, self.upper, 1, rng=self.rng)[0, :]
        if X.shape[0] == 1:
            return self.initial_design(self.lower, self.upper, 1, rng=self.rng)[0, :]
        t0 = time.time()
        self.model.train(X, y, do_optimize=do_optimize)
        self.time_train.append(time.time() - t0)
        self.acquisition_func.update(self.model)
        t0 = time.time()
        x = self.maximize_func.maximize()
        self.time_maximize.append(time.time() - t0)
        return x
