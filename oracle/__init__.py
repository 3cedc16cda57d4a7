"""CPU oracle for the GP-posterior + acquisition hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  ``robo_b200`` never imports this package.

Contents
--------
george_oracle   numpy/scipy restatement of the subset of george 0.3 that RoBO
                calls (kernels, GP.compute / log_likelihood / predict).  george
                itself is an un-vendored third-party dependency of the reference
                (``requirements.txt:8``: git+https://github.com/automl/george.git
                @development, no version pin) and cannot be installed here.
robo_oracle     numpy/scipy restatement of robo/models/gaussian_process.py and
                robo/acquisition_functions/{ei,log_ei,pi,lcb}.py.
make_golden.py  runs the *real* reference classes from /root/reference on top of
                george_oracle (registered as ``sys.modules['george']``) and writes
                tests/golden/*.npz; it also asserts robo_oracle == reference run.

Parity status: the RoBO layer is pinned (the reference's own Python code is
executed to produce the golden vectors).  The george layer is pinned only to
the closed-form known answer of test/test_models/test_gaussian_process.py:44-49
and to independent implementations (sklearn Matern/RBF, mpmath 50-digit
arithmetic); no numeric output of george itself exists anywhere in the
reference, so at the george boundary parity is "unpinned" in the strict sense.
"""
