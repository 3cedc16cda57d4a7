"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/gpk.h declares; host-side logic of the george-compatible kernels; and the
product never falls back to a CPU path (it must fail loudly without a GPU)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gpk.h")).read()
    return sorted(set(re.findall(r"\b(gpk_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from robo_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libgpk.so does not export %s" % name
    assert set(declared) == set(_lib.exported_symbols())
    assert b"sm_100a" in lib.gpk_version()


def test_library_is_sm100a_with_tma_and_dmma():
    """the built cubin must be Blackwell-native: TMA (UTMALDG) staging + DMMA tensor ops."""
    import subprocess
    from robo_b200 import _lib
    _lib.load()
    out = subprocess.run(["cuobjdump", "-sass", _lib.library_path()], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True).stdout
    if "sm_100a" not in out and "SM100" not in out.upper():
        pytest.skip("cuobjdump not available")
    assert "UTMALDG" in out and "DMMA" in out and "SYNCS" in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from robo_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Handle(0)
    from robo_b200.models.gaussian_process import GaussianProcess
    from robo_b200 import kernels as K
    m = GaussianProcess(K.Matern52Kernel(np.ones(2), ndim=2), normalize_input=False)
    with pytest.raises(RuntimeError):
        m.train(np.random.rand(5, 2), np.random.rand(5), do_optimize=False)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "robo_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_kernel_algebra_matches_george_layout():
    from robo_b200 import kernels as K
    k = 2 * K.Matern52Kernel(np.ones(3), ndim=3)          # fmin/bayesian_optimization.py:79-81
    assert len(k) == 4
    np.testing.assert_allclose(k.get_parameter_vector(), [np.log(2.0 / 3), 0, 0, 0])
    k.set_parameter_vector([0.3, -1, -2, -3])
    np.testing.assert_allclose(k[:], [0.3, -1, -2, -3])
    f = k.flatten()
    assert f["family"] == 0 and f["group"] == [0, 0, 0] and f["axis"] == [0, 1, 2]
    assert f["log_amp"] == pytest.approx(0.3) and f["log_metric"] == [-1, -2, -3]
    # fabolas.py:104-110: product of 1-D kernels, one group per axis
    k = 1
    for d in range(3):
        k *= K.Matern52Kernel(np.ones([1]) * 0.01, ndim=4, axes=d)
    f = k.flatten()
    assert len(k) == 4 and f["group"] == [0, 1, 2] and f["axis"] == [0, 1, 2]
    assert f["log_amp"] == pytest.approx(np.log(1.0 / 4))
    import copy
    k2 = copy.deepcopy(k)
    k2.set_parameter_vector([0.0, 1.0, 2.0, 3.0])
    assert k[1] == pytest.approx(np.log(0.01)) and k2[1] == 1.0
    with pytest.raises(NotImplementedError):
        (K.Matern52Kernel(1.0, ndim=1) + K.Matern52Kernel(2.0, ndim=1)).flatten()
    iso = K.ExpSquaredKernel(0.5, ndim=3)
    assert len(iso) == 1 and iso.flatten()["log_metric"] == [np.log(0.5)] * 3


def test_model_api_surface_matches_reference():
    import inspect
    from robo_b200.models.gaussian_process import GaussianProcess
    sig = inspect.signature(GaussianProcess.__init__)
    names = list(sig.parameters)[1:]
    assert names[:9] == ["kernel", "prior", "noise", "use_gradients", "normalize_output",
                         "normalize_input", "lower", "upper", "rng"]
    assert sig.parameters["noise"].default == 1e-3 and sig.parameters["normalize_input"].default is True
    for meth in ["train", "predict", "nll", "grad_nll", "optimize", "predict_variance", "sample_functions",
                 "get_incumbent", "get_noise", "update", "get_json_data"]:
        assert callable(getattr(GaussianProcess, meth))
    from robo_b200 import kernels as K
    m = GaussianProcess(K.Matern52Kernel(np.ones(2), ndim=2))
    with pytest.raises(Exception, match="Model has to be trained first!"):
        m.predict(np.zeros((3, 2)))
    with pytest.raises(AssertionError):
        m.train(np.zeros((3, 2)), np.zeros((4,)))
