import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "raw_abi: drives libmetrics_b200.so directly (deselected when the kernel stand-ins replace the wrappers)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available() or config.pluginmanager.has_plugin("tests.host_twin_plugin"):
        return  # (the host-twin plugin re-targets the GPU suites at CPU tensors + kernel stand-ins)
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_cls():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "classification.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_curves():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "curves.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_det():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "detection.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_reg():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "regression.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_binned():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "binned.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_multilabel():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "multilabel.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_consumers():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "consumers.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_fuzz():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "fuzz.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_atfixed():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "atfixed.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_logauc():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "logauc.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_coco_format():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "coco_format.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_fairness():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "fairness.npz"), allow_pickle=False)


@pytest.fixture
def cpu_kernel_standins(monkeypatch):
    """HOST-LAYER tests only.  For the duration of one test the kernel wrappers of `metrics_b200._native` are replaced by
    the torch-CPU stand-ins of tests/reference_runtime/cpu_kernels.py (same contracts), so that the Python above the C-ABI
    — validation, format seams, state handling, reducers — can be checked against the reference's goldens without a GPU.
    Like `oracle/`, the stand-ins are test infrastructure: nothing under `metrics_b200/` can reach them, and the product
    raises `NativeLibraryError` on CPU tensors (tests/test_native_abi.py)."""
    import importlib.util

    from metrics_b200 import _native

    spec = importlib.util.spec_from_file_location(
        "mb200_cpu_kernels", os.path.join(os.path.dirname(__file__), "reference_runtime", "cpu_kernels.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    for name, fn in module.standins().items():
        monkeypatch.setattr(_native, name, fn)
    return module


@pytest.fixture(scope="session")
def golden_tweedie():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "tweedie.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_fuzz2():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "fuzz2.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_csi():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "csi.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_kld():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "kld.npz"), allow_pickle=False)
