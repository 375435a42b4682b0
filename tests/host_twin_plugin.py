"""pytest plugin (`-p tests.host_twin_plugin`, used by tests/test_host_twins.py): re-run the GPU parity suites on CPU
tensors with the kernel wrappers of `metrics_b200._native` replaced by their torch stand-ins
(tests/reference_runtime/cpu_kernels.py).  Every module-level ``DEV = "cuda:0"`` of a collected test module becomes "cpu".
This checks the HOST layer (and the stand-ins themselves) against the same reference goldens the kernels are held to; it
says nothing about the kernels.  Test infrastructure only."""
import importlib.util
import os

import pytest

_HERE = os.path.dirname(os.path.abspath(__file__))


def _standins():
    spec = importlib.util.spec_from_file_location("mb200_cpu_kernels", os.path.join(_HERE, "reference_runtime", "cpu_kernels.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module.standins()


def pytest_collection_modifyitems(config, items):
    for item in items:
        module = getattr(item, "module", None)
        if module is not None and isinstance(getattr(module, "DEV", None), str) and module.DEV.startswith("cuda"):
            module.DEV = "cpu"
    # tests that drive the C-ABI library directly (not through a wrapper that has a stand-in) only make sense on the GPU box
    raw = [item for item in items if item.get_closest_marker("raw_abi")]
    if raw:
        config.hook.pytest_deselected(items=raw)
        items[:] = [item for item in items if item not in raw]


@pytest.fixture(autouse=True)
def _kernel_standins(monkeypatch):
    from metrics_b200 import _native

    for name, fn in _standins().items():
        monkeypatch.setattr(_native, name, fn)
