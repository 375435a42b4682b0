"""CPU: the C-ABI shared library loads and exports every symbol declared in include/metrics_b200.h."""
import ctypes
import os
import re

from tests.conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "metrics_b200.h")).read()
    return sorted(set(re.findall(r"MB200_API[^;(]*?\b(mb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from metrics_b200 import _native

    assert os.path.exists(_native.lib_path()), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = ctypes.CDLL(_native.lib_path())
    names = _declared_symbols()
    assert len(names) >= 6
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"


def test_abi_version_and_error_string():
    from metrics_b200 import _native

    lib = _native.lib()
    assert lib.mb200_abi_version() == 1
    assert isinstance(lib.mb200_last_error(), bytes)


def test_cpu_tensors_are_rejected_loudly():
    import pytest
    import torch

    from metrics_b200.classification import MulticlassConfusionMatrix
    from metrics_b200._native import NativeLibraryError

    m = MulticlassConfusionMatrix(num_classes=3, validate_args=False)
    with pytest.raises(NativeLibraryError, match="no CPU fallback"):
        m.update(torch.randn(4, 3), torch.tensor([0, 1, 2, 0]))


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "metrics_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product code must not import oracle/: {bad}"
