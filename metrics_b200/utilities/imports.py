"""Availability / version flags under the names the reference exposes (`torchmetrics/utilities/imports.py:21-65`), for
code and tests that branch on them.  Only the flags that matter to the accelerated path are kept; each is a plain bool
evaluated once at import (the reference uses `lightning_utilities.RequirementCache`, a dependency this package avoids)."""
from __future__ import annotations

import sys
from importlib import metadata
from importlib.util import find_spec

from packaging.version import Version


def _at_least(dist: str, minimum: str) -> bool:
    try:
        return Version(Version(metadata.version(dist)).base_version) >= Version(minimum)
    except metadata.PackageNotFoundError:
        return False


def _available(module: str) -> bool:
    try:
        return find_spec(module) is not None
    except (ImportError, ValueError):
        return False


_PYTHON_VERSION = ".".join(str(v) for v in sys.version_info[:3])
_TORCH_GREATER_EQUAL_2_1 = _at_least("torch", "2.1.0")
_TORCH_GREATER_EQUAL_2_2 = _at_least("torch", "2.2.0")
_TORCH_GREATER_EQUAL_2_5 = _at_least("torch", "2.5.0")
_TORCH_LESS_THAN_2_6 = not _at_least("torch", "2.6.0")
_SCIPY_AVAILABLE = _available("scipy")
_SKLEARN_GREATER_EQUAL_1_3 = _at_least("scikit-learn", "1.3.0")
_PYCOCOTOOLS_AVAILABLE = _available("pycocotools")
_FASTER_COCO_EVAL_AVAILABLE = _available("faster_coco_eval")
_TORCHVISION_AVAILABLE = _available("torchvision")
_MATPLOTLIB_AVAILABLE = _available("matplotlib")
_PYTDC_AVAILABLE = _available("tdc")
_XLA_AVAILABLE = _available("torch_xla")
