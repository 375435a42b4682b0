"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/minkowski.py`; here they live in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _minkowski_distance_compute,
    _minkowski_distance_update,
    minkowski_distance,
)
