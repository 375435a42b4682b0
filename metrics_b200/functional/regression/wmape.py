"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/wmape.py`; here they live in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _weighted_mean_absolute_percentage_error_compute,
    _weighted_mean_absolute_percentage_error_update,
    weighted_mean_absolute_percentage_error,
)
