"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/mape.py`; here they live in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _mean_absolute_percentage_error_compute,
    _mean_absolute_percentage_error_update,
    mean_absolute_percentage_error,
)
