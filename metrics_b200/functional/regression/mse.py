"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/mse.py`; here they live in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _mean_squared_error_compute,
    _mean_squared_error_update,
    mean_squared_error,
)
