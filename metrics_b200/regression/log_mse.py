"""Import-path alias: the reference keeps these in `torchmetrics/regression/log_mse.py`; here they live in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.regression.metrics import (  # noqa: F401
    MeanSquaredLogError,
)
