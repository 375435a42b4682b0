"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/log_cosh.py`; here they live in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _check_data_shape_to_num_outputs,
    _log_cosh_error_compute,
    _log_cosh_error_update,
    log_cosh_error,
)
