"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/csi.py`; here they live in `metrics.py`."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _critical_success_index_compute,
    _critical_success_index_update,
    critical_success_index,
)
