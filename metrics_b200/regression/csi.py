"""Import-path alias: the reference keeps this in `torchmetrics/regression/csi.py`; here it lives in `metrics.py`."""
from metrics_b200.regression.metrics import (  # noqa: F401
    CriticalSuccessIndex,
)
