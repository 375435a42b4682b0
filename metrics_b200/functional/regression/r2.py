"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/r2.py`; here they live in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _r2_score_compute,
    _r2_score_update,
    r2_score,
)
