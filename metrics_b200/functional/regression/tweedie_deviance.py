"""Import-path alias: the reference keeps these in `torchmetrics/functional/regression/tweedie_deviance.py`; here they live
in `metrics.py` (one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    _tweedie_deviance_score_compute,
    _tweedie_deviance_score_update,
    tweedie_deviance_score,
)
