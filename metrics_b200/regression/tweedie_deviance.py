"""Import-path alias: the reference keeps this in `torchmetrics/regression/tweedie_deviance.py`; here it lives in `metrics.py`
(one module for the whole running-sum family, all served by kernel K9)."""
from metrics_b200.regression.metrics import (  # noqa: F401
    TweedieDevianceScore,
)
