"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/negative_predictive_value.py`; here they are rows of the
table-driven `ratio_metrics` module."""
from metrics_b200.functional.classification.ratio_metrics import (  # noqa: F401
    binary_negative_predictive_value,
    multiclass_negative_predictive_value,
    multilabel_negative_predictive_value,
    negative_predictive_value,
)
