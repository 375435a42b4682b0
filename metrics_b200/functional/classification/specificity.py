"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/specificity.py`; here they are rows of the
table-driven `ratio_metrics` module."""
from metrics_b200.functional.classification.ratio_metrics import (  # noqa: F401
    binary_specificity,
    multiclass_specificity,
    multilabel_specificity,
    specificity,
)
