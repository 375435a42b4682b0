"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/precision_fixed_recall.py`; here they are rows of the
table-driven `at_fixed` module."""
from metrics_b200.functional.classification.at_fixed import (  # noqa: F401
    binary_precision_at_fixed_recall,
    multiclass_precision_at_fixed_recall,
    multilabel_precision_at_fixed_recall,
    precision_at_fixed_recall,
)
