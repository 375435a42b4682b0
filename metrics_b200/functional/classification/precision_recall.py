"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/precision_recall.py`; here they are rows of the
table-driven `ratio_metrics` module."""
from metrics_b200.functional.classification.ratio_metrics import (  # noqa: F401
    binary_precision,
    binary_recall,
    multiclass_precision,
    multiclass_recall,
    multilabel_precision,
    multilabel_recall,
    precision,
    recall,
)
