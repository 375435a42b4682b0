"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/recall_fixed_precision.py`; here they are rows of the
table-driven `at_fixed` module."""
from metrics_b200.functional.classification.at_fixed import (  # noqa: F401
    binary_recall_at_fixed_precision,
    multiclass_recall_at_fixed_precision,
    multilabel_recall_at_fixed_precision,
    recall_at_fixed_precision,
)
