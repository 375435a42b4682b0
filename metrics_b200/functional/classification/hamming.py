"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/hamming.py`; here they are rows of the
table-driven `ratio_metrics` module."""
from metrics_b200.functional.classification.ratio_metrics import (  # noqa: F401
    binary_hamming_distance,
    hamming_distance,
    multiclass_hamming_distance,
    multilabel_hamming_distance,
)
