"""Import-path alias: the reference keeps these in `torchmetrics/classification/negative_predictive_value.py`; here they are rows of the
table-driven `ratio_metrics` module."""
from metrics_b200.classification.ratio_metrics import (  # noqa: F401
    BinaryNegativePredictiveValue,
    MulticlassNegativePredictiveValue,
    MultilabelNegativePredictiveValue,
    NegativePredictiveValue,
)
