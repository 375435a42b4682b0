"""Import-path alias: the reference keeps these in `torchmetrics/classification/recall_fixed_precision.py`; here they are rows of the
table-driven `at_fixed` module."""
from metrics_b200.classification.at_fixed import (  # noqa: F401
    BinaryRecallAtFixedPrecision,
    MulticlassRecallAtFixedPrecision,
    MultilabelRecallAtFixedPrecision,
    RecallAtFixedPrecision,
)
