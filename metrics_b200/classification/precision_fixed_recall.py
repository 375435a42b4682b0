"""Import-path alias: the reference keeps these in `torchmetrics/classification/precision_fixed_recall.py`; here they are rows of the
table-driven `at_fixed` module."""
from metrics_b200.classification.at_fixed import (  # noqa: F401
    BinaryPrecisionAtFixedRecall,
    MulticlassPrecisionAtFixedRecall,
    MultilabelPrecisionAtFixedRecall,
    PrecisionAtFixedRecall,
)
