"""Import-path alias: the reference keeps these in `torchmetrics/classification/matthews_corrcoef.py`; here they are rows of the
table-driven `confmat_metrics` module."""
from metrics_b200.classification.confmat_metrics import (  # noqa: F401
    BinaryMatthewsCorrCoef,
    MatthewsCorrCoef,
    MulticlassMatthewsCorrCoef,
    MultilabelMatthewsCorrCoef,
    _matthews_corrcoef_reduce,
)
