"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/matthews_corrcoef.py`; here they are rows of the
table-driven `confmat_metrics` module."""
from metrics_b200.functional.classification.confmat_metrics import (  # noqa: F401
    _matthews_corrcoef_reduce,
    binary_matthews_corrcoef,
    matthews_corrcoef,
    multiclass_matthews_corrcoef,
    multilabel_matthews_corrcoef,
)
