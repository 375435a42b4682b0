"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/cohen_kappa.py`; here they are rows of the
table-driven `confmat_metrics` module."""
from metrics_b200.functional.classification.confmat_metrics import (  # noqa: F401
    _cohen_kappa_reduce,
    _cohen_kappa_weights_validation,
    binary_cohen_kappa,
    cohen_kappa,
    multiclass_cohen_kappa,
)
