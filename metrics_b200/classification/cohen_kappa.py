"""Import-path alias: the reference keeps these in `torchmetrics/classification/cohen_kappa.py`; here they are rows of the
table-driven `confmat_metrics` module."""
from metrics_b200.classification.confmat_metrics import (  # noqa: F401
    BinaryCohenKappa,
    CohenKappa,
    MulticlassCohenKappa,
    _cohen_kappa_reduce,
    _cohen_kappa_weights_validation,
)
