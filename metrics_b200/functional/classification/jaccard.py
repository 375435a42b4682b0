"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/jaccard.py`; here they are rows of the
table-driven `confmat_metrics` module."""
from metrics_b200.functional.classification.confmat_metrics import (  # noqa: F401
    _jaccard_average_validation,
    _jaccard_index_reduce,
    binary_jaccard_index,
    jaccard_index,
    multiclass_jaccard_index,
    multilabel_jaccard_index,
)
