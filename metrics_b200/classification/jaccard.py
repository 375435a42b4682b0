"""Import-path alias: the reference keeps these in `torchmetrics/classification/jaccard.py`; here they are rows of the
table-driven `confmat_metrics` module."""
from metrics_b200.classification.confmat_metrics import (  # noqa: F401
    BinaryJaccardIndex,
    JaccardIndex,
    MulticlassJaccardIndex,
    MultilabelJaccardIndex,
    _jaccard_average_validation,
    _jaccard_index_reduce,
)
