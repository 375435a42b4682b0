"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/specificity_sensitivity.py`; here they are rows of the
table-driven `at_fixed` module."""
from metrics_b200.functional.classification.at_fixed import (  # noqa: F401
    _convert_fpr_to_specificity,
    binary_specificity_at_sensitivity,
    multiclass_specificity_at_sensitivity,
    multilabel_specificity_at_sensitivity,
    specificity_at_sensitivity,
)
