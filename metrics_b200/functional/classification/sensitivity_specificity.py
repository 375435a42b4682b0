"""Import-path alias: the reference keeps these in `torchmetrics/functional/classification/sensitivity_specificity.py`; here they are rows of the
table-driven `at_fixed` module."""
from metrics_b200.functional.classification.at_fixed import (  # noqa: F401
    _convert_fpr_to_specificity,
    binary_sensitivity_at_specificity,
    multiclass_sensitivity_at_specificity,
    multilabel_sensitivity_at_specificity,
    sensitivity_at_specificity,
)
