"""Import-path alias: the reference keeps these in `torchmetrics/classification/sensitivity_specificity.py`; here they are rows of the
table-driven `at_fixed` module."""
from metrics_b200.classification.at_fixed import (  # noqa: F401
    BinarySensitivityAtSpecificity,
    MulticlassSensitivityAtSpecificity,
    MultilabelSensitivityAtSpecificity,
    SensitivityAtSpecificity,
)
