"""ROC metric classes (reference: classification/roc.py)."""
from __future__ import annotations

from torch import Tensor

from metrics_b200.classification.precision_recall_curve import BinaryPrecisionRecallCurve, MulticlassPrecisionRecallCurve
from metrics_b200.functional.classification.roc import _binary_roc_compute, _multiclass_roc_compute


class BinaryROC(BinaryPrecisionRecallCurve):
    """Reference :44-165."""

    def compute(self) -> tuple[Tensor, Tensor, Tensor]:
        return _binary_roc_compute(self._state(), self.thresholds)


class MulticlassROC(MulticlassPrecisionRecallCurve):
    """Reference :168-330."""

    def compute(self):
        return _multiclass_roc_compute(self._state(), self.num_classes, self.thresholds, self.average)
