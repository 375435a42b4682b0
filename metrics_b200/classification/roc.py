"""ROC metric classes (reference: classification/roc.py)."""
from __future__ import annotations

from torch import Tensor

from metrics_b200.classification.precision_recall_curve import (
    BinaryPrecisionRecallCurve,
    MulticlassPrecisionRecallCurve,
    MultilabelPrecisionRecallCurve,
)
from metrics_b200.functional.classification.roc import _binary_roc_compute, _multiclass_roc_compute, _multilabel_roc_compute


class BinaryROC(BinaryPrecisionRecallCurve):
    """Reference :44-165."""

    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0

    def compute(self) -> tuple[Tensor, Tensor, Tensor]:
        return _binary_roc_compute(self._state(), self.thresholds)


class MulticlassROC(MulticlassPrecisionRecallCurve):
    """Reference :168-330."""

    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Class"

    def compute(self):
        return _multiclass_roc_compute(self._state(), self.num_classes, self.thresholds, self.average)


class MultilabelROC(MultilabelPrecisionRecallCurve):
    """Reference :333-497."""

    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Label"

    def compute(self):
        return _multilabel_roc_compute(self._state(), self.num_labels, self.thresholds, self.ignore_index)


from typing import Any, List, Optional, Union  # noqa: E402

from typing_extensions import Literal  # noqa: E402

from metrics_b200.classification.base import _ClassificationTaskWrapper  # noqa: E402
from metrics_b200.metric import Metric  # noqa: E402
from metrics_b200.classification._curve_common import build_for_task  # noqa: E402


class ROC(_ClassificationTaskWrapper):
    """Task wrapper (reference :500-596)."""

    def __new__(  # type: ignore[misc]
        cls,
        task: Literal["binary", "multiclass", "multilabel"],
        thresholds: Optional[Union[int, List[float], Tensor]] = None,
        num_classes: Optional[int] = None,
        num_labels: Optional[int] = None,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> Metric:
        shared = dict(kwargs, thresholds=thresholds, ignore_index=ignore_index, validate_args=validate_args)
        return build_for_task(task, num_classes, num_labels, lambda: BinaryROC(**shared),
                              lambda c: MulticlassROC(c, **shared), lambda n: MultilabelROC(n, **shared))
