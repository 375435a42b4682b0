"""LogAUC metric classes (reference: classification/logauc.py): ROC state holders whose `compute` integrates the curve over a
logarithmic false-positive-rate window."""
from __future__ import annotations

from typing import Any, List, Optional, Tuple, Union

from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification.base import _ClassificationTaskWrapper
from metrics_b200.classification.roc import BinaryROC, MulticlassROC, MultilabelROC
from metrics_b200.functional.classification.logauc import _binary_logauc_compute, _reduce_logauc, _validate_fpr_range
from metrics_b200.metric import Metric
from metrics_b200.utilities.enums import ClassificationTask

_Thr = Optional[Union[int, List[float], Tensor]]


class BinaryLogAUC(BinaryROC):
    """Reference :35-159."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0

    def __init__(self, fpr_range: Tuple[float, float] = (0.001, 0.1), thresholds: _Thr = None,
                 ignore_index: Optional[int] = None, validate_args: bool = False, **kwargs: Any) -> None:
        super().__init__(thresholds=thresholds, ignore_index=ignore_index, validate_args=validate_args, **kwargs)
        if validate_args:
            _validate_fpr_range(fpr_range)
        self.fpr_range = fpr_range

    def compute(self) -> Tensor:  # type: ignore[override]
        fpr, tpr, _ = super().compute()
        return _binary_logauc_compute(fpr, tpr, fpr_range=self.fpr_range)


class MulticlassLogAUC(MulticlassROC):
    """Reference :162-309.  The per-class averaging mode is kept in ``average2`` (``average`` belongs to the ROC base)."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Class"

    def __init__(self, num_classes: int, fpr_range: Tuple[float, float] = (0.001, 0.1),
                 average: Optional[Literal["macro", "none"]] = None, thresholds: _Thr = None,
                 ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> None:
        super().__init__(num_classes=num_classes, thresholds=thresholds, average=None, ignore_index=ignore_index,
                         validate_args=validate_args, **kwargs)
        if validate_args:
            _validate_fpr_range(fpr_range)
        self.fpr_range = fpr_range
        self.average2 = average

    def compute(self) -> Tensor:  # type: ignore[override]
        fpr, tpr, _ = super().compute()
        return _reduce_logauc(fpr, tpr, fpr_range=self.fpr_range, average=self.average2)


class MultilabelLogAUC(MultilabelROC):
    """Reference :312-459."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Label"

    def __init__(self, num_labels: int, fpr_range: Tuple[float, float] = (0.001, 0.1),
                 average: Optional[Literal["macro", "none"]] = None, thresholds: _Thr = None,
                 ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> None:
        super().__init__(num_labels=num_labels, thresholds=thresholds, ignore_index=ignore_index,
                         validate_args=validate_args, **kwargs)
        if validate_args:
            _validate_fpr_range(fpr_range)
        self.fpr_range = fpr_range
        self.average2 = average

    def compute(self) -> Tensor:  # type: ignore[override]
        fpr, tpr, _ = super().compute()
        return _reduce_logauc(fpr, tpr, fpr_range=self.fpr_range, average=self.average2)


class LogAUC(_ClassificationTaskWrapper):
    """Task wrapper (reference :462-528)."""

    def __new__(cls, task: Literal["binary", "multiclass", "multilabel"], thresholds: _Thr = None,  # type: ignore[misc]
                fpr_range: Tuple[float, float] = (0.001, 0.1), num_classes: Optional[int] = None,
                num_labels: Optional[int] = None, ignore_index: Optional[int] = None, validate_args: bool = True,
                **kwargs: Any) -> Metric:
        task = ClassificationTask.from_str(task)
        kwargs.update({"thresholds": thresholds, "fpr_range": fpr_range, "ignore_index": ignore_index,
                       "validate_args": validate_args})
        if task == ClassificationTask.BINARY:
            return BinaryLogAUC(**kwargs)
        if task == ClassificationTask.MULTICLASS:
            if not isinstance(num_classes, int):
                raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
            return MulticlassLogAUC(num_classes, **kwargs)  # `average`, if any, travels in kwargs like in the reference
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return MultilabelLogAUC(num_labels, **kwargs)
