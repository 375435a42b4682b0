"""AUROC metric classes (reference: classification/auroc.py)."""
from __future__ import annotations

from typing import Any, List, Optional, Union

from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification import precision_recall_curve as _prc
from metrics_b200.classification._curve_common import _RankingScore, build_for_task, finish_score_init
from metrics_b200.functional.classification.auroc import (
    _binary_auroc_arg_validation,
    _binary_auroc_compute,
    _multiclass_auroc_arg_validation,
    _multiclass_auroc_compute,
    _multilabel_auroc_arg_validation,
    _multilabel_auroc_compute,
)


class BinaryAUROC(_RankingScore, _prc.BinaryPrecisionRecallCurve):
    """Reference :44-125."""

    def __init__(
        self,
        max_fpr: Optional[float] = None,
        thresholds: Optional[Union[int, List[float], Tensor]] = None,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        super().__init__(thresholds=thresholds, ignore_index=ignore_index, validate_args=False, **kwargs)
        if validate_args:
            _binary_auroc_arg_validation(max_fpr, thresholds, ignore_index)
        self.validate_args = validate_args
        self.max_fpr = max_fpr

    def compute(self) -> Tensor:
        if self.thresholds is None and (self.max_fpr is None or self.max_fpr == 1):
            return _binary_auroc_compute(None, self.thresholds, self.max_fpr, scalars=self._curve_scalars())
        return _binary_auroc_compute(self._state(), self.thresholds, self.max_fpr)


class MulticlassAUROC(_RankingScore, _prc.MulticlassPrecisionRecallCurve):
    """Reference :170-282."""

    plot_legend_name: str = "Class"

    def _compute_distributed(self):
        """Class-sharded multi-GPU evaluation (metrics_b200/parallel_curves.py) instead of all-gathering the score lists;
        returns ``NotImplemented`` when the generic sync has to be used."""
        from metrics_b200.parallel_curves import sharded_applicable

        if not sharded_applicable(self):
            return NotImplemented
        self._sharded_now = True
        try:
            return self._compute_local()
        finally:
            self._sharded_now = False

    def __init__(
        self,
        num_classes: int,
        average: Optional[Literal["macro", "weighted", "none"]] = "macro",
        thresholds: Optional[Union[int, List[float], Tensor]] = None,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        super().__init__(
            num_classes=num_classes, thresholds=thresholds, ignore_index=ignore_index, validate_args=False, **kwargs
        )
        if validate_args:
            _multiclass_auroc_arg_validation(num_classes, average, thresholds, ignore_index)
        self.average = average  # the parent stored its own (curve) `average=None`; this one drives the class reduction
        self.validate_args = validate_args

    def _compute_local(self) -> Tensor:
        if self.thresholds is not None:
            return _multiclass_auroc_compute(self._state(), self.num_classes, self.average, self.thresholds)
        return _multiclass_auroc_compute(None, self.num_classes, self.average, self.thresholds,
                                         scalars=self._curve_scalars(self.num_classes))

    def compute(self) -> Tensor:
        return self._compute_local()


from metrics_b200.classification.base import _ClassificationTaskWrapper  # noqa: E402
from metrics_b200.metric import Metric  # noqa: E402


class MultilabelAUROC(_RankingScore, _prc.MultilabelPrecisionRecallCurve):
    """Reference :281-429."""

    plot_legend_name: str = "Label"

    def __init__(
        self,
        num_labels: int,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        thresholds: Optional[Union[int, List[float], Tensor]] = None,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        super().__init__(num_labels=num_labels, thresholds=thresholds, ignore_index=ignore_index, validate_args=False, **kwargs)
        finish_score_init(self, average, validate_args,
                          lambda: _multilabel_auroc_arg_validation(num_labels, average, thresholds, ignore_index))

    def compute(self) -> Tensor:
        scalars = None if self.average == "micro" else self._curve_scalars()
        return _multilabel_auroc_compute(self._state(), self.num_labels, self.average, self.thresholds, self.ignore_index,
                                         scalars=scalars)


class AUROC(_ClassificationTaskWrapper):
    """Task wrapper (reference :432-547)."""

    def __new__(  # type: ignore[misc]
        cls,
        task: Literal["binary", "multiclass", "multilabel"],
        thresholds: Optional[Union[int, List[float], Tensor]] = None,
        num_classes: Optional[int] = None,
        num_labels: Optional[int] = None,
        average: Optional[Literal["macro", "weighted", "none"]] = "macro",
        max_fpr: Optional[float] = None,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> Metric:
        shared = dict(kwargs, thresholds=thresholds, ignore_index=ignore_index, validate_args=validate_args)
        return build_for_task(task, num_classes, num_labels,
                              lambda: BinaryAUROC(max_fpr, **shared),
                              lambda c: MulticlassAUROC(c, average, **shared),
                              lambda n: MultilabelAUROC(n, average, **shared))
