"""Average-precision metric classes (reference: classification/average_precision.py)."""
from __future__ import annotations

from typing import Any, List, Optional, Union

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification import precision_recall_curve as _prc
from metrics_b200.classification._curve_common import _RankingScore, build_for_task, finish_score_init
from metrics_b200.functional.classification.average_precision import (
    _binary_average_precision_compute,
    _multiclass_average_precision_arg_validation,
    _multiclass_average_precision_compute,
    _multilabel_average_precision_arg_validation,
    _multilabel_average_precision_compute,
)


class BinaryAveragePrecision(_RankingScore, _prc.BinaryPrecisionRecallCurve):
    """Reference :47-119."""


    def compute(self) -> Tensor:
        return _binary_average_precision_compute(self._state(), self.thresholds, scalars=self._curve_scalars())


class MulticlassAveragePrecision(_RankingScore, _prc.MulticlassPrecisionRecallCurve):
    """Reference :170-290."""

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
        finish_score_init(self, average, validate_args,
                          lambda: _multiclass_average_precision_arg_validation(num_classes, average, thresholds, ignore_index))

    def _compute_local(self) -> Tensor:
        scalars = self._curve_scalars(self.num_classes)
        if getattr(self, "_sharded_now", False):
            # stand-in `target` for the reference's `(target == 0).all()` guard: the class-0 counts say it
            counts = scalars[2]
            all_zero = bool(counts[0, 1] == 0)  # no sample is a negative for class 0
            state = (None, torch.zeros(1, dtype=torch.int64) if all_zero else torch.ones(1, dtype=torch.int64))
        else:
            state = self._state()
        return _multiclass_average_precision_compute(state, self.num_classes, self.average, self.thresholds, scalars=scalars)

    def compute(self) -> Tensor:
        return self._compute_local()


class MultilabelAveragePrecision(_RankingScore, _prc.MultilabelPrecisionRecallCurve):
    """Reference :293-441."""

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
                          lambda: _multilabel_average_precision_arg_validation(num_labels, average, thresholds, ignore_index))

    def compute(self) -> Tensor:
        scalars = None if self.average == "micro" else self._curve_scalars()
        return _multilabel_average_precision_compute(self._state(), self.num_labels, self.average, self.thresholds,
                                                     self.ignore_index, scalars=scalars)


from metrics_b200.classification.base import _ClassificationTaskWrapper  # noqa: E402
from metrics_b200.metric import Metric  # noqa: E402


class AveragePrecision(_ClassificationTaskWrapper):
    """Task wrapper (reference :430-544)."""

    def __new__(  # type: ignore[misc]
        cls,
        task: Literal["binary", "multiclass", "multilabel"],
        thresholds: Optional[Union[int, List[float], Tensor]] = None,
        num_classes: Optional[int] = None,
        num_labels: Optional[int] = None,
        average: Optional[Literal["macro", "weighted", "none"]] = "macro",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> Metric:
        shared = dict(kwargs, thresholds=thresholds, ignore_index=ignore_index, validate_args=validate_args)
        return build_for_task(task, num_classes, num_labels,
                              lambda: BinaryAveragePrecision(**shared),
                              lambda c: MulticlassAveragePrecision(c, average, **shared),
                              lambda n: MultilabelAveragePrecision(n, average, **shared))
