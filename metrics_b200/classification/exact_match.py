"""ExactMatch metric classes (reference: classification/exact_match.py): states ``correct`` (int64 sum, or a ``cat`` list of
per-sample flags when samplewise) and ``total``."""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification.base import _ClassificationTaskWrapper
from metrics_b200.functional.classification.exact_match import (
    _exact_match_reduce,
    _label_values_check,
    _multiclass_exact_match_format,
    _multiclass_exact_match_update,
    _multilabel_exact_match_format,
    _multilabel_exact_match_update,
)
from metrics_b200.functional.classification.stat_scores import (
    _multiclass_stat_scores_arg_validation,
    _multiclass_stat_scores_tensor_validation,
    _multilabel_stat_scores_arg_validation,
    _multilabel_stat_scores_tensor_validation,
)
from metrics_b200.metric import Metric
from metrics_b200.utilities.data import dim_zero_cat
from metrics_b200.utilities.enums import ClassificationTaskNoBinary


class _ExactMatchBase(Metric):
    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0

    def _create_states(self) -> None:
        samplewise = self.multidim_average == "samplewise"
        self.add_state("correct", [] if samplewise else torch.zeros(1, dtype=torch.long),
                       dist_reduce_fx="cat" if samplewise else "sum")
        self.add_state("total", torch.zeros(1, dtype=torch.long), dist_reduce_fx="mean" if samplewise else "sum")

    def _accumulate(self, correct: Tensor, total: Tensor) -> None:
        if self.multidim_average == "samplewise":
            self.correct.append(correct)
            self.total = total
        else:
            self.correct += correct
            self.total += total

    def compute(self) -> Tensor:
        correct = dim_zero_cat(self.correct) if isinstance(self.correct, list) else self.correct
        return _exact_match_reduce(correct, self.total)


class MulticlassExactMatch(_ExactMatchBase):
    """Reference :45-197."""

    plot_legend_name: str = "Class"

    def __init__(self, num_classes: int, multidim_average: Literal["global", "samplewise"] = "global",
                 ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if validate_args:
            _multiclass_stat_scores_arg_validation(num_classes, 1, None, multidim_average, ignore_index)
        self.num_classes = num_classes
        self.multidim_average = multidim_average
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self._create_states()

    def update(self, preds: Tensor, target: Tensor) -> None:
        if self.validate_args:
            _multiclass_stat_scores_tensor_validation(preds, target, self.num_classes, self.multidim_average, self.ignore_index)
            _label_values_check(target, self.num_classes, self.ignore_index, "target")
            if not preds.is_floating_point():
                _label_values_check(preds, self.num_classes, None, "preds")
        preds, target = _multiclass_exact_match_format(preds, target)
        self._accumulate(*_multiclass_exact_match_update(preds, target, self.multidim_average, self.ignore_index))


class MultilabelExactMatch(_ExactMatchBase):
    """Reference :200-366."""

    plot_legend_name: str = "Label"

    def __init__(self, num_labels: int, threshold: float = 0.5,
                 multidim_average: Literal["global", "samplewise"] = "global", ignore_index: Optional[int] = None,
                 validate_args: bool = True, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if validate_args:
            _multilabel_stat_scores_arg_validation(num_labels, threshold, None, multidim_average, ignore_index)
        self.num_labels = num_labels
        self.threshold = threshold
        self.multidim_average = multidim_average
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self._create_states()

    def update(self, preds: Tensor, target: Tensor) -> None:
        if self.validate_args:
            _multilabel_stat_scores_tensor_validation(preds, target, self.num_labels, self.multidim_average, self.ignore_index)
        preds, target = _multilabel_exact_match_format(preds, target, self.num_labels, self.threshold, self.ignore_index)
        self._accumulate(*_multilabel_exact_match_update(preds, target, self.num_labels, self.multidim_average))


class ExactMatch(_ClassificationTaskWrapper):
    """Task wrapper (reference :369-430); multiclass and multilabel only."""

    def __new__(cls, task: Literal["multiclass", "multilabel"], threshold: float = 0.5,  # type: ignore[misc]
                num_classes: Optional[int] = None, num_labels: Optional[int] = None,
                multidim_average: Literal["global", "samplewise"] = "global", ignore_index: Optional[int] = None,
                validate_args: bool = True, **kwargs: Any) -> Metric:
        task = ClassificationTaskNoBinary.from_str(task)
        kwargs.update({"multidim_average": multidim_average, "ignore_index": ignore_index, "validate_args": validate_args})
        if task == ClassificationTaskNoBinary.MULTICLASS:
            if not isinstance(num_classes, int):
                raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
            return MulticlassExactMatch(num_classes, **kwargs)
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return MultilabelExactMatch(num_labels, threshold, **kwargs)
