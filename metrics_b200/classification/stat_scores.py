"""Stat-scores metric classes (reference: classification/stat_scores.py)."""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.stat_scores import (
    _multiclass_stat_scores_arg_validation,
    _multiclass_stat_scores_compute,
    _multiclass_stat_scores_tensor_validation,
    _multiclass_stat_scores_update_,
    _require_kernel_mode,
    stat_scores_workspace,
)
from metrics_b200.metric import Metric
from metrics_b200.utilities.data import dim_zero_cat


class _AbstractStatScores(Metric):
    """Holds the four counters ``tp, fp, tn, fn`` (reference :43-88): int64 tensors with ``sum`` reduction, or
    lists with ``cat`` reduction when ``multidim_average="samplewise"``."""


    def _create_state(self, size: int, multidim_average: str = "global") -> None:
        for name in ("tp", "fp", "tn", "fn"):
            if multidim_average == "samplewise":
                self.add_state(name, [], dist_reduce_fx="cat")
            else:
                self.add_state(name, torch.zeros(size, dtype=torch.long), dist_reduce_fx="sum")
        self._scratch: Optional[Tensor] = None

    def _workspace(self, n_slots: int, device: torch.device) -> Tensor:
        """Per-instance kernel scratch (NOT a metric state: never synced, saved or compared)."""
        ws = self._scratch
        if ws is None or ws.device != device or ws.numel() != 3 * n_slots + 2:
            ws = stat_scores_workspace(n_slots, device)
            self._scratch = ws
        return ws

    def _update_state(self, tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor) -> None:
        if self.multidim_average == "samplewise":
            self.tp.append(tp)
            self.fp.append(fp)
            self.tn.append(tn)
            self.fn.append(fn)
        else:
            self.tp += tp
            self.fp += fp
            self.tn += tn
            self.fn += fn

    def _final_state(self) -> tuple[Tensor, Tensor, Tensor, Tensor]:
        return dim_zero_cat(self.tp), dim_zero_cat(self.fp), dim_zero_cat(self.tn), dim_zero_cat(self.fn)


class MulticlassStatScores(_AbstractStatScores):
    """tp / fp / tn / fn / support for multiclass tasks (reference :198-352)."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = None
    full_state_update: bool = False

    def __init__(
        self,
        num_classes: Optional[int] = None,
        top_k: int = 1,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        zero_division = kwargs.pop("zero_division", 0)
        super(_AbstractStatScores, self).__init__(**kwargs)
        if validate_args:
            _multiclass_stat_scores_arg_validation(num_classes, top_k, average, multidim_average, ignore_index, zero_division)
        _require_kernel_mode(top_k, multidim_average)
        self.num_classes = num_classes
        self.top_k = top_k
        self.average = average
        self.multidim_average = multidim_average
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self.zero_division = zero_division
        self._create_state(
            size=1 if (average == "micro" and top_k == 1) else (num_classes or 1), multidim_average=multidim_average
        )

    def update(self, preds: Tensor, target: Tensor) -> None:
        if self.validate_args:
            _multiclass_stat_scores_tensor_validation(
                preds, target, self.num_classes, self.multidim_average, self.ignore_index
            )
        if self.num_classes is None:
            self._update_unknown_class_count(preds, target)
            return
        num_classes = self.num_classes
        if self.multidim_average == "samplewise":
            from metrics_b200.functional.classification.stat_scores import _multiclass_stat_scores_states

            self._update_state(*_multiclass_stat_scores_states(
                preds, target, num_classes, self.top_k, self.average, "samplewise", self.ignore_index, self.validate_args))
            return
        _multiclass_stat_scores_update_(
            self.tp, self.fp, self.tn, self.fn, self._workspace(num_classes, self.tp.device), preds, target,
            num_classes, self.top_k, self.average, self.multidim_average, self.ignore_index, self.validate_args,
        )

    def _update_unknown_class_count(self, preds: Tensor, target: Tensor) -> None:
        """``num_classes=None`` with ``average="micro"`` (functional/classification/stat_scores.py::_class_count_bound)."""
        if self.multidim_average == "samplewise" or self.top_k != 1:
            raise NotImplementedError("`num_classes=None` is supported for global top-1 micro statistics only")
        from metrics_b200.functional.classification.stat_scores import _multiclass_micro_update_unknown_classes_

        _multiclass_micro_update_unknown_classes_(self.tp, self.fp, self.tn, self.fn, preds, target, self.ignore_index)

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _multiclass_stat_scores_compute(tp, fp, tn, fn, self.average, self.multidim_average)


# =========================================================================================================
# binary / multilabel
# =========================================================================================================
from metrics_b200.classification.base import _ClassificationTaskWrapper  # noqa: E402
from metrics_b200.functional.classification.stat_scores import (  # noqa: E402
    _binary_stat_scores_arg_validation,
    _binary_stat_scores_compute,
    _binary_stat_scores_tensor_validation,
    _binary_stat_scores_update,
    _multilabel_stat_scores_arg_validation,
    _multilabel_stat_scores_compute,
    _multilabel_stat_scores_tensor_validation,
    _multilabel_stat_scores_update,
)
from metrics_b200.utilities.enums import ClassificationTask  # noqa: E402


class BinaryStatScores(_AbstractStatScores):
    """Reference :91-195."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = None
    full_state_update: bool = False

    def __init__(
        self,
        threshold: float = 0.5,
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        zero_division = kwargs.pop("zero_division", 0)
        super(_AbstractStatScores, self).__init__(**kwargs)
        if validate_args:
            _binary_stat_scores_arg_validation(threshold, multidim_average, ignore_index, zero_division)
        self.threshold = threshold
        self.multidim_average = multidim_average
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self.zero_division = zero_division
        self._create_state(size=1, multidim_average=multidim_average)

    def update(self, preds: Tensor, target: Tensor) -> None:
        if self.validate_args:
            _binary_stat_scores_tensor_validation(preds, target, self.multidim_average, self.ignore_index)
        tp, fp, tn, fn = _binary_stat_scores_update(
            preds, target, self.threshold, self.multidim_average, self.ignore_index, self.validate_args
        )
        self._update_state(tp, fp, tn, fn)

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _binary_stat_scores_compute(tp, fp, tn, fn, self.multidim_average)


class MultilabelStatScores(_AbstractStatScores):
    """Reference :355-500."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = None
    full_state_update: bool = False

    def __init__(
        self,
        num_labels: int,
        threshold: float = 0.5,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        zero_division = kwargs.pop("zero_division", 0)
        super(_AbstractStatScores, self).__init__(**kwargs)
        if validate_args:
            _multilabel_stat_scores_arg_validation(num_labels, threshold, average, multidim_average, ignore_index, zero_division)
        self.num_labels = num_labels
        self.threshold = threshold
        self.average = average
        self.multidim_average = multidim_average
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self.zero_division = zero_division
        self._create_state(size=num_labels, multidim_average=multidim_average)

    def update(self, preds: Tensor, target: Tensor) -> None:
        if self.validate_args:
            _multilabel_stat_scores_tensor_validation(preds, target, self.num_labels, self.multidim_average, self.ignore_index)
        tp, fp, tn, fn = _multilabel_stat_scores_update(
            preds, target, self.num_labels, self.threshold, self.multidim_average, self.ignore_index, self.validate_args
        )
        self._update_state(tp, fp, tn, fn)

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _multilabel_stat_scores_compute(tp, fp, tn, fn, self.average, self.multidim_average)


def _dispatch(cls_binary, cls_multiclass, cls_multilabel, task, threshold, num_classes, num_labels, average, top_k, kwargs,
              multiclass_extra=None):
    """Shared body of the task wrappers' ``__new__`` (reference e.g. classification/stat_scores.py:529-562)."""
    task = ClassificationTask.from_str(task)
    if task == ClassificationTask.BINARY:
        return cls_binary(threshold, **kwargs)
    if task == ClassificationTask.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        if not isinstance(top_k, int):
            raise ValueError(f"`top_k` is expected to be `int` but `{type(top_k)} was passed.`")
        return cls_multiclass(num_classes, top_k, average, **kwargs)
    if task == ClassificationTask.MULTILABEL:
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return cls_multilabel(num_labels, threshold, average, **kwargs)
    raise ValueError(f"Task {task} not supported!")


class StatScores(_ClassificationTaskWrapper):
    """Task wrapper (reference :503-562)."""

    def __new__(  # type: ignore[misc]
        cls,
        task: Literal["binary", "multiclass", "multilabel"],
        threshold: float = 0.5,
        num_classes: Optional[int] = None,
        num_labels: Optional[int] = None,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
        multidim_average: Optional[Literal["global", "samplewise"]] = "global",
        top_k: Optional[int] = 1,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> Metric:
        assert multidim_average is not None  # noqa: S101
        kwargs.update({"multidim_average": multidim_average, "ignore_index": ignore_index, "validate_args": validate_args})
        return _dispatch(BinaryStatScores, MulticlassStatScores, MultilabelStatScores, task, threshold, num_classes,
                         num_labels, average, top_k, kwargs)
