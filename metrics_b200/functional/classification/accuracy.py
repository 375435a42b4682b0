"""Accuracy reducers and functionals (reference: functional/classification/accuracy.py)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.stat_scores import (
    _multiclass_stat_scores_arg_validation,
    _multiclass_stat_scores_tensor_validation,
    _multiclass_stat_scores_states,
)
from metrics_b200.utilities.compute import _adjust_weights_safe_divide, _safe_divide


def _accuracy_reduce(
    tp: Tensor,
    fp: Tensor,
    tn: Tensor,
    fn: Tensor,
    average: Optional[str],
    multidim_average: str = "global",
    multilabel: bool = False,
    top_k: int = 1,
) -> Tensor:
    """tp/fp/tn/fn -> accuracy (reference :37-88).  Counts are widened to f32 and divided once."""
    axis = 0 if multidim_average == "global" else 1
    if average == "binary":
        return _safe_divide(tp + tn, tp + tn + fp + fn)
    if average == "micro":
        tp, fn = tp.sum(dim=axis), fn.sum(dim=axis)
        if not multilabel:
            return _safe_divide(tp, tp + fn)
        fp, tn = fp.sum(dim=axis), tn.sum(dim=axis)
        return _safe_divide(tp + tn, tp + tn + fp + fn)
    per_class = _safe_divide(tp + tn, tp + tn + fp + fn) if multilabel else _safe_divide(tp, tp + fn)
    return _adjust_weights_safe_divide(per_class, average, multilabel, tp, fp, fn, top_k)


def multiclass_accuracy(
    preds: Tensor,
    target: Tensor,
    num_classes: Optional[int] = None,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    top_k: int = 1,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Multiclass accuracy (reference :263-369)."""
    if validate_args:
        _multiclass_stat_scores_arg_validation(num_classes, top_k, average, multidim_average, ignore_index)
        _multiclass_stat_scores_tensor_validation(preds, target, num_classes, multidim_average, ignore_index)
    if num_classes is None:  # micro only (validated above; reference :265-270 passes `num_classes or 1` on)
        if average != "micro" or multidim_average != "global" or top_k != 1:
            raise NotImplementedError("`num_classes=None` is supported for global top-1 micro accuracy only")
        from metrics_b200.functional.classification.stat_scores import _multiclass_micro_update_unknown_classes_

        states = [torch.zeros(1, dtype=torch.int64, device=preds.device) for _ in range(4)]
        _multiclass_micro_update_unknown_classes_(*states, preds, target, ignore_index)
        states = [s_.reshape(()) for s_ in states]
    else:
        states = _multiclass_stat_scores_states(preds, target, num_classes, top_k, average, multidim_average, ignore_index, validate_args)
    return _accuracy_reduce(*states, average=average, multidim_average=multidim_average, top_k=top_k)


# ---- binary / multilabel ---------------------------------------------------------------------------------
from metrics_b200.functional.classification.stat_scores import (  # noqa: E402
    _binary_stat_scores_arg_validation,
    _binary_stat_scores_tensor_validation,
    _binary_stat_scores_update,
    _multilabel_stat_scores_arg_validation,
    _multilabel_stat_scores_tensor_validation,
    _multilabel_stat_scores_update,
)


def binary_accuracy(
    preds: Tensor,
    target: Tensor,
    threshold: float = 0.5,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Reference :91-160."""
    if validate_args:
        _binary_stat_scores_arg_validation(threshold, multidim_average, ignore_index)
        _binary_stat_scores_tensor_validation(preds, target, multidim_average, ignore_index)
    tp, fp, tn, fn = _binary_stat_scores_update(preds, target, threshold, multidim_average, ignore_index, validate_args)
    return _accuracy_reduce(tp, fp, tn, fn, average="binary", multidim_average=multidim_average)


def multilabel_accuracy(
    preds: Tensor,
    target: Tensor,
    num_labels: int,
    threshold: float = 0.5,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Reference :372-470."""
    if validate_args:
        _multilabel_stat_scores_arg_validation(num_labels, threshold, average, multidim_average, ignore_index)
        _multilabel_stat_scores_tensor_validation(preds, target, num_labels, multidim_average, ignore_index)
    tp, fp, tn, fn = _multilabel_stat_scores_update(
        preds, target, num_labels, threshold, multidim_average, ignore_index, validate_args
    )
    return _accuracy_reduce(tp, fp, tn, fn, average=average, multidim_average=multidim_average, multilabel=True)


def accuracy(preds: Tensor, target: Tensor, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,
             num_classes: Optional[int] = None, num_labels: Optional[int] = None,
             average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
             multidim_average: Literal["global", "samplewise"] = "global", top_k: Optional[int] = 1,
             ignore_index: Optional[int] = None, validate_args: bool = True) -> Tensor:
    """Task wrapper (reference :473-548)."""
    from metrics_b200.functional.classification._task import call_for_task

    def mc(c: int) -> Tensor:
        if not isinstance(top_k, int):
            raise ValueError(f"`top_k` is expected to be `int` but `{type(top_k)} was passed.`")
        return multiclass_accuracy(preds, target, c, average, top_k, multidim_average, ignore_index, validate_args)

    return call_for_task(
        task, num_classes, num_labels,
        lambda: binary_accuracy(preds, target, threshold, multidim_average, ignore_index, validate_args), mc,
        lambda n: multilabel_accuracy(preds, target, n, threshold, average, multidim_average, ignore_index, validate_args))
