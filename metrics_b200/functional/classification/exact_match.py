"""Exact match (subset accuracy) for multiclass / multilabel inputs with extra dimensions.

Reference: functional/classification/exact_match.py.  The per-sample "all positions agree" test is a handful of device ops
on label tensors; the class dimension is reduced by the K1 argmax kernel (`mb200_argmax_rows`) and logits are normalised
by the K6 kernel, like in the stat-scores family.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200 import _native
from metrics_b200.functional.classification.stat_scores import (
    _multiclass_stat_scores_arg_validation,
    _multiclass_stat_scores_tensor_validation,
    _multilabel_stat_scores_arg_validation,
    _multilabel_stat_scores_tensor_validation,
)
from metrics_b200.utilities.compute import _safe_divide


def _exact_match_reduce(correct: Tensor, total: Tensor) -> Tensor:
    """Reference :32-37."""
    return _safe_divide(correct, total)


def _label_values_check(x: Tensor, num_classes: int, ignore_index: Optional[int], name: str) -> None:
    """The reference counts `unique` values (stat_scores.py:300-314); here: one fused range test, one host read."""
    bad = (x < 0) | (x >= num_classes)
    if ignore_index is not None:
        bad &= x != ignore_index
    if bool(bad.any()):
        raise RuntimeError(
            f"Detected more unique values in `{name}` than expected. Expected only {num_classes if ignore_index is None else num_classes + 1}"
            f" but found values outside of [0, {num_classes}) in `{name}`."
        )


def _multiclass_exact_match_format(preds: Tensor, target: Tensor) -> tuple[Tensor, Tensor]:
    """``[N, C, ...]`` scores -> ``[N, P]`` labels via the argmax kernel; label inputs are only reshaped
    (reference stat_scores.py:328-344 with top_k = 1)."""
    if preds.ndim == target.ndim + 1:
        preds = _native.argmax_rows(preds)
    return preds.reshape(preds.shape[0], -1), target.reshape(target.shape[0], -1)


def _multiclass_exact_match_update(
    preds: Tensor, target: Tensor, multidim_average: str = "global", ignore_index: Optional[int] = None
) -> tuple[Tensor, Tensor]:
    """Reference :40-54: a sample is correct when every (non-ignored) position matches."""
    agree = preds == target
    if ignore_index is not None:
        agree = agree | (target == ignore_index)
    correct = agree.all(dim=1)
    correct = correct.to(torch.int64) if multidim_average == "samplewise" else correct.sum()
    total = torch.tensor(preds.shape[0] if multidim_average == "global" else 1, device=preds.device)
    return correct, total


def multiclass_exact_match(
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Reference :57-121."""
    if validate_args:
        _multiclass_stat_scores_arg_validation(num_classes, 1, None, multidim_average, ignore_index)
        _multiclass_stat_scores_tensor_validation(preds, target, num_classes, multidim_average, ignore_index)
        _label_values_check(target, num_classes, ignore_index, "target")
        if not preds.is_floating_point():
            _label_values_check(preds, num_classes, None, "preds")
    preds, target = _multiclass_exact_match_format(preds, target)
    correct, total = _multiclass_exact_match_update(preds, target, multidim_average, ignore_index)
    return _exact_match_reduce(correct, total)


def _multilabel_exact_match_format(
    preds: Tensor, target: Tensor, num_labels: int, threshold: float = 0.5, ignore_index: Optional[int] = None
) -> tuple[Tensor, Tensor]:
    """Reference stat_scores.py:681-702: sigmoid-if-logits + threshold, ``[N, L, P]`` layout, ignored targets -> -1
    (which can never equal a 0/1 prediction: an ignored position makes its sample incorrect, like in the reference)."""
    if preds.is_floating_point():
        preds = _native.sigmoid_if_logits(preds) > threshold
    preds = preds.reshape(*preds.shape[:2], -1)
    target = target.reshape(*target.shape[:2], -1)
    if ignore_index is not None:
        target = torch.where(target == ignore_index, torch.full_like(target, -1), target)
    return preds, target


def _multilabel_exact_match_update(
    preds: Tensor, target: Tensor, num_labels: int, multidim_average: str = "global"
) -> tuple[Tensor, Tensor]:
    """Reference :124-134."""
    if multidim_average == "global":
        preds = torch.movedim(preds, 1, -1).reshape(-1, num_labels)
        target = torch.movedim(target, 1, -1).reshape(-1, num_labels)
    correct = ((preds == target).sum(1) == num_labels).sum(dim=-1)
    total = torch.tensor(preds.shape[0 if multidim_average == "global" else 2], device=correct.device)
    return correct, total


def multilabel_exact_match(
    preds: Tensor,
    target: Tensor,
    num_labels: int,
    threshold: float = 0.5,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Reference :137-205."""
    if validate_args:
        _multilabel_stat_scores_arg_validation(num_labels, threshold, None, multidim_average, ignore_index)
        _multilabel_stat_scores_tensor_validation(preds, target, num_labels, multidim_average, ignore_index)
    preds, target = _multilabel_exact_match_format(preds, target, num_labels, threshold, ignore_index)
    correct, total = _multilabel_exact_match_update(preds, target, num_labels, multidim_average)
    return _exact_match_reduce(correct, total)


def exact_match(
    preds: Tensor,
    target: Tensor,
    task: Literal["multiclass", "multilabel"],
    num_classes: Optional[int] = None,
    num_labels: Optional[int] = None,
    threshold: float = 0.5,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Task wrapper (reference :208-258)."""
    from metrics_b200.utilities.enums import ClassificationTaskNoBinary

    task_ = ClassificationTaskNoBinary.from_str(task)
    if task_ == ClassificationTaskNoBinary.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        return multiclass_exact_match(preds, target, num_classes, multidim_average, ignore_index, validate_args)
    if not isinstance(num_labels, int):
        raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
    return multilabel_exact_match(preds, target, num_labels, threshold, multidim_average, ignore_index, validate_args)
