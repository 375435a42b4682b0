"""Metrics reduced from the confusion-matrix state: Jaccard index, Cohen's kappa, Matthews correlation coefficient.

Reference: functional/classification/{jaccard,cohen_kappa,matthews_corrcoef}.py.  The state comes from the K1 / K2 kernels
(`confmat` [C,C] / [2,2] / [L,2,2], int64); only the small float reduction lives here.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.confusion_matrix import (
    _binary_confusion_matrix_arg_validation,
    _multiclass_confusion_matrix_arg_validation,
    _multilabel_confusion_matrix_arg_validation,
    binary_confusion_matrix,
    multiclass_confusion_matrix,
    multilabel_confusion_matrix,
)
from metrics_b200.utilities.compute import _safe_divide


# ----------------------------------------------------------------------------------------------------------------------
# Jaccard index (reference jaccard.py:38-97)
# ----------------------------------------------------------------------------------------------------------------------
def _jaccard_index_reduce(confmat: Tensor, average: Optional[str], ignore_index: Optional[int] = None,
                          zero_division: float = 0.0) -> Tensor:
    """intersection / union per class (or label) from an un-normalised confusion matrix, then the class average."""
    allowed = ("binary", "micro", "macro", "weighted", "none", None)
    if average not in allowed:
        raise ValueError(f"The `average` has to be one of {list(allowed)}, got {average}.")
    cm = confmat.float()
    if average == "binary":
        return _safe_divide(cm[1, 1], cm[0, 1] + cm[1, 0] + cm[1, 1], zero_division)
    multilabel = cm.ndim == 3
    drop = ignore_index is not None and 0 <= ignore_index < cm.shape[0]
    if multilabel:
        inter = cm[:, 1, 1]
        union = inter + cm[:, 0, 1] + cm[:, 1, 0]
    else:
        inter = cm.diagonal()
        union = cm.sum(0) + cm.sum(1) - inter
    if average == "micro":
        total_union = union.sum() - (union[ignore_index] if drop else 0.0)
        return _safe_divide(inter.sum(), total_union, zero_division)
    score = _safe_divide(inter, union, zero_division)
    if average in (None, "none"):
        return score
    if average == "weighted":
        w = cm[:, 1, 1] + cm[:, 1, 0] if multilabel else cm.sum(1)
    else:  # macro: classes that never occur (and the ignored class) do not count
        w = torch.ones_like(score)
        if drop:
            w[ignore_index] = 0.0
        if not multilabel:
            w = torch.where(cm.sum(1) + cm.sum(0) == 0, torch.zeros_like(w), w)
    return ((w * score) / w.sum()).sum()


def _jaccard_average_validation(average: Optional[str]) -> None:
    allowed = ("micro", "macro", "weighted", "none", None)
    if average not in allowed:
        raise ValueError(f"Expected argument `average` to be one of {allowed}, but got {average}.")


def binary_jaccard_index(preds: Tensor, target: Tensor, threshold: float = 0.5, ignore_index: Optional[int] = None,
                         validate_args: bool = True, zero_division: float = 0.0) -> Tensor:
    """Reference jaccard.py:100-160."""
    cm = binary_confusion_matrix(preds, target, threshold, None, ignore_index, validate_args)
    return _jaccard_index_reduce(cm, "binary", zero_division=zero_division)


def multiclass_jaccard_index(preds: Tensor, target: Tensor, num_classes: int,
                             average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
                             ignore_index: Optional[int] = None, validate_args: bool = True,
                             zero_division: float = 0.0) -> Tensor:
    """Reference jaccard.py:176-250."""
    if validate_args:
        _multiclass_confusion_matrix_arg_validation(num_classes, ignore_index)
        _jaccard_average_validation(average)
    cm = multiclass_confusion_matrix(preds, target, num_classes, None, ignore_index, validate_args)
    return _jaccard_index_reduce(cm, average, ignore_index, zero_division)


def multilabel_jaccard_index(preds: Tensor, target: Tensor, num_labels: int, threshold: float = 0.5,
                             average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
                             ignore_index: Optional[int] = None, validate_args: bool = True,
                             zero_division: float = 0.0) -> Tensor:
    """Reference jaccard.py:267-345."""
    if validate_args:
        _multilabel_confusion_matrix_arg_validation(num_labels, threshold, ignore_index)
        _jaccard_average_validation(average)
    cm = multilabel_confusion_matrix(preds, target, num_labels, threshold, None, ignore_index, validate_args)
    return _jaccard_index_reduce(cm, average, zero_division=zero_division)


# ----------------------------------------------------------------------------------------------------------------------
# Cohen's kappa (reference cohen_kappa.py:33-54)
# ----------------------------------------------------------------------------------------------------------------------
def _cohen_kappa_weights_validation(weights: Optional[str]) -> None:
    allowed = ("linear", "quadratic", "none", None)
    if weights not in allowed:
        raise ValueError(f"Expected argument `weight` to be one of {allowed}, but got {weights}.")


def _cohen_kappa_reduce(confmat: Tensor, weights: Optional[str] = None) -> Tensor:
    """1 - sum(W * observed) / sum(W * expected) with W = off-diagonal ones, |i-j| or (i-j)^2."""
    cm = confmat if confmat.is_floating_point() else confmat.float()
    c = cm.shape[0]
    col = cm.sum(dim=0, keepdim=True)
    row = cm.sum(dim=1, keepdim=True)
    expected = row @ col / col.sum()
    if weights is None or weights == "none":
        w = 1.0 - torch.eye(c, dtype=cm.dtype, device=cm.device)
    elif weights in ("linear", "quadratic"):
        idx = torch.arange(c, dtype=cm.dtype, device=cm.device)
        diff = idx[None, :] - idx[:, None]
        w = diff.abs() if weights == "linear" else diff.pow(2.0)
    else:
        raise ValueError(f"Received {weights} for argument ``weights`` but should be either None, 'linear' or 'quadratic'")
    return 1 - torch.sum(w * cm) / torch.sum(w * expected)


def binary_cohen_kappa(preds: Tensor, target: Tensor, threshold: float = 0.5,
                       weights: Optional[Literal["linear", "quadratic", "none"]] = None, ignore_index: Optional[int] = None,
                       validate_args: bool = True) -> Tensor:
    """Reference cohen_kappa.py:73-135."""
    if validate_args:
        _binary_confusion_matrix_arg_validation(threshold, ignore_index)
        _cohen_kappa_weights_validation(weights)
    return _cohen_kappa_reduce(binary_confusion_matrix(preds, target, threshold, None, ignore_index, validate_args), weights)


def multiclass_cohen_kappa(preds: Tensor, target: Tensor, num_classes: int,
                           weights: Optional[Literal["linear", "quadratic", "none"]] = None,
                           ignore_index: Optional[int] = None, validate_args: bool = True) -> Tensor:
    """Reference cohen_kappa.py:154-225."""
    if validate_args:
        _multiclass_confusion_matrix_arg_validation(num_classes, ignore_index)
        _cohen_kappa_weights_validation(weights)
    return _cohen_kappa_reduce(multiclass_confusion_matrix(preds, target, num_classes, None, ignore_index, validate_args), weights)


# ----------------------------------------------------------------------------------------------------------------------
# Matthews correlation coefficient (reference matthews_corrcoef.py:37-80)
# ----------------------------------------------------------------------------------------------------------------------
def _matthews_corrcoef_reduce(confmat: Tensor) -> Tensor:
    """MCC = cov(t, p) / sqrt(cov(t, t) cov(p, p)) from an un-normalised confusion matrix; multilabel matrices are summed
    into one 2x2 first.  The degenerate 2x2 cases follow the reference's rules (:46-52, :66-79); they are decided on the
    host from the four counts (one 32-byte read at `compute()` time)."""
    cm = confmat.sum(0) if confmat.ndim == 3 else confmat
    binary = cm.numel() == 4
    if binary:
        tn, fp, fn, tp = (int(v) for v in cm.reshape(-1).tolist())
        if tp + tn != 0 and fp + fn == 0:
            return torch.tensor(1.0, dtype=cm.dtype, device=cm.device)
        if tp + tn == 0 and fp + fn != 0:
            return torch.tensor(-1.0, dtype=cm.dtype, device=cm.device)
    t_k = cm.sum(dim=-1).float()
    p_k = cm.sum(dim=-2).float()
    correct = cm.diagonal().sum().float()
    n = cm.sum().float()
    # the reference accumulates these dot products with Python's `sum` over 0-d tensors (sequential fp32 adds)
    numerator = correct * n - _seq_sum(t_k * p_k)
    denom = (n**2 - _seq_sum(p_k * p_k)) * (n**2 - _seq_sum(t_k * t_k))
    if bool(denom == 0):
        if not binary:
            return torch.tensor(0, dtype=cm.dtype, device=cm.device)
        if fn == 0 and tn == 0:
            a, b = tp, fp
        elif fp == 0 and tn == 0:
            a, b = tp, fn
        elif tp == 0 and fn == 0:
            a, b = tn, fp
        else:  # tp == 0 and fp == 0
            a, b = tn, fn
        eps = torch.tensor(torch.finfo(torch.float32).eps, dtype=torch.float32, device=cm.device)
        numerator = torch.sqrt(eps) * (a - b)
        denom = (tp + fp + eps) * (tp + fn + eps) * (tn + fp + eps) * (tn + fn + eps)
    return numerator / torch.sqrt(denom)


def _seq_sum(v: Tensor) -> Tensor:
    """Left-to-right fp32 sum (what Python's builtin `sum` over a 1-d tensor does in the reference): a cumsum's last
    element adds in the same order on the device."""
    return torch.cumsum(v, dim=0)[-1]


def binary_matthews_corrcoef(preds: Tensor, target: Tensor, threshold: float = 0.5, ignore_index: Optional[int] = None,
                             validate_args: bool = True) -> Tensor:
    """Reference matthews_corrcoef.py:83-140."""
    return _matthews_corrcoef_reduce(binary_confusion_matrix(preds, target, threshold, None, ignore_index, validate_args))


def multiclass_matthews_corrcoef(preds: Tensor, target: Tensor, num_classes: int, ignore_index: Optional[int] = None,
                                 validate_args: bool = True) -> Tensor:
    """Reference matthews_corrcoef.py:143-205."""
    return _matthews_corrcoef_reduce(multiclass_confusion_matrix(preds, target, num_classes, None, ignore_index, validate_args))


def multilabel_matthews_corrcoef(preds: Tensor, target: Tensor, num_labels: int, threshold: float = 0.5,
                                 ignore_index: Optional[int] = None, validate_args: bool = True) -> Tensor:
    """Reference matthews_corrcoef.py:208-270."""
    return _matthews_corrcoef_reduce(
        multilabel_confusion_matrix(preds, target, num_labels, threshold, None, ignore_index, validate_args))


def cohen_kappa(preds: Tensor, target: Tensor, task: Literal["binary", "multiclass"], threshold: float = 0.5,
                num_classes: Optional[int] = None, weights: Optional[Literal["linear", "quadratic", "none"]] = None,
                ignore_index: Optional[int] = None, validate_args: bool = True) -> Tensor:
    """Task wrapper (reference cohen_kappa.py:228-278); binary and multiclass only."""
    from metrics_b200.functional.classification._task import call_for_task

    return call_for_task(
        task, num_classes, None,
        lambda: binary_cohen_kappa(preds, target, threshold, weights, ignore_index, validate_args),
        lambda c: multiclass_cohen_kappa(preds, target, c, weights, ignore_index, validate_args), None)


def jaccard_index(preds: Tensor, target: Tensor, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,
                  num_classes: Optional[int] = None, num_labels: Optional[int] = None,
                  average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
                  ignore_index: Optional[int] = None, validate_args: bool = True, zero_division: float = 0.0) -> Tensor:
    """Task wrapper (reference jaccard.py:348-420)."""
    from metrics_b200.functional.classification._task import call_for_task

    return call_for_task(
        task, num_classes, num_labels,
        lambda: binary_jaccard_index(preds, target, threshold, ignore_index, validate_args, zero_division),
        lambda c: multiclass_jaccard_index(preds, target, c, average, ignore_index, validate_args, zero_division),
        lambda n: multilabel_jaccard_index(preds, target, n, threshold, average, ignore_index, validate_args, zero_division))


def matthews_corrcoef(preds: Tensor, target: Tensor, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,
                      num_classes: Optional[int] = None, num_labels: Optional[int] = None,
                      ignore_index: Optional[int] = None, validate_args: bool = True) -> Tensor:
    """Task wrapper (reference matthews_corrcoef.py:273-330)."""
    from metrics_b200.functional.classification._task import call_for_task

    return call_for_task(
        task, num_classes, num_labels,
        lambda: binary_matthews_corrcoef(preds, target, threshold, ignore_index, validate_args),
        lambda c: multiclass_matthews_corrcoef(preds, target, c, ignore_index, validate_args),
        lambda n: multilabel_matthews_corrcoef(preds, target, n, threshold, ignore_index, validate_args))
