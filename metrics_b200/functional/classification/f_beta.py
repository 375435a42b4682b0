"""F-beta / F1 reducers and functionals (reference: functional/classification/f_beta.py)."""
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


def _fbeta_reduce(
    tp: Tensor,
    fp: Tensor,
    tn: Tensor,
    fn: Tensor,
    beta: float,
    average: Optional[str],
    multidim_average: str = "global",
    multilabel: bool = False,
    zero_division: float = 0,
) -> Tensor:
    """tp/fp/fn -> F-beta (reference :37-58): ``(1+b^2) tp / ((1+b^2) tp + b^2 fn + fp)``."""
    b2 = beta**2
    if average == "micro":
        axis = 0 if multidim_average == "global" else 1
        tp, fn, fp = tp.sum(dim=axis), fn.sum(dim=axis), fp.sum(dim=axis)
    score = _safe_divide((1 + b2) * tp, (1 + b2) * tp + b2 * fn + fp, zero_division)
    if average in ("binary", "micro"):
        return score
    return _adjust_weights_safe_divide(score, average, multilabel, tp, fp, fn)


def _fbeta_arg_validation(beta: float) -> None:
    if not (isinstance(beta, float) and beta > 0):
        raise ValueError(f"Expected argument `beta` to be a float larger than 0, but got {beta}.")


def multiclass_fbeta_score(
    preds: Tensor,
    target: Tensor,
    beta: float,
    num_classes: int,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    top_k: int = 1,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
    zero_division: float = 0,
) -> Tensor:
    """Multiclass F-beta (reference :356-470)."""
    if validate_args:
        _fbeta_arg_validation(beta)
        _multiclass_stat_scores_arg_validation(num_classes, top_k, average, multidim_average, ignore_index, zero_division)
        _multiclass_stat_scores_tensor_validation(preds, target, num_classes, multidim_average, ignore_index)
    states = _multiclass_stat_scores_states(preds, target, num_classes, top_k, average, multidim_average, ignore_index, validate_args)
    return _fbeta_reduce(*states, beta, average=average, multidim_average=multidim_average, zero_division=zero_division)


def multiclass_f1_score(
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    top_k: int = 1,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
    zero_division: float = 0,
) -> Tensor:
    """Multiclass F1 = F-beta with beta 1 (reference :745-860)."""
    return multiclass_fbeta_score(
        preds, target, 1.0, num_classes, average, top_k, multidim_average, ignore_index, validate_args, zero_division
    )


# ---- binary / multilabel ---------------------------------------------------------------------------------
from metrics_b200.functional.classification.stat_scores import (  # noqa: E402
    _binary_stat_scores_arg_validation,
    _binary_stat_scores_tensor_validation,
    _binary_stat_scores_update,
    _multilabel_stat_scores_arg_validation,
    _multilabel_stat_scores_tensor_validation,
    _multilabel_stat_scores_update,
)


def binary_fbeta_score(
    preds: Tensor,
    target: Tensor,
    beta: float,
    threshold: float = 0.5,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
    zero_division: float = 0,
) -> Tensor:
    """Reference :80-160."""
    if validate_args:
        _fbeta_arg_validation(beta)
        _binary_stat_scores_arg_validation(threshold, multidim_average, ignore_index, zero_division)
        _binary_stat_scores_tensor_validation(preds, target, multidim_average, ignore_index)
    tp, fp, tn, fn = _binary_stat_scores_update(preds, target, threshold, multidim_average, ignore_index, validate_args)
    return _fbeta_reduce(tp, fp, tn, fn, beta, average="binary", multidim_average=multidim_average, zero_division=zero_division)


def binary_f1_score(
    preds: Tensor,
    target: Tensor,
    threshold: float = 0.5,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
    zero_division: float = 0,
) -> Tensor:
    return binary_fbeta_score(preds, target, 1.0, threshold, multidim_average, ignore_index, validate_args, zero_division)


def multilabel_fbeta_score(
    preds: Tensor,
    target: Tensor,
    beta: float,
    num_labels: int,
    threshold: float = 0.5,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
    zero_division: float = 0,
) -> Tensor:
    """Reference :290-400."""
    if validate_args:
        _fbeta_arg_validation(beta)
        _multilabel_stat_scores_arg_validation(num_labels, threshold, average, multidim_average, ignore_index, zero_division)
        _multilabel_stat_scores_tensor_validation(preds, target, num_labels, multidim_average, ignore_index)
    tp, fp, tn, fn = _multilabel_stat_scores_update(
        preds, target, num_labels, threshold, multidim_average, ignore_index, validate_args
    )
    return _fbeta_reduce(
        tp, fp, tn, fn, beta, average=average, multidim_average=multidim_average, multilabel=True, zero_division=zero_division
    )


def multilabel_f1_score(
    preds: Tensor,
    target: Tensor,
    num_labels: int,
    threshold: float = 0.5,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
    zero_division: float = 0,
) -> Tensor:
    return multilabel_fbeta_score(
        preds, target, 1.0, num_labels, threshold, average, multidim_average, ignore_index, validate_args, zero_division
    )


def fbeta_score(preds: Tensor, target: Tensor, task: Literal["binary", "multiclass", "multilabel"], beta: float = 1.0,
                threshold: float = 0.5, num_classes: Optional[int] = None, num_labels: Optional[int] = None,
                average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
                multidim_average: Optional[Literal["global", "samplewise"]] = "global", top_k: Optional[int] = 1,
                ignore_index: Optional[int] = None, validate_args: bool = True, zero_division: float = 0) -> Tensor:
    """Task wrapper (reference :636-707)."""
    from metrics_b200.functional.classification._task import call_for_task

    def mc(c: int) -> Tensor:
        if not isinstance(top_k, int):
            raise ValueError(f"`top_k` is expected to be `int` but `{type(top_k)} was passed.`")
        return multiclass_fbeta_score(preds, target, beta, c, average, top_k, multidim_average, ignore_index, validate_args,
                                      zero_division)

    return call_for_task(
        task, num_classes, num_labels,
        lambda: binary_fbeta_score(preds, target, beta, threshold, multidim_average, ignore_index, validate_args, zero_division),
        mc,
        lambda n: multilabel_fbeta_score(preds, target, beta, n, threshold, average, multidim_average, ignore_index,
                                         validate_args, zero_division))


def f1_score(preds: Tensor, target: Tensor, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,
             num_classes: Optional[int] = None, num_labels: Optional[int] = None,
             average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
             multidim_average: Optional[Literal["global", "samplewise"]] = "global", top_k: Optional[int] = 1,
             ignore_index: Optional[int] = None, validate_args: bool = True, zero_division: float = 0) -> Tensor:
    """Task wrapper (reference :710-780): F-beta with beta = 1."""
    return fbeta_score(preds, target, task, 1.0, threshold, num_classes, num_labels, average, multidim_average, top_k,
                       ignore_index, validate_args, zero_division)
