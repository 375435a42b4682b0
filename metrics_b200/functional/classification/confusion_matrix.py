"""Confusion-matrix functionals (reference: functional/classification/confusion_matrix.py).

The reference's seam is `_x_format` (argmax + flatten + ignore drop) -> `_x_update` (bincount) -> `_x_compute`
(normalise).  Here format+update are ONE kernel (`mb200_multiclass_confmat_update`, csrc/confmat.cu) that adds
straight into a `[C, C]` int64 tensor; the seam functions are kept for callers that use them piecewise.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200 import _native
from metrics_b200.functional.classification._validation import check_multiclass_shapes, labels_as_int, new_flag, raise_if_flagged
from metrics_b200.utilities.prints import rank_zero_warn

_NORMALIZE = ("true", "pred", "all", "none", None)


def _confusion_matrix_reduce(confmat: Tensor, normalize: Optional[str] = None) -> Tensor:
    """Optional normalisation over targets / predictions / everything (reference :27-60); NaN -> 0 with a warning."""
    if normalize not in _NORMALIZE:
        raise ValueError(f"Argument `normalize` needs to one of the following: {_NORMALIZE}")
    if normalize is None or normalize == "none":
        return confmat
    cm = confmat if confmat.is_floating_point() else confmat.float()
    if normalize == "true":
        cm = cm / cm.sum(dim=-1, keepdim=True)
    elif normalize == "pred":
        cm = cm / cm.sum(dim=-2, keepdim=True)
    else:
        cm = cm / cm.sum(dim=[-2, -1], keepdim=True)
    nan_mask = torch.isnan(cm)
    n_nan = int(nan_mask.sum())
    if n_nan:
        cm = cm.masked_fill(nan_mask, 0.0)
        rank_zero_warn(f"{n_nan} NaN values found in confusion matrix have been replaced with zeros.")
    return cm


# ---------------------------------------------------------------------------------------------------------
# multiclass
# ---------------------------------------------------------------------------------------------------------
def _multiclass_confusion_matrix_arg_validation(
    num_classes: int, ignore_index: Optional[int] = None, normalize: Optional[str] = None
) -> None:
    if not isinstance(num_classes, int) or num_classes < 2:
        raise ValueError(f"Expected argument `num_classes` to be an integer larger than 1, but got {num_classes}")
    if ignore_index is not None and not isinstance(ignore_index, int):
        raise ValueError(f"Expected argument `ignore_index` to either be `None` or an integer, but got {ignore_index}")
    if normalize not in _NORMALIZE:
        raise ValueError(f"Expected argument `normalize` to be one of {_NORMALIZE}, but got {normalize}.")


def _multiclass_confusion_matrix_tensor_validation(
    preds: Tensor, target: Tensor, num_classes: int, ignore_index: Optional[int] = None
) -> None:
    """Host-side shape rules (reference :250-286).  Label *values* are checked inside the update kernel."""
    check_multiclass_shapes(preds, target, num_classes)


def _multiclass_confusion_matrix_update_(
    confmat: Tensor,
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    ignore_index: Optional[int] = None,
    validate_args: bool = False,
) -> None:
    """FUSED format+update: ``confmat[target, argmax(preds)] += 1`` in place, one pass over ``preds``.

    ``validate_args=True``: the label range check runs INSIDE the kernel, so the batch is first counted into a scratch matrix
    and folded into the state only after the error word came back clean — a caller that catches the error finds the state
    untouched, like the reference, which validates before it counts (confusion_matrix.py:287-294)."""
    if not validate_args:
        _native.multiclass_confmat_update_(confmat, labels_as_int(preds, target), target, num_classes, ignore_index, None)
        return
    flag = new_flag(confmat.device)
    scratch = torch.zeros_like(confmat)
    _native.multiclass_confmat_update_(scratch, labels_as_int(preds, target), target, num_classes, ignore_index, flag)
    raise_if_flagged(flag, num_classes, ignore_index)
    confmat += scratch


def _multiclass_confusion_matrix_format(
    preds: Tensor, target: Tensor, ignore_index: Optional[int] = None, convert_to_labels: bool = True
) -> tuple[Tensor, Tensor]:
    """Piecewise seam (reference :297-321): argmax kernel + flatten + boolean drop of ignored rows."""
    if preds.ndim == target.ndim + 1 and convert_to_labels:
        preds = _native.argmax_rows(preds)
    preds = preds.flatten() if convert_to_labels else torch.movedim(preds, 1, -1).reshape(-1, preds.shape[1])
    target = target.flatten()
    if ignore_index is not None:
        keep = target != ignore_index
        preds, target = preds[keep], target[keep]
    return preds, target


def _multiclass_confusion_matrix_update(preds: Tensor, target: Tensor, num_classes: int) -> Tensor:
    """Piecewise seam (reference :324-328): label preds + label target -> fresh ``[C, C]`` int64 counts."""
    confmat = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=preds.device)
    _native.multiclass_confmat_update_(confmat, labels_as_int(preds, target), target, num_classes, None, None)
    return confmat


def _multiclass_confusion_matrix_compute(confmat: Tensor, normalize: Optional[str] = None) -> Tensor:
    return _confusion_matrix_reduce(confmat, normalize)


def multiclass_confusion_matrix(
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    normalize: Optional[Literal["true", "pred", "all", "none"]] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """``[num_classes, num_classes]`` confusion matrix (rows = target, cols = prediction); reference :333-400.

    ``preds``: ``(N, ...)`` integer labels or ``(N, C, ...)`` floating scores (argmax over ``C``);
    ``target``: ``(N, ...)`` integer labels.  Inputs must be CUDA tensors.
    """
    if validate_args:
        _multiclass_confusion_matrix_arg_validation(num_classes, ignore_index, normalize)
        _multiclass_confusion_matrix_tensor_validation(preds, target, num_classes, ignore_index)
    confmat = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=preds.device)
    _multiclass_confusion_matrix_update_(confmat, preds, target, num_classes, ignore_index, validate_args)
    return _multiclass_confusion_matrix_compute(confmat, normalize)


# =========================================================================================================
# binary / multilabel (kernel K2): the 2x2 matrix is [[tn, fp], [fn, tp]]
# =========================================================================================================
from metrics_b200.functional.classification import _binary_counts as _bc  # noqa: E402


def _counts_to_confmat(c: Tensor) -> Tensor:
    tp, fp, tn, fn = c.unbind(-1)
    return torch.stack([torch.stack([tn, fp], -1), torch.stack([fn, tp], -1)], -2)


def _binary_confusion_matrix_arg_validation(
    threshold: float = 0.5, ignore_index: Optional[int] = None, normalize: Optional[str] = None
) -> None:
    if not (isinstance(threshold, float) and (0 <= threshold <= 1)):
        raise ValueError(f"Expected argument `threshold` to be a float in the [0,1] range, but got {threshold}.")
    if ignore_index is not None and not isinstance(ignore_index, int):
        raise ValueError(f"Expected argument `ignore_index` to either be `None` or an integer, but got {ignore_index}")
    if normalize not in _NORMALIZE:
        raise ValueError(f"Expected argument `normalize` to be one of {_NORMALIZE}, but got {normalize}.")


def _binary_confusion_matrix_tensor_validation(preds: Tensor, target: Tensor, ignore_index: Optional[int] = None) -> None:
    _bc.binary_shape_validation(preds, target, "global")


def _binary_confusion_matrix_update(
    preds: Tensor, target: Tensor, threshold: float = 0.5, ignore_index: Optional[int] = None, validate_args: bool = False
) -> Tensor:
    """FUSED format+update (reference :119-152): ``[2, 2]`` int64 counts of this batch."""
    return _counts_to_confmat(_bc.counts(preds, target, 1, threshold, ignore_index, False, validate_args))[0]


def _binary_confusion_matrix_compute(confmat: Tensor, normalize: Optional[str] = None) -> Tensor:
    return _confusion_matrix_reduce(confmat, normalize)


def binary_confusion_matrix(
    preds: Tensor,
    target: Tensor,
    threshold: float = 0.5,
    normalize: Optional[Literal["true", "pred", "all", "none"]] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """``[[tn, fp], [fn, tp]]`` (reference :164-228)."""
    if validate_args:
        _binary_confusion_matrix_arg_validation(threshold, ignore_index, normalize)
        _binary_confusion_matrix_tensor_validation(preds, target, ignore_index)
    confmat = _binary_confusion_matrix_update(preds, target, threshold, ignore_index, validate_args)
    return _binary_confusion_matrix_compute(confmat, normalize)


def _multilabel_confusion_matrix_arg_validation(
    num_labels: int, threshold: float = 0.5, ignore_index: Optional[int] = None, normalize: Optional[str] = None
) -> None:
    if not isinstance(num_labels, int) or num_labels < 2:
        raise ValueError(f"Expected argument `num_labels` to be an integer larger than 1, but got {num_labels}")
    _binary_confusion_matrix_arg_validation(threshold, ignore_index, normalize)


def _multilabel_confusion_matrix_tensor_validation(
    preds: Tensor, target: Tensor, num_labels: int, ignore_index: Optional[int] = None
) -> None:
    _bc.multilabel_shape_validation(preds, target, num_labels, "global")


def _multilabel_confusion_matrix_update(
    preds: Tensor, target: Tensor, num_labels: int, threshold: float = 0.5, ignore_index: Optional[int] = None,
    validate_args: bool = False,
) -> Tensor:
    """FUSED format+update (reference :477-516): ``[L, 2, 2]`` int64 counts of this batch."""
    return _counts_to_confmat(_bc.counts(preds, target, num_labels, threshold, ignore_index, False, validate_args))


def _multilabel_confusion_matrix_compute(confmat: Tensor, normalize: Optional[str] = None) -> Tensor:
    return _confusion_matrix_reduce(confmat, normalize)


def multilabel_confusion_matrix(
    preds: Tensor,
    target: Tensor,
    num_labels: int,
    threshold: float = 0.5,
    normalize: Optional[Literal["true", "pred", "all", "none"]] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Per-label ``[[tn, fp], [fn, tp]]`` (reference :527-600)."""
    if validate_args:
        _multilabel_confusion_matrix_arg_validation(num_labels, threshold, ignore_index, normalize)
        _multilabel_confusion_matrix_tensor_validation(preds, target, num_labels, ignore_index)
    confmat = _multilabel_confusion_matrix_update(preds, target, num_labels, threshold, ignore_index, validate_args)
    return _multilabel_confusion_matrix_compute(confmat, normalize)


def confusion_matrix(
    preds: Tensor,
    target: Tensor,
    task: Literal["binary", "multiclass", "multilabel"],
    threshold: float = 0.5,
    num_classes: Optional[int] = None,
    num_labels: Optional[int] = None,
    normalize: Optional[str] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Task dispatcher (reference :603-655)."""
    from metrics_b200.utilities.enums import ClassificationTask

    task = ClassificationTask.from_str(task)
    if task == ClassificationTask.BINARY:
        return binary_confusion_matrix(preds, target, threshold, normalize, ignore_index, validate_args)
    if task == ClassificationTask.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        return multiclass_confusion_matrix(preds, target, num_classes, normalize, ignore_index, validate_args)
    if not isinstance(num_labels, int):
        raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
    return multilabel_confusion_matrix(preds, target, num_labels, threshold, normalize, ignore_index, validate_args)
