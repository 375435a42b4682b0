"""Log-AUC: area under the ROC curve over a logarithmic false-positive-rate window, rescaled to [0, 1].

Reference: functional/classification/logauc.py.  The ROC curve comes from the sort + scan kernels (exact) or the K4 state
(binned); the window integration below is a few device ops on the curve arrays with no host synchronisation (the reference
locates the window ends with `torch.where(...)[0][-1]` on float equality of log10 values computed on two devices).
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.roc import binary_roc, multiclass_roc, multilabel_roc
from metrics_b200.utilities.compute import _safe_divide
from metrics_b200.utilities.data import interp
from metrics_b200.utilities.prints import rank_zero_warn


def _validate_fpr_range(fpr_range: Tuple[float, float]) -> None:
    if not isinstance(fpr_range, tuple) and not len(fpr_range) == 2:
        raise ValueError(f"The `fpr_range` should be a tuple of two floats, but got {type(fpr_range)}.")
    if not (0 <= fpr_range[0] < fpr_range[1] <= 1):
        raise ValueError(f"The `fpr_range` should be a tuple of two floats in the range [0, 1], but got {fpr_range}.")


def _binary_logauc_compute(fpr: Tensor, tpr: Tensor, fpr_range: Tuple[float, float] = (0.001, 0.1)) -> Tensor:
    """Trapezoid area of tpr over log10(fpr) between the two window ends, divided by the window's log width
    (reference :35-61).  The window ends are inserted into the curve (tpr by linear interpolation) first."""
    if fpr.numel() < 2 or tpr.numel() < 2:
        rank_zero_warn("At least two values on for the fpr and tpr are required to compute the log AUC. Returns 0 score.")
        return torch.tensor(0.0, device=fpr.device)
    ends = torch.tensor(fpr_range, dtype=fpr.dtype, device=fpr.device)
    y = torch.cat([tpr, interp(ends, fpr, tpr)]).sort().values
    x = torch.cat([fpr, ends]).sort().values
    # last occurrence of each window end in the merged, sorted fpr axis
    lower = torch.searchsorted(x, ends[0:1], right=True) - 1
    upper = torch.searchsorted(x, ends[1:2], right=True) - 1
    lx = torch.log10(x)
    seg = (lx[1:] - lx[:-1]) * (y[1:] + y[:-1]) * 0.5  # may hold nan / inf outside the window (log10(0)); masked below
    pos = torch.arange(seg.numel(), device=x.device)
    inside = (pos >= lower) & (pos < upper)
    area = torch.where(inside, seg, torch.zeros_like(seg)).sum()
    width = torch.log10(ends[1]) - torch.log10(ends[0])
    return area / width


def _reduce_logauc(
    fpr: Union[Tensor, List[Tensor]],
    tpr: Union[Tensor, List[Tensor]],
    fpr_range: Tuple[float, float] = (0.001, 0.1),
    average: Optional[Literal["macro", "weighted", "none"]] = "macro",
    weights: Optional[Tensor] = None,
) -> Tensor:
    """Per-curve scores, then none / macro / weighted averaging over the non-NaN ones (reference :64-88)."""
    scores = torch.stack([_binary_logauc_compute(f, t, fpr_range) for f, t in zip(fpr, tpr)])
    nan = torch.isnan(scores)
    if bool(nan.any()):
        rank_zero_warn(
            f"LogAUC score for one or more classes/labels was `nan`. Ignoring these classes in {average}-average."
        )
    if average is None or average == "none":
        return scores
    keep = ~nan
    if average == "macro":
        return scores[keep].mean()
    if average == "weighted" and weights is not None:
        w = _safe_divide(weights[keep], weights[keep].sum())
        return (scores[keep] * w).sum()
    raise ValueError(f"Got unknown average parameter: {average}. Please choose one of ['macro', 'weighted', 'none'].")


def binary_logauc(
    preds: Tensor,
    target: Tensor,
    fpr_range: Tuple[float, float] = (0.001, 0.1),
    thresholds: Optional[Union[int, List[float], Tensor]] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Reference :91-157."""
    _validate_fpr_range(fpr_range)
    fpr, tpr, _ = binary_roc(preds, target, thresholds, ignore_index, validate_args)
    return _binary_logauc_compute(fpr, tpr, fpr_range)


def multiclass_logauc(
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    fpr_range: Tuple[float, float] = (0.001, 0.1),
    average: Optional[Literal["macro", "none"]] = "macro",
    thresholds: Optional[Union[int, List[float], Tensor]] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Reference :160-239."""
    if validate_args:
        _validate_fpr_range(fpr_range)
    fpr, tpr, _ = multiclass_roc(preds, target, num_classes, thresholds, average=None, ignore_index=ignore_index,
                                 validate_args=validate_args)
    return _reduce_logauc(fpr, tpr, fpr_range, average)


def multilabel_logauc(
    preds: Tensor,
    target: Tensor,
    num_labels: int,
    fpr_range: Tuple[float, float] = (0.001, 0.1),
    average: Optional[Literal["macro", "none"]] = "macro",
    thresholds: Optional[Union[int, List[float], Tensor]] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Reference :242-320."""
    fpr, tpr, _ = multilabel_roc(preds, target, num_labels, thresholds, ignore_index, validate_args)
    return _reduce_logauc(fpr, tpr, fpr_range, average=average)


def logauc(
    preds: Tensor,
    target: Tensor,
    task: Literal["binary", "multiclass", "multilabel"],
    thresholds: Optional[Union[int, List[float], Tensor]] = None,
    num_classes: Optional[int] = None,
    num_labels: Optional[int] = None,
    fpr_range: Tuple[float, float] = (0.001, 0.1),
    average: Optional[Literal["macro", "none"]] = None,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Task wrapper (reference :323-356)."""
    from metrics_b200.utilities.enums import ClassificationTask

    task_ = ClassificationTask.from_str(task)
    if task_ == ClassificationTask.BINARY:
        return binary_logauc(preds, target, fpr_range, thresholds, ignore_index, validate_args)
    if task_ == ClassificationTask.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        return multiclass_logauc(preds, target, num_classes, fpr_range, average, thresholds, ignore_index, validate_args)
    if not isinstance(num_labels, int):
        raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
    return multilabel_logauc(preds, target, num_labels, fpr_range, average, thresholds, ignore_index, validate_args)
