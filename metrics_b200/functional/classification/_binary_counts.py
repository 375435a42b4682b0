"""Shared front-end of the K2 counting kernel for the binary / multilabel stat-score and confusion-matrix families."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from metrics_b200 import _native
from metrics_b200.functional.classification._validation import new_flag
from metrics_b200.utilities.checks import _check_same_shape


def counts(
    preds: Tensor, target: Tensor, num_labels: int, threshold: float, ignore_index: Optional[int], samplewise: bool,
    validate_args: bool, into: Optional[Tensor] = None,
) -> Tensor:
    """``[G, 4]`` int64 ``(tp, fp, tn, fn)``; with ``validate_args`` the kernel's range flags are read back (one 4-byte
    D2H) and turned into the reference's RuntimeErrors."""
    flag = new_flag(preds.device) if validate_args else None
    if target.is_floating_point():
        # The reference has no dtype rule for binary / multilabel targets (stat_scores.py:49-88), only the value rule
        # {0, 1, ignore_index}; the kernel reads integer targets, so a float 0./1. tensor is cast once on the device.
        as_int = target.long()
        if validate_args and bool((as_int != target).any()):
            raise RuntimeError(
                f"Detected the following values in `target`: {torch.unique(target)} but expected only"
                f" the following values {[0, 1] if ignore_index is None else [ignore_index]}."
            )
        target = as_int
    out = _native.binary_stat_counts(preds, target, num_labels, threshold, ignore_index, samplewise, into, flag)
    if flag is not None:
        bits = int(flag.item())
        if bits & _native.FLAG_TARGET_RANGE:
            raise RuntimeError(
                f"Detected the following values in `target`: {torch.unique(target)} but expected only"
                f" the following values {[0, 1] if ignore_index is None else [ignore_index]}."
            )
        if bits & _native.FLAG_PREDS_RANGE:
            raise RuntimeError(
                f"Detected the following values in `preds`: {torch.unique(preds)} but expected only"
                " the following values [0,1] since `preds` is a label tensor."
            )
    return out


def check_threshold(threshold: float) -> None:
    if not (isinstance(threshold, float) and (0 <= threshold <= 1)):
        raise ValueError(f"Expected argument `threshold` to be a float in the [0,1] range, but got {threshold}.")


def check_common(multidim_average: str, ignore_index: Optional[int], zero_division: float = 0) -> None:
    allowed = ("global", "samplewise")
    if multidim_average not in allowed:
        raise ValueError(f"Expected argument `multidim_average` to be one of {allowed}, but got {multidim_average}")
    if ignore_index is not None and not isinstance(ignore_index, int):
        raise ValueError(f"Expected argument `ignore_index` to either be `None` or an integer, but got {ignore_index}")
    if zero_division not in [0, 1]:
        raise ValueError(f"Expected argument `zero_division` to be 0 or 1, but got {zero_division}.")


def binary_shape_validation(preds: Tensor, target: Tensor, multidim_average: str) -> None:
    _check_same_shape(preds, target)
    if multidim_average != "global" and preds.ndim < 2:
        raise ValueError("Expected input to be at least 2D when multidim_average is set to `samplewise`")


def multilabel_shape_validation(preds: Tensor, target: Tensor, num_labels: int, multidim_average: str) -> None:
    _check_same_shape(preds, target)
    if preds.shape[1] != num_labels:
        raise ValueError(
            "Expected both `target.shape[1]` and `preds.shape[1]` to be equal to the number of labels"
            f" but got {preds.shape[1]} and expected {num_labels}"
        )
    if multidim_average != "global" and preds.ndim < 3:
        raise ValueError("Expected input to be at least 3D when multidim_average is set to `samplewise`")
