"""Validation shared by the multiclass front-ends.

The reference checks label *content* with ``len(torch.unique(x)) > num_classes`` (a device sort + host sync per
tensor: functional/classification/confusion_matrix.py:287-294, stat_scores.py:317-325).  Here the update kernels
range-check every label they consume and OR a bit into a 4-byte device flag; ``validate_args=True`` costs one
4-byte read-back after the update instead of two sorts before it.  Out-of-range labels always imply what the
reference tests for on valid label sets (more than ``num_classes`` distinct values), and additionally catch the
inputs the reference silently mis-counts.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from metrics_b200 import _native


def new_flag(device: torch.device) -> Tensor:
    return torch.zeros(1, dtype=torch.int32, device=device)


def raise_if_flagged(flag: Tensor, num_classes: int, ignore_index: Optional[int]) -> None:
    """Read the kernel's error word (host sync) and raise the reference's RuntimeError when it is set."""
    bits = int(flag.item())
    if bits == 0:
        return
    flag.zero_()
    expected = num_classes if ignore_index is None else num_classes + 1
    if bits & _native.FLAG_SPIN_TIMEOUT:
        raise RuntimeError("metrics_b200: internal kernel wait timed out; results are invalid")
    name = "target" if bits & _native.FLAG_TARGET_RANGE else "preds"
    raise RuntimeError(
        f"Detected more unique values in `{name}` than expected. Expected only {expected} but found"
        f" values outside of [0, {num_classes}) in `{name}`."
    )


def check_multiclass_shapes(preds: Tensor, target: Tensor, num_classes: Optional[int]) -> None:
    """Shape rules common to multiclass confusion-matrix and stat-scores inputs."""
    if preds.ndim == target.ndim + 1:
        if not preds.is_floating_point():
            raise ValueError("If `preds` have one dimension more than `target`, `preds` should be a float tensor.")
        if num_classes is not None and preds.shape[1] != num_classes:
            raise ValueError(
                "If `preds` have one dimension more than `target`, `preds.shape[1]` should be"
                " equal to number of classes."
            )
        if preds.shape[2:] != target.shape[1:]:
            raise ValueError(
                "If `preds` have one dimension more than `target`, the shape of `preds` should be"
                " (N, C, ...), and the shape of `target` should be (N, ...)."
            )
    elif preds.ndim == target.ndim:
        if preds.shape != target.shape:
            raise ValueError(
                "The `preds` and `target` should have the same shape,"
                f" got `preds` with shape={preds.shape} and `target` with shape={target.shape}."
            )
    else:
        raise ValueError(
            "Either `preds` and `target` both should have the (same) shape (N, ...), or `target` should be (N, ...)"
            " and `preds` should be (N, C, ...)."
        )


def labels_as_int(preds: Tensor, target: Tensor) -> Tensor:
    """Label-format predictions (same shape as ``target``) in a floating dtype: the reference has no dtype rule for them
    and truncates with ``.long()`` when it builds the bincount index (functional/classification/stat_scores.py:442,
    confusion_matrix.py:323); the kernels read integer labels, so do that cast once here."""
    if preds.ndim == target.ndim and preds.is_floating_point():
        return preds.long()
    return preds
