"""Small floating-point helpers shared by the `_compute` halves (reference: utilities/compute.py).

These run on tiny `[C]`-sized tensors after the kernels have produced the integer states; they define the
rounding of the final ratios (int64 -> float32, then divide), so they follow the reference's op order exactly.
"""
from __future__ import annotations

from typing import Optional, Union

import torch
from torch import Tensor


def _safe_divide(num: Tensor, denom: Tensor, zero_division: float = 0.0) -> Tensor:
    """``num / denom`` in floating point, ``zero_division`` where ``denom == 0`` (compute.py:47-68).

    Integer inputs are widened with ``.float()`` first — this is what fixes the rounding of every ratio metric
    (int64 counts -> f32 -> one IEEE division).
    """
    if not num.is_floating_point():
        num = num.float()
    if not denom.is_floating_point():
        denom = denom.float()
    fallback = torch.full((), zero_division, dtype=num.dtype, device=num.device)
    return torch.where(denom != 0, num / denom, fallback)


def _adjust_weights_safe_divide(
    score: Tensor, average: Optional[str], multilabel: bool, tp: Tensor, fp: Tensor, fn: Tensor, top_k: int = 1
) -> Tensor:
    """Class averaging of a per-class score (compute.py:71-82).

    ``weighted`` uses the support ``tp + fn``; ``macro`` weighs every class 1 except (multiclass only) classes
    that never occur in preds or target, which get weight 0.
    """
    if average in (None, "none"):
        return score
    if average == "weighted":
        w = tp + fn
    else:
        w = torch.ones_like(score)
        if not multilabel:
            absent = (tp + fp + fn == 0) if top_k == 1 else (tp + fn == 0)
            w = w.masked_fill(absent, 0.0)
    return _safe_divide(w * score, w.sum(-1, keepdim=True)).sum(-1)


def _auc_compute_without_check(x: Tensor, y: Tensor, direction: float, axis: int = -1) -> Tensor:
    """Trapezoidal area (compute.py:101-109)."""
    with torch.no_grad():
        return torch.trapz(y, x, dim=axis) * direction


def interp(x: Tensor, xp: Tensor, fp: Tensor) -> Tensor:
    """1-D piecewise-linear interpolation, numpy.interp-like (compute.py:157-187)."""
    m = _safe_divide(fp[1:] - fp[:-1], xp[1:] - xp[:-1])  # repeated sample points: slope 0, like the reference (:181)
    b = fp[:-1] - m * xp[:-1]
    idx = torch.sum(torch.ge(x[:, None], xp[None, :]), 1) - 1
    idx = torch.clamp(idx, 0, len(m) - 1)
    return m[idx] * x + b[idx]
