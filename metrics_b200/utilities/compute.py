"""Small floating-point helpers shared by the `_compute` halves (reference: utilities/compute.py).

These run on tiny `[C]`-sized tensors after the kernels have produced the integer states; they define the
rounding of the final ratios (int64 -> float32, then divide), so they follow the reference's op order exactly.
"""
from __future__ import annotations

from typing import Optional

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


def _safe_matmul(x: Tensor, y: Tensor) -> Tensor:
    """``x @ y.T``; half inputs are multiplied in float32 and rounded back (reference compute.py:21-29)."""
    if torch.float16 in (x.dtype, y.dtype):
        return (x.float() @ y.float().T).half()
    return x @ y.T


def _safe_xlogy(x: Tensor, y: Tensor) -> Tensor:
    """``x * log(y)`` with the convention ``0 * log(anything) = 0`` (reference compute.py:32-44)."""
    return torch.where(x == 0, torch.zeros_like(x), x * torch.log(y))


def _auc_format_inputs(x: Tensor, y: Tensor) -> tuple[Tensor, Tensor]:
    """Squeeze to 1-d and check the two lengths (reference compute.py:85-98)."""
    x = x.squeeze() if x.ndim > 1 else x
    y = y.squeeze() if y.ndim > 1 else y
    if x.ndim > 1 or y.ndim > 1:
        raise ValueError(f"Expected both `x` and `y` tensor to be 1d, but got tensors with dimension {x.ndim} and {y.ndim}")
    if x.numel() != y.numel():
        raise ValueError(f"Expected the same number of elements in `x` and `y` tensor but received {x.numel()} and {y.numel()}")
    return x, y


def _auc_compute(x: Tensor, y: Tensor, reorder: bool = False) -> Tensor:
    """Trapezoidal area under ``y(x)`` for monotone ``x`` (either direction); ``reorder`` sorts first (reference :112-138)."""
    with torch.no_grad():
        if reorder:
            x, order = torch.sort(x, stable=True)
            y = y[order]
        dx = x[1:] - x[:-1]
        direction = 1.0
        if bool((dx < 0).any()):
            if not bool((dx <= 0).all()):
                raise ValueError(
                    "The `x` tensor is neither increasing or decreasing. Try setting the reorder argument to `True`."
                )
            direction = -1.0
        return _auc_compute_without_check(x, y, direction)


def auc(x: Tensor, y: Tensor, reorder: bool = False) -> Tensor:
    """Area under the curve by the trapezoidal rule (reference :141-154)."""
    return _auc_compute(*_auc_format_inputs(x, y), reorder=reorder)


def normalize_logits_if_needed(tensor: Tensor, normalization: str) -> Tensor:
    """Sigmoid / softmax (over dim 1) when any value lies outside [0, 1] — kernel K6 with its device-side vote, no host
    sync (reference :190-230).  CUDA tensors only, like every kernel of this package."""
    from metrics_b200 import _native

    if normalization not in ("sigmoid", "softmax"):
        raise ValueError(f"Expected `normalization` to be 'sigmoid' or 'softmax' but got {normalization}")
    return _native.sigmoid_if_logits(tensor) if normalization == "sigmoid" else _native.softmax_if_logits(tensor)
