"""KL divergence (reference: functional/regression/kl_divergence.py).  The per-row measures come from ONE kernel
(`mb200_kl_divergence_rows`, csrc/kldiv.cu) instead of the reference's chain of eight ATen passes."""
from __future__ import annotations

from typing import Optional, Tuple, Union

from torch import Tensor
from typing_extensions import Literal

from metrics_b200 import _native
from metrics_b200.utilities.checks import _check_same_shape


def _kld_update(p: Tensor, q: Tensor, log_prob: bool) -> Tuple[Tensor, int]:
    """Per-observation KL divergences and the number of observations (reference :25-46).  Probabilities are normalised to
    sum 1 per row first; log-probabilities are taken as they are."""
    _check_same_shape(p, q)
    if p.ndim != 2 or q.ndim != 2:
        raise ValueError(f"Expected both p and q distribution to be 2D but got {p.ndim} and {q.ndim} respectively")
    return _native.kl_divergence_rows(p, q, bool(log_prob)), p.shape[0]


def _kld_compute(measures: Tensor, total: Union[int, Tensor], reduction: Optional[Literal["mean", "sum", "none"]] = "mean") -> Tensor:
    """Reduce over the observations (reference :49-78)."""
    if reduction == "sum":
        return measures.sum()
    if reduction == "mean":
        return measures.sum() / total
    if reduction is None or reduction == "none":
        return measures
    return measures / total


def kl_divergence(p: Tensor, q: Tensor, log_prob: bool = False,
                  reduction: Optional[Literal["mean", "sum", "none"]] = "mean") -> Tensor:
    """``D_KL(P || Q) = sum_x P(x) log(P(x) / Q(x))`` for the rows of ``p`` and ``q`` (``[N, d]``), reference :81-115."""
    measures, total = _kld_update(p, q, log_prob)
    return _kld_compute(measures, total, reduction)
