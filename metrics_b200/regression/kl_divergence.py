"""KLDivergence (reference: regression/kl_divergence.py:31-125).  States as in the reference — `measures` (a float32 running
sum for the "mean" / "sum" reductions, a list of per-observation tensors for "none") and `total` (observation count, "sum") —
fed by the row kernel K13 through `_kld_update`."""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.regression.kl_divergence import _kld_compute, _kld_update
from metrics_b200.metric import Metric
from metrics_b200.utilities.data import dim_zero_cat

_REDUCTIONS = ["mean", "sum", "none", None]


class KLDivergence(Metric):
    """Running KL divergence ``D_KL(P || Q)`` over rows of ``p`` and ``q`` (``[N, d]``, probabilities or log-probabilities)."""

    is_differentiable: bool = False  # kernel launches carry no autograd graph (reference: True)
    higher_is_better: bool = False
    full_state_update: bool = False
    plot_lower_bound: float = 0.0

    def __init__(self, log_prob: bool = False, reduction: Optional[Literal["mean", "sum", "none"]] = "mean", **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if not isinstance(log_prob, bool):
            raise TypeError(f"Expected argument `log_prob` to be bool but got {log_prob}")
        if reduction not in _REDUCTIONS:
            raise ValueError(f"Expected argument `reduction` to be one of {_REDUCTIONS} but got {reduction}")
        self.log_prob, self.reduction = log_prob, reduction
        self._keeps_rows = reduction in ("none", None)
        if self._keeps_rows:
            self.add_state("measures", [], dist_reduce_fx="cat")
        else:
            self.add_state("measures", torch.tensor(0.0), dist_reduce_fx="sum")
        self.add_state("total", torch.tensor(0), dist_reduce_fx="sum")

    def update(self, p: Tensor, q: Tensor) -> None:
        rows, count = _kld_update(p, q, self.log_prob)
        if self._keeps_rows:  # (the reference leaves `total` untouched in this mode too, :116-120; compute does not use it)
            self.measures.append(rows)
            return
        self.measures += rows.sum()
        self.total += count

    def compute(self) -> Tensor:
        return _kld_compute(dim_zero_cat(self.measures) if self._keeps_rows else self.measures, self.total, self.reduction)
