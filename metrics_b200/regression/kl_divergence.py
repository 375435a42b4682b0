"""KLDivergence (reference: regression/kl_divergence.py)."""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.regression.kl_divergence import _kld_compute, _kld_update
from metrics_b200.metric import Metric
from metrics_b200.utilities.data import dim_zero_cat


class KLDivergence(Metric):
    """Running KL divergence of row distributions (reference :31-125): ``measures`` is a running sum for the "mean" / "sum"
    reductions and a list of per-observation values for "none"; ``total`` counts the observations."""

    is_differentiable: bool = False  # kernel launches carry no autograd graph (reference: True)
    higher_is_better: bool = False
    full_state_update: bool = False
    plot_lower_bound: float = 0.0

    def __init__(self, log_prob: bool = False, reduction: Optional[Literal["mean", "sum", "none"]] = "mean", **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if not isinstance(log_prob, bool):
            raise TypeError(f"Expected argument `log_prob` to be bool but got {log_prob}")
        self.log_prob = log_prob
        allowed_reduction = ["mean", "sum", "none", None]
        if reduction not in allowed_reduction:
            raise ValueError(f"Expected argument `reduction` to be one of {allowed_reduction} but got {reduction}")
        self.reduction = reduction
        if self.reduction in ["mean", "sum"]:
            self.add_state("measures", torch.tensor(0.0), dist_reduce_fx="sum")
        else:
            self.add_state("measures", [], dist_reduce_fx="cat")
        self.add_state("total", torch.tensor(0), dist_reduce_fx="sum")

    def update(self, p: Tensor, q: Tensor) -> None:
        measures, total = _kld_update(p, q, self.log_prob)
        if self.reduction is None or self.reduction == "none":
            self.measures.append(measures)  # (the reference does not count `total` in this mode either, :116-120)
        else:
            self.measures += measures.sum()
            self.total += total

    def compute(self) -> Tensor:
        measures: Tensor = dim_zero_cat(self.measures) if self.reduction in ["none", None] else self.measures
        return _kld_compute(measures, self.total, self.reduction)
