"""Modular group-fairness metrics (reference: classification/group_fairness.py:36-326): four ``[num_groups]`` int64
``sum`` states filled by one counting-kernel launch per ``update`` (functional/classification/group_fairness.py)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.group_fairness import (
    _check_task,
    _compute_binary_demographic_parity,
    _compute_binary_equal_opportunity,
    _group_counts,
)
from metrics_b200.functional.classification.stat_scores import _binary_stat_scores_arg_validation
from metrics_b200.metric import Metric
from metrics_b200.utilities.prints import rank_zero_warn


class _AbstractGroupStatScores(Metric):
    """tp / fp / tn / fn per group id (reference :36-57)."""

    is_differentiable: bool = False
    higher_is_better: bool = False
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0

    def _create_states(self, num_groups: int) -> None:
        for name in ("tp", "fp", "tn", "fn"):
            self.add_state(name, torch.zeros(num_groups, dtype=torch.long), dist_reduce_fx="sum")

    def _setup(self, num_groups: int, threshold: float, ignore_index: Optional[int], validate_args: bool) -> None:
        if validate_args:
            _binary_stat_scores_arg_validation(threshold, "global", ignore_index)
        if not isinstance(num_groups, int) or num_groups < 2:
            raise ValueError(f"Expected argument `num_groups` to be an int larger than 1, but got {num_groups}")
        self.num_groups = num_groups
        self.threshold = threshold
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self._create_states(num_groups)

    def _accumulate(self, preds: Tensor, target: Tensor, groups: Tensor) -> None:
        counts = _group_counts(preds, target, groups, self.num_groups, self.threshold, self.ignore_index, self.validate_args)
        self.tp += counts[:, 0]
        self.fp += counts[:, 1]
        self.tn += counts[:, 2]
        self.fn += counts[:, 3]


class BinaryGroupStatRates(_AbstractGroupStatScores):
    """``{"group_i": [tp, fp, tn, fn] / total}`` per group (reference :60-155)."""

    def __init__(self, num_groups: int, threshold: float = 0.5, ignore_index: Optional[int] = None,
                 validate_args: bool = True, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self._setup(num_groups, threshold, ignore_index, validate_args)

    def update(self, preds: Tensor, target: Tensor, groups: Tensor) -> None:
        self._accumulate(preds, target, groups)

    def compute(self) -> Dict[str, Tensor]:
        table = torch.stack((self.tp, self.fp, self.tn, self.fn), dim=1)
        return {f"group_{i}": row / row.sum() for i, row in enumerate(table)}


class BinaryFairness(_AbstractGroupStatScores):
    """Demographic parity and / or equal opportunity between the groups (reference :158-284)."""

    def __init__(self, num_groups: int, task: Literal["demographic_parity", "equal_opportunity", "all"] = "all",
                 threshold: float = 0.5, ignore_index: Optional[int] = None, validate_args: bool = True,
                 **kwargs: Any) -> None:
        super().__init__(**kwargs)
        _check_task(task)
        self.task = task
        self._setup(num_groups, threshold, ignore_index, validate_args)

    def update(self, preds: Tensor, target: Tensor, groups: Tensor) -> None:
        if self.task == "demographic_parity":
            if target is not None:
                rank_zero_warn("The task demographic_parity does not require a target.", UserWarning)
            target = torch.zeros(preds.shape, dtype=torch.long, device=preds.device)
        self._accumulate(preds, target, groups)

    def compute(self) -> Dict[str, Tensor]:
        out: Dict[str, Tensor] = {}
        if self.task in ("demographic_parity", "all"):
            out.update(_compute_binary_demographic_parity(self.tp, self.fp, self.tn, self.fn))
        if self.task in ("equal_opportunity", "all"):
            out.update(_compute_binary_equal_opportunity(self.tp, self.fp, self.tn, self.fn))
        return out
