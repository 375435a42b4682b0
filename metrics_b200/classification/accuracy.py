"""Accuracy metric classes (reference: classification/accuracy.py)."""
from __future__ import annotations

from typing import Optional

from torch import Tensor

from metrics_b200.classification.stat_scores import MulticlassStatScores
from metrics_b200.functional.classification.accuracy import _accuracy_reduce


class MulticlassAccuracy(MulticlassStatScores):
    """Multiclass accuracy from the stat-scores state (reference :152-262)."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Class"

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _accuracy_reduce(
            tp, fp, tn, fn, average=self.average, multidim_average=self.multidim_average, top_k=self.top_k
        )
