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


# ---- binary / multilabel / task wrapper ------------------------------------------------------------------
from typing import Any  # noqa: E402

from typing_extensions import Literal  # noqa: E402

from metrics_b200.classification.base import _ClassificationTaskWrapper  # noqa: E402
from metrics_b200.classification.stat_scores import BinaryStatScores, MultilabelStatScores, _dispatch  # noqa: E402
from metrics_b200.metric import Metric  # noqa: E402


class BinaryAccuracy(BinaryStatScores):
    """Reference :33-150."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _accuracy_reduce(tp, fp, tn, fn, average="binary", multidim_average=self.multidim_average)


class MultilabelAccuracy(MultilabelStatScores):
    """Reference :265-410."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Label"

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _accuracy_reduce(tp, fp, tn, fn, average=self.average, multidim_average=self.multidim_average, multilabel=True)


class Accuracy(_ClassificationTaskWrapper):
    """Task wrapper (reference :413-530)."""

    def __new__(  # type: ignore[misc]
        cls,
        task: Literal["binary", "multiclass", "multilabel"],
        threshold: float = 0.5,
        num_classes: Optional[int] = None,
        num_labels: Optional[int] = None,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
        multidim_average: Literal["global", "samplewise"] = "global",
        top_k: Optional[int] = 1,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> Metric:
        kwargs.update({"multidim_average": multidim_average, "ignore_index": ignore_index, "validate_args": validate_args})
        return _dispatch(BinaryAccuracy, MulticlassAccuracy, MultilabelAccuracy, task, threshold, num_classes, num_labels,
                         average, top_k, kwargs)
