"""Pieces shared by the curve metric classes (AUROC / AveragePrecision / ROC / PrecisionRecallCurve families)."""
from __future__ import annotations

from typing import Any, Callable, Optional

from metrics_b200.metric import Metric
from metrics_b200.utilities.enums import ClassificationTask


class _RankingScore:
    """Class-level metadata of the scalar curve summaries (area-like scores in [0, 1], larger is better)."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0


def build_for_task(
    task: str,
    num_classes: Optional[int],
    num_labels: Optional[int],
    binary: Callable[[], Metric],
    multiclass: Callable[[int], Metric],
    multilabel: Callable[[int], Metric],
) -> Metric:
    """Body of the task wrappers' ``__new__``: pick the concrete class, insisting on the size argument it needs."""
    kind = ClassificationTask.from_str(task)
    if kind == ClassificationTask.BINARY:
        return binary()
    if kind == ClassificationTask.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        return multiclass(num_classes)
    if kind == ClassificationTask.MULTILABEL:
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return multilabel(num_labels)
    raise ValueError(f"Task {task} not supported!")


def finish_score_init(metric: Any, average: Optional[str], validate_args: bool, check: Callable[[], None]) -> None:
    """Tail of the AUROC / AP constructors: run the family's argument validation, remember ``average``."""
    if validate_args:
        check()
    metric.average = average
    metric.validate_args = validate_args
