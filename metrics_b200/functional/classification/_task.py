"""Task dispatch for the functional ``<metric>(preds, target, task=...)`` wrappers."""
from __future__ import annotations

from typing import Any, Callable, Optional

from metrics_b200.utilities.enums import ClassificationTask, ClassificationTaskNoMultilabel


def call_for_task(task: str, num_classes: Optional[int], num_labels: Optional[int], binary: Callable[[], Any],
                  multiclass: Callable[[int], Any], multilabel: Optional[Callable[[int], Any]]) -> Any:
    """Run the binary / multiclass / multilabel functional, insisting on the size argument the task needs."""
    kind = (ClassificationTask if multilabel is not None else ClassificationTaskNoMultilabel).from_str(task)
    if kind == ClassificationTask.BINARY:
        return binary()
    if kind == ClassificationTask.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        return multiclass(num_classes)
    if not isinstance(num_labels, int):
        raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
    return multilabel(num_labels)
