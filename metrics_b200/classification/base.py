"""Base of the task-dispatching wrapper classes (reference: classification/base.py:19-32): calling e.g.
``Accuracy(task="multiclass", num_classes=3)`` returns a ``MulticlassAccuracy``."""
from metrics_b200.metric import Metric


class _ClassificationTaskWrapper(Metric):
    def update(self, *args, **kwargs) -> None:
        raise NotImplementedError(f"{self.__class__.__name__} metric does not have a global `update` method. Use the task specific metric.")

    def compute(self) -> None:
        raise NotImplementedError(f"{self.__class__.__name__} metric does not have a global `compute` method. Use the task specific metric.")
