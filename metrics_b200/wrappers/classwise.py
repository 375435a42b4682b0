"""`ClasswiseWrapper`: name the entries of a per-class result (reference: src/torchmetrics/wrappers/classwise.py:32-237).

This is the usual way the per-class outputs of the accelerated metrics (``average=None`` stat-score consumers, per-class
AUROC / AP) reach a logger, and it has to be transparent to `MetricCollection`'s compute groups: the wrapper exposes the
wrapped metric's states as its own attributes, so grouped members can share them.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from torch import Tensor

from metrics_b200.metric import Metric
from metrics_b200.wrappers.abstract import WrapperMetric


def _optional_str(value: Any, name: str) -> Optional[str]:
    if value is not None and not isinstance(value, str):
        raise ValueError(f"Expected argument `{name}` to either be `None` or a string but got {value}")
    return value


class ClasswiseWrapper(WrapperMetric):
    """``{f"{prefix}{label}{postfix}": value}`` for every entry of the wrapped metric's 1-d result.  Keys default to
    ``"<metricclassname>_<index>"``; ``labels`` replaces the index."""

    def __init__(self, metric: Metric, labels: Optional[List[str]] = None, prefix: Optional[str] = None,
                 postfix: Optional[str] = None) -> None:
        super().__init__()
        if not isinstance(metric, Metric):
            raise ValueError(f"Expected argument `metric` to be an instance of `torchmetrics.Metric` but got {metric}")
        self.metric = metric
        if labels is not None and not (isinstance(labels, list) and all(isinstance(lab, str) for lab in labels)):
            raise ValueError(f"Expected argument `labels` to either be `None` or a list of strings but got {labels}")
        self.labels = labels
        self._prefix = _optional_str(prefix, "prefix")
        self._postfix = _optional_str(postfix, "postfix")
        self._update_count = 1

    # ------------------------------------------------------------------ naming
    def _convert_output(self, x: Tensor) -> Dict[str, Any]:
        if self._prefix or self._postfix:
            prefix, postfix = self._prefix or "", self._postfix or ""
        else:
            prefix, postfix = f"{type(self.metric).__name__.lower()}_", ""
        names = range(len(x)) if self.labels is None else self.labels
        return {f"{prefix}{name}{postfix}": value for name, value in zip(names, x)}

    # ------------------------------------------------------------------ delegation
    def _filter_kwargs(self, **kwargs: Any) -> Dict[str, Any]:
        return self.metric._filter_kwargs(**kwargs)

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        return self._convert_output(self.metric(*args, **kwargs))

    def update(self, *args: Any, **kwargs: Any) -> None:
        self.metric.update(*args, **kwargs)

    def compute(self) -> Dict[str, Tensor]:
        return self._convert_output(self.metric.compute())

    def reset(self) -> None:
        self.metric.reset()

    # ------------------------------------------------------------------ state transparency (compute groups)
    def __getattr__(self, name: str) -> Any:
        # only reached when normal lookup failed: module attributes of the wrapper first, then the wrapped metric
        if name == "metric" or (name in self.__dict__ and name not in self.metric.__dict__):
            return super().__getattr__(name)
        return getattr(self.metric, name)

    def __setattr__(self, name: str, value: Any) -> None:
        if "metric" in self.__dict__.get("_modules", {}) and name in self.metric._defaults:
            setattr(self.metric, name, value)  # a state: it lives on the wrapped metric
            return
        super().__setattr__(name, value)
        if name == "_update_count" and "metric" in self.__dict__.get("_modules", {}) and isinstance(value, int):
            # a compute group hands its leader's update count to the followers: the wrapped metric shares the leader's
            # states, so it has been "updated" as often (otherwise its `compute` warns about a missing update)
            self.metric._update_count = max(self.metric._update_count, value)
        if name == "metric":  # share the state registry, so the collection sees the wrapped metric's states
            self._defaults = self.metric._defaults
            self._persistent = self.metric._persistent
            self._reductions = self.metric._reductions
