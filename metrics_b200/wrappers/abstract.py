"""Base of metrics that only forward to another metric (reference: src/torchmetrics/wrappers/abstract.py:19-42)."""
from __future__ import annotations

from typing import Any, Callable

from metrics_b200.metric import Metric


class WrapperMetric(Metric):
    """A wrapper owns no state of its own: the wrapped metric does the syncing, caching and bookkeeping, so the
    `update` / `compute` decorations of `Metric` (sync context, result cache, update counter) are switched off here."""

    def _wrap_update(self, update: Callable) -> Callable:
        return update

    def _wrap_compute(self, compute: Callable) -> Callable:
        return compute

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError
