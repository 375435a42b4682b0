"""metrics_b200 — the TorchMetrics update()/compute()/sync hot path, rebuilt for B200 (sm_100a).

Drop-in for the reference's `Metric` / `MetricCollection` API and the classification / detection / regression
metric families named in DESIGN.md; every per-batch `_update` runs in a hand-written CUDA kernel reached through
the C-ABI in `include/metrics_b200.h` (there is no CPU fallback: inputs must be CUDA tensors).
"""
from metrics_b200.collections import MetricCollection  # noqa: F401
from metrics_b200.metric import CompositionalMetric, Metric  # noqa: F401

__version__ = "0.1.0"
__all__ = ["Metric", "MetricCollection", "CompositionalMetric"]
