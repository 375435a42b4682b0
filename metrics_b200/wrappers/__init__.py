"""Metric wrappers that sit directly on the accelerated path (reference: src/torchmetrics/wrappers/)."""
from metrics_b200.wrappers.abstract import WrapperMetric  # noqa: F401
from metrics_b200.wrappers.classwise import ClasswiseWrapper  # noqa: F401
