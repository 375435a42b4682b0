"""Detection metrics (reference: src/torchmetrics/detection/)."""
from metrics_b200.detection.mean_ap import MeanAveragePrecision  # noqa: F401
