"""Modular classification metrics (reference: src/torchmetrics/classification/)."""
from metrics_b200.classification.accuracy import MulticlassAccuracy  # noqa: F401
from metrics_b200.classification.confusion_matrix import MulticlassConfusionMatrix  # noqa: F401
from metrics_b200.classification.f_beta import MulticlassF1Score, MulticlassFBetaScore  # noqa: F401
from metrics_b200.classification.stat_scores import MulticlassStatScores  # noqa: F401
