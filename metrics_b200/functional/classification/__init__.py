"""Functional classification metrics (reference: src/torchmetrics/functional/classification/)."""
from metrics_b200.functional.classification.accuracy import multiclass_accuracy  # noqa: F401
from metrics_b200.functional.classification.confusion_matrix import multiclass_confusion_matrix  # noqa: F401
from metrics_b200.functional.classification.f_beta import multiclass_f1_score, multiclass_fbeta_score  # noqa: F401
from metrics_b200.functional.classification.stat_scores import multiclass_stat_scores  # noqa: F401
