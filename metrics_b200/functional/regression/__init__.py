"""Functional regression metrics (reference: src/torchmetrics/functional/regression/)."""
from metrics_b200.functional.regression.metrics import (  # noqa: F401
    explained_variance,
    log_cosh_error,
    mean_absolute_error,
    mean_absolute_percentage_error,
    mean_squared_error,
    mean_squared_log_error,
    minkowski_distance,
    r2_score,
    relative_squared_error,
    symmetric_mean_absolute_percentage_error,
    weighted_mean_absolute_percentage_error,
)
