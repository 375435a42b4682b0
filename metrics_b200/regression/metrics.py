"""Regression metric classes: running-sum states updated by the K9 kernel (reference: regression/*.py).

State names, dtypes and reductions are the reference's (e.g. `sum_squared_error` / `total` with "sum")."""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor, tensor

from metrics_b200.functional.regression import metrics as F
from metrics_b200.metric import Metric
from metrics_b200.utilities.exceptions import TorchMetricsUserError


class _SumOverN(Metric):
    is_differentiable: bool = False  # kernel launches carry no autograd graph (reference: True)
    full_state_update: bool = False
    plot_lower_bound: float = 0.0


class MeanSquaredError(_SumOverN):
    """Reference regression/mse.py:27-110."""

    higher_is_better: bool = False

    def __init__(self, squared: bool = True, num_outputs: int = 1, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if not isinstance(squared, bool):
            raise ValueError(f"Expected argument `squared` to be a boolean but got {squared}")
        self.squared = squared
        if not (isinstance(num_outputs, int) and num_outputs > 0):
            raise ValueError(f"Expected num_outputs to be a positive integer but got {num_outputs}")
        self.num_outputs = num_outputs
        self.add_state("sum_squared_error", default=torch.zeros(num_outputs), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        sse, n = F._mean_squared_error_update(preds, target, self.num_outputs)
        self.sum_squared_error += sse
        self.total += n

    def compute(self) -> Tensor:
        return F._mean_squared_error_compute(self.sum_squared_error, self.total, self.squared)


class MeanAbsoluteError(_SumOverN):
    """Reference regression/mae.py:27-100."""

    higher_is_better: bool = False

    def __init__(self, num_outputs: int = 1, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if not (isinstance(num_outputs, int) and num_outputs > 0):
            raise ValueError(f"Expected num_outputs to be a positive integer but got {num_outputs}")
        self.num_outputs = num_outputs
        self.add_state("sum_abs_error", default=torch.zeros(num_outputs), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        s, n = F._mean_absolute_error_update(preds, target, self.num_outputs)
        self.sum_abs_error += s
        self.total += n

    def compute(self) -> Tensor:
        return F._mean_absolute_error_compute(self.sum_abs_error, self.total)


class MeanAbsolutePercentageError(_SumOverN):
    """Reference regression/mape.py:30-100."""

    higher_is_better: bool = False

    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state("sum_abs_per_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0.0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        s, n = F._mean_absolute_percentage_error_update(preds, target)
        self.sum_abs_per_error += s
        self.total += n

    def compute(self) -> Tensor:
        return F._mean_absolute_percentage_error_compute(self.sum_abs_per_error, self.total)


class SymmetricMeanAbsolutePercentageError(_SumOverN):
    """Reference regression/symmetric_mape.py:30-100."""

    higher_is_better: bool = False
    plot_upper_bound: float = 2.0

    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state("sum_abs_per_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0.0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        s, n = F._symmetric_mean_absolute_percentage_error_update(preds, target)
        self.sum_abs_per_error += s
        self.total += n

    def compute(self) -> Tensor:
        return F._symmetric_mean_absolute_percentage_error_compute(self.sum_abs_per_error, self.total)


class WeightedMeanAbsolutePercentageError(_SumOverN):
    """Reference regression/wmape.py:30-100."""

    higher_is_better: bool = False

    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state("sum_abs_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("sum_scale", default=tensor(0.0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        a, b = F._weighted_mean_absolute_percentage_error_update(preds, target)
        self.sum_abs_error += a
        self.sum_scale += b

    def compute(self) -> Tensor:
        return F._weighted_mean_absolute_percentage_error_compute(self.sum_abs_error, self.sum_scale)


class MeanSquaredLogError(_SumOverN):
    """Reference regression/log_mse.py:27-100."""

    higher_is_better: bool = False

    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state("sum_squared_log_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        s, n = F._mean_squared_log_error_update(preds, target)
        self.sum_squared_log_error += s
        self.total += n

    def compute(self) -> Tensor:
        return F._mean_squared_log_error_compute(self.sum_squared_log_error, self.total)


class LogCoshError(_SumOverN):
    """Reference regression/log_cosh.py:27-100."""

    higher_is_better: bool = False

    def __init__(self, num_outputs: int = 1, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if not isinstance(num_outputs, int) and num_outputs < 1:
            raise ValueError(f"Expected argument `num_outputs` to be an int larger than 0, but got {num_outputs}")
        self.num_outputs = num_outputs
        self.add_state("sum_log_cosh_error", default=torch.zeros(num_outputs), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        s, n = F._log_cosh_error_update(preds, target, self.num_outputs)
        self.sum_log_cosh_error += s
        self.total += n

    def compute(self) -> Tensor:
        return F._log_cosh_error_compute(self.sum_log_cosh_error, self.total)


class MinkowskiDistance(Metric):
    """Reference regression/minkowski.py:27-100."""

    is_differentiable: Optional[bool] = False  # kernel launches carry no autograd graph (reference: True)
    higher_is_better: Optional[bool] = False
    full_state_update: Optional[bool] = False
    plot_lower_bound: float = 0.0

    def __init__(self, p: float, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if not (isinstance(p, (float, int)) and p >= 1):
            raise TorchMetricsUserError(f"Argument ``p`` must be a float or int greater than 1, but got {p}")
        self.p = p
        self.add_state("minkowski_dist_sum", default=tensor(0.0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, targets: Tensor) -> None:
        self.minkowski_dist_sum += F._minkowski_distance_update(preds, targets, self.p)

    def compute(self) -> Tensor:
        return F._minkowski_distance_compute(self.minkowski_dist_sum, self.p)


class R2Score(Metric):
    """Reference regression/r2.py:27-160."""

    is_differentiable: bool = False  # kernel launches carry no autograd graph (reference: True)
    higher_is_better: bool = True
    full_state_update: bool = False
    plot_upper_bound: float = 1.0

    def __init__(self, adjusted: int = 0, multioutput: str = "uniform_average", **kwargs: Any) -> None:
        super().__init__(**kwargs)
        if adjusted < 0 or not isinstance(adjusted, int):
            raise ValueError("`adjusted` parameter should be an integer larger or equal to 0.")
        self.adjusted = adjusted
        allowed = ("raw_values", "uniform_average", "variance_weighted")
        if multioutput not in allowed:
            raise ValueError(f"Invalid input to argument `multioutput`. Choose one of the following: {allowed}")
        self.multioutput = multioutput
        self.add_state("sum_squared_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("sum_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("residual", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        sso, so, rss, n = F._r2_score_update(preds, target)
        self.sum_squared_error = self.sum_squared_error + sso
        self.sum_error = self.sum_error + so
        self.residual = self.residual + rss
        self.total = self.total + n

    def compute(self) -> Tensor:
        return F._r2_score_compute(self.sum_squared_error, self.sum_error, self.residual, self.total, self.adjusted, self.multioutput)


class RelativeSquaredError(Metric):
    """Reference regression/rse.py:27-110."""

    is_differentiable: bool = False  # kernel launches carry no autograd graph (reference: True)
    higher_is_better: bool = False
    full_state_update: bool = False

    def __init__(self, num_outputs: int = 1, squared: bool = True, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.num_outputs = num_outputs
        # state names as in the reference (rse.py:84-87), whose checkpoints must load: `sum_squared_error` holds the sum of
        # squared TARGETS and `sum_error` the sum of targets (they receive `_r2_score_update`'s first two outputs, :93-96);
        # the residual sum of squares is `residual`
        self.add_state("sum_squared_error", default=torch.zeros(num_outputs), dist_reduce_fx="sum")
        self.add_state("sum_error", default=torch.zeros(num_outputs), dist_reduce_fx="sum")
        self.add_state("residual", default=torch.zeros(num_outputs), dist_reduce_fx="sum")
        self.add_state("total", default=tensor(0), dist_reduce_fx="sum")
        self.squared = squared

    def update(self, preds: Tensor, target: Tensor) -> None:
        sum_squared_obs, sum_obs, rss, n = F._r2_score_update(preds, target)
        self.sum_squared_error += sum_squared_obs
        self.sum_error += sum_obs
        self.residual += rss
        self.total += n

    def compute(self) -> Tensor:
        return F._relative_squared_error_compute(self.sum_squared_error, self.sum_error, self.residual, self.total, self.squared)


class ExplainedVariance(Metric):
    """Reference regression/explained_variance.py:30-130."""

    is_differentiable: bool = False  # kernel launches carry no autograd graph (reference: True)
    higher_is_better: bool = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0

    def __init__(self, multioutput: str = "uniform_average", **kwargs: Any) -> None:
        super().__init__(**kwargs)
        allowed = ("raw_values", "uniform_average", "variance_weighted")
        if multioutput not in allowed:
            raise ValueError(f"Invalid input to argument `multioutput`. Choose one of the following: {allowed}")
        self.multioutput = multioutput
        self.add_state("sum_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("sum_squared_error", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("sum_target", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("sum_squared_target", default=tensor(0.0), dist_reduce_fx="sum")
        self.add_state("num_obs", default=tensor(0.0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        n, se, sse, st, sst = F._explained_variance_update(preds, target)
        self.num_obs = self.num_obs + n
        self.sum_error = self.sum_error + se
        self.sum_squared_error = self.sum_squared_error + sse
        self.sum_target = self.sum_target + st
        self.sum_squared_target = self.sum_squared_target + sst

    def compute(self) -> Tensor:
        return F._explained_variance_compute(
            self.num_obs, self.sum_error, self.sum_squared_error, self.sum_target, self.sum_squared_target, self.multioutput
        )


class TweedieDevianceScore(Metric):
    """Mean Tweedie deviance of the given ``power`` (reference regression/tweedie_deviance.py:30-111): two ``sum`` states,
    filled by one K9 pass per update (which also carries the domain check)."""

    is_differentiable: bool = False  # kernel launches carry no autograd graph (reference: True)
    higher_is_better: Optional[bool] = None
    full_state_update: bool = False
    plot_lower_bound: float = 0.0

    def __init__(self, power: float = 0.0, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        F._tweedie_power_check(power)
        self.power: float = power
        self.add_state("sum_deviance_score", torch.tensor(0.0), dist_reduce_fx="sum")
        self.add_state("num_observations", torch.tensor(0), dist_reduce_fx="sum")

    def update(self, preds: Tensor, targets: Tensor) -> None:
        sum_deviance_score, num_observations = F._tweedie_deviance_score_update(preds, targets, self.power)
        self.sum_deviance_score += sum_deviance_score
        self.num_observations += num_observations

    def compute(self) -> Tensor:
        return F._tweedie_deviance_score_compute(self.sum_deviance_score, self.num_observations)


class CriticalSuccessIndex(Metric):
    """Reference regression/csi.py:27-110: three ``sum`` states, or three ``cat`` list states (one ``[S]`` entry per update)
    when a sequence dimension is kept."""

    is_differentiable: bool = False
    higher_is_better: bool = True

    def __init__(self, threshold: float, keep_sequence_dim: Optional[int] = None, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.threshold = float(threshold)
        if keep_sequence_dim and (not isinstance(keep_sequence_dim, int) or keep_sequence_dim < 0):
            raise ValueError(f"Expected keep_sequence_dim to be a non-negative integer but got {keep_sequence_dim}")
        self.keep_sequence_dim = keep_sequence_dim
        for name in ("hits", "misses", "false_alarms"):
            if keep_sequence_dim is None:
                self.add_state(name, default=torch.tensor(0), dist_reduce_fx="sum")
            else:
                self.add_state(name + "_list", default=[], dist_reduce_fx="cat")

    def update(self, preds: Tensor, target: Tensor) -> None:
        hits, misses, false_alarms = F._critical_success_index_update(preds, target, self.threshold, self.keep_sequence_dim)
        if self.keep_sequence_dim is None:
            self.hits += hits
            self.misses += misses
            self.false_alarms += false_alarms
        else:
            self.hits_list.append(hits)
            self.misses_list.append(misses)
            self.false_alarms_list.append(false_alarms)

    def compute(self) -> Tensor:
        if self.keep_sequence_dim is None:
            return F._critical_success_index_compute(self.hits, self.misses, self.false_alarms)
        from metrics_b200.utilities.data import dim_zero_cat

        return F._critical_success_index_compute(dim_zero_cat(self.hits_list), dim_zero_cat(self.misses_list),
                                                 dim_zero_cat(self.false_alarms_list))
