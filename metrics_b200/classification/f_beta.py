"""F-beta / F1 metric classes (reference: classification/f_beta.py)."""
from __future__ import annotations

from typing import Any, Optional

from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification.stat_scores import MulticlassStatScores
from metrics_b200.functional.classification.f_beta import _fbeta_arg_validation, _fbeta_reduce


class MulticlassFBetaScore(MulticlassStatScores):
    """Multiclass F-beta from the stat-scores state (reference :205-358)."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Class"

    def __init__(
        self,
        beta: float,
        num_classes: int,
        top_k: int = 1,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> None:
        super().__init__(
            num_classes=num_classes,
            top_k=top_k,
            average=average,
            multidim_average=multidim_average,
            ignore_index=ignore_index,
            validate_args=False,
            **kwargs,
        )
        if validate_args:
            _fbeta_arg_validation(beta)
            from metrics_b200.functional.classification.stat_scores import _multiclass_stat_scores_arg_validation

            _multiclass_stat_scores_arg_validation(num_classes, top_k, average, multidim_average, ignore_index, zero_division)
        self.validate_args = validate_args
        self.zero_division = zero_division
        self.beta = beta

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _fbeta_reduce(
            tp, fp, tn, fn, self.beta, average=self.average, multidim_average=self.multidim_average,
            zero_division=self.zero_division,
        )


class MulticlassF1Score(MulticlassFBetaScore):
    """Multiclass F1 (reference :739-871)."""

    def __init__(
        self,
        num_classes: int,
        top_k: int = 1,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> None:
        super().__init__(
            beta=1.0,
            num_classes=num_classes,
            top_k=top_k,
            average=average,
            multidim_average=multidim_average,
            ignore_index=ignore_index,
            validate_args=validate_args,
            zero_division=zero_division,
            **kwargs,
        )


# ---- binary / multilabel / task wrappers -----------------------------------------------------------------
from metrics_b200.classification.base import _ClassificationTaskWrapper  # noqa: E402
from metrics_b200.classification.stat_scores import BinaryStatScores, MultilabelStatScores  # noqa: E402
from metrics_b200.metric import Metric  # noqa: E402
from metrics_b200.utilities.enums import ClassificationTask  # noqa: E402


class BinaryFBetaScore(BinaryStatScores):
    """Reference :41-202."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0

    def __init__(
        self,
        beta: float,
        threshold: float = 0.5,
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> None:
        super().__init__(threshold=threshold, multidim_average=multidim_average, ignore_index=ignore_index,
                         validate_args=False, **kwargs)
        if validate_args:
            from metrics_b200.functional.classification.stat_scores import _binary_stat_scores_arg_validation

            _fbeta_arg_validation(beta)
            _binary_stat_scores_arg_validation(threshold, multidim_average, ignore_index, zero_division)
        self.validate_args = validate_args
        self.zero_division = zero_division
        self.beta = beta

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _fbeta_reduce(tp, fp, tn, fn, self.beta, average="binary", multidim_average=self.multidim_average,
                             zero_division=self.zero_division)


class BinaryF1Score(BinaryFBetaScore):
    """Reference :560-736."""

    def __init__(
        self,
        threshold: float = 0.5,
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> None:
        super().__init__(beta=1.0, threshold=threshold, multidim_average=multidim_average, ignore_index=ignore_index,
                         validate_args=validate_args, zero_division=zero_division, **kwargs)


class MultilabelFBetaScore(MultilabelStatScores):
    """Reference :361-557."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0
    plot_legend_name: str = "Label"

    def __init__(
        self,
        beta: float,
        num_labels: int,
        threshold: float = 0.5,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> None:
        super().__init__(num_labels=num_labels, threshold=threshold, average=average, multidim_average=multidim_average,
                         ignore_index=ignore_index, validate_args=False, **kwargs)
        if validate_args:
            from metrics_b200.functional.classification.stat_scores import _multilabel_stat_scores_arg_validation

            _fbeta_arg_validation(beta)
            _multilabel_stat_scores_arg_validation(num_labels, threshold, average, multidim_average, ignore_index, zero_division)
        self.validate_args = validate_args
        self.zero_division = zero_division
        self.beta = beta

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _fbeta_reduce(tp, fp, tn, fn, self.beta, average=self.average, multidim_average=self.multidim_average,
                             multilabel=True, zero_division=self.zero_division)


class MultilabelF1Score(MultilabelFBetaScore):
    """Reference :874-1060."""

    def __init__(
        self,
        num_labels: int,
        threshold: float = 0.5,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> None:
        super().__init__(beta=1.0, num_labels=num_labels, threshold=threshold, average=average,
                         multidim_average=multidim_average, ignore_index=ignore_index, validate_args=validate_args,
                         zero_division=zero_division, **kwargs)


def _fbeta_dispatch(beta, f1, task, threshold, num_classes, num_labels, average, multidim_average, top_k, ignore_index,
                    validate_args, zero_division, kwargs):
    task = ClassificationTask.from_str(task)
    kwargs.update({"multidim_average": multidim_average, "ignore_index": ignore_index, "validate_args": validate_args,
                   "zero_division": zero_division})
    pre = () if f1 else (beta,)
    if task == ClassificationTask.BINARY:
        return (BinaryF1Score if f1 else BinaryFBetaScore)(*pre, threshold, **kwargs)
    if task == ClassificationTask.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        if not isinstance(top_k, int):
            raise ValueError(f"`top_k` is expected to be `int` but `{type(top_k)} was passed.`")
        return (MulticlassF1Score if f1 else MulticlassFBetaScore)(*pre, num_classes, top_k, average, **kwargs)
    if task == ClassificationTask.MULTILABEL:
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return (MultilabelF1Score if f1 else MultilabelFBetaScore)(*pre, num_labels, threshold, average, **kwargs)
    raise ValueError(f"Task {task} not supported!")


class FBetaScore(_ClassificationTaskWrapper):
    """Task wrapper (reference :1063-1140)."""

    def __new__(  # type: ignore[misc]
        cls,
        task: Literal["binary", "multiclass", "multilabel"],
        beta: float = 1.0,
        threshold: float = 0.5,
        num_classes: Optional[int] = None,
        num_labels: Optional[int] = None,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
        multidim_average: Optional[Literal["global", "samplewise"]] = "global",
        top_k: Optional[int] = 1,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> Metric:
        return _fbeta_dispatch(beta, False, task, threshold, num_classes, num_labels, average, multidim_average, top_k,
                               ignore_index, validate_args, zero_division, kwargs)


class F1Score(_ClassificationTaskWrapper):
    """Task wrapper (reference :1143-1221)."""

    def __new__(  # type: ignore[misc]
        cls,
        task: Literal["binary", "multiclass", "multilabel"],
        threshold: float = 0.5,
        num_classes: Optional[int] = None,
        num_labels: Optional[int] = None,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
        multidim_average: Optional[Literal["global", "samplewise"]] = "global",
        top_k: Optional[int] = 1,
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        zero_division: float = 0,
        **kwargs: Any,
    ) -> Metric:
        return _fbeta_dispatch(1.0, True, task, threshold, num_classes, num_labels, average, multidim_average, top_k,
                               ignore_index, validate_args, zero_division, kwargs)
