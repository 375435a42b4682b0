"""JaccardIndex / CohenKappa / MatthewsCorrCoef metric classes (reference: classification/{jaccard,cohen_kappa,
matthews_corrcoef}.py): confusion-matrix state holders (K1 / K2 kernels) with a different `compute`."""
from __future__ import annotations

from typing import Any, Optional

from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification.base import _ClassificationTaskWrapper
from metrics_b200.classification.confusion_matrix import (
    BinaryConfusionMatrix,
    MulticlassConfusionMatrix,
    MultilabelConfusionMatrix,
)
from metrics_b200.functional.classification.confmat_metrics import (
    _cohen_kappa_reduce,
    _cohen_kappa_weights_validation,
    _jaccard_average_validation,
    _jaccard_index_reduce,
    _matthews_corrcoef_reduce,
)
from metrics_b200.metric import Metric
from metrics_b200.utilities.enums import ClassificationTask, ClassificationTaskNoMultilabel


class _Score01:
    is_differentiable: bool = False
    higher_is_better: Optional[bool] = True
    full_state_update: bool = False
    plot_lower_bound: float = 0.0
    plot_upper_bound: float = 1.0


# ---- Jaccard ------------------------------------------------------------------------------------------------------------
class BinaryJaccardIndex(_Score01, BinaryConfusionMatrix):
    """Reference jaccard.py:40-155."""

    def __init__(self, threshold: float = 0.5, ignore_index: Optional[int] = None, validate_args: bool = True,
                 zero_division: float = 0, **kwargs: Any) -> None:
        super().__init__(threshold=threshold, ignore_index=ignore_index, normalize=None, validate_args=validate_args, **kwargs)
        self.zero_division = zero_division

    def compute(self) -> Tensor:
        return _jaccard_index_reduce(self.confmat, average="binary", zero_division=self.zero_division)


class MulticlassJaccardIndex(_Score01, MulticlassConfusionMatrix):
    """Reference jaccard.py:158-292."""

    plot_legend_name: str = "Class"

    def __init__(self, num_classes: int, average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
                 ignore_index: Optional[int] = None, validate_args: bool = True, zero_division: float = 0,
                 **kwargs: Any) -> None:
        super().__init__(num_classes=num_classes, ignore_index=ignore_index, normalize=None, validate_args=validate_args, **kwargs)
        if validate_args:
            _jaccard_average_validation(average)
        self.average = average
        self.zero_division = zero_division

    def compute(self) -> Tensor:
        return _jaccard_index_reduce(self.confmat, average=self.average, ignore_index=self.ignore_index,
                                     zero_division=self.zero_division)


class MultilabelJaccardIndex(_Score01, MultilabelConfusionMatrix):
    """Reference jaccard.py:295-431."""

    plot_legend_name: str = "Label"

    def __init__(self, num_labels: int, threshold: float = 0.5,
                 average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
                 ignore_index: Optional[int] = None, validate_args: bool = True, zero_division: float = 0,
                 **kwargs: Any) -> None:
        super().__init__(num_labels=num_labels, threshold=threshold, ignore_index=ignore_index, normalize=None,
                         validate_args=validate_args, **kwargs)
        if validate_args:
            _jaccard_average_validation(average)
        self.average = average
        self.zero_division = zero_division

    def compute(self) -> Tensor:
        return _jaccard_index_reduce(self.confmat, average=self.average, zero_division=self.zero_division)


class JaccardIndex(_ClassificationTaskWrapper):
    """Task wrapper (reference jaccard.py:434-492)."""

    def __new__(cls, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,  # type: ignore[misc]
                num_classes: Optional[int] = None, num_labels: Optional[int] = None,
                average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
                ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> Metric:
        task = ClassificationTask.from_str(task)
        kwargs.update({"ignore_index": ignore_index, "validate_args": validate_args})
        if task == ClassificationTask.BINARY:
            return BinaryJaccardIndex(threshold, **kwargs)
        if task == ClassificationTask.MULTICLASS:
            if not isinstance(num_classes, int):
                raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
            return MulticlassJaccardIndex(num_classes, average, **kwargs)
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return MultilabelJaccardIndex(num_labels, threshold, average, **kwargs)


# ---- Cohen's kappa -------------------------------------------------------------------------------------------------------
class BinaryCohenKappa(_Score01, BinaryConfusionMatrix):
    """Reference cohen_kappa.py:36-158."""

    plot_lower_bound: float = 0.0  # as declared by the reference (the score itself can be negative)

    def __init__(self, threshold: float = 0.5, ignore_index: Optional[int] = None,
                 weights: Optional[Literal["linear", "quadratic", "none"]] = None, validate_args: bool = True,
                 **kwargs: Any) -> None:
        super().__init__(threshold, ignore_index, normalize=None, validate_args=validate_args, **kwargs)
        if validate_args:
            _cohen_kappa_weights_validation(weights)
        self.weights = weights

    def compute(self) -> Tensor:
        return _cohen_kappa_reduce(self.confmat, self.weights)


class MulticlassCohenKappa(_Score01, MulticlassConfusionMatrix):
    """Reference cohen_kappa.py:161-287."""

    plot_lower_bound: float = 0.0  # as declared by the reference (the score itself can be negative)
    plot_legend_name: str = "Class"

    def __init__(self, num_classes: int, ignore_index: Optional[int] = None,
                 weights: Optional[Literal["linear", "quadratic", "none"]] = None, validate_args: bool = True,
                 **kwargs: Any) -> None:
        super().__init__(num_classes, ignore_index, normalize=None, validate_args=validate_args, **kwargs)
        if validate_args:
            _cohen_kappa_weights_validation(weights)
        self.weights = weights

    def compute(self) -> Tensor:
        return _cohen_kappa_reduce(self.confmat, self.weights)


class CohenKappa(_ClassificationTaskWrapper):
    """Task wrapper (reference cohen_kappa.py:290-338); binary and multiclass only."""

    def __new__(cls, task: Literal["binary", "multiclass"], threshold: float = 0.5,  # type: ignore[misc]
                num_classes: Optional[int] = None, weights: Optional[Literal["linear", "quadratic", "none"]] = None,
                ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> Metric:
        task = ClassificationTaskNoMultilabel.from_str(task)
        kwargs.update({"weights": weights, "ignore_index": ignore_index, "validate_args": validate_args})
        if task == ClassificationTaskNoMultilabel.BINARY:
            return BinaryCohenKappa(threshold, **kwargs)
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        return MulticlassCohenKappa(num_classes, **kwargs)


# ---- Matthews correlation coefficient ---------------------------------------------------------------------------------------
class BinaryMatthewsCorrCoef(_Score01, BinaryConfusionMatrix):
    """Reference matthews_corrcoef.py:40-145."""

    plot_lower_bound: float = 0.0  # as declared by the reference (the score itself can be negative)

    def __init__(self, threshold: float = 0.5, ignore_index: Optional[int] = None, validate_args: bool = True,
                 **kwargs: Any) -> None:
        super().__init__(threshold, ignore_index, normalize=None, validate_args=validate_args, **kwargs)

    def compute(self) -> Tensor:
        return _matthews_corrcoef_reduce(self.confmat)


class MulticlassMatthewsCorrCoef(_Score01, MulticlassConfusionMatrix):
    """Reference matthews_corrcoef.py:148-257."""

    plot_lower_bound: float = 0.0  # as declared by the reference (the score itself can be negative)
    plot_legend_name: str = "Class"

    def __init__(self, num_classes: int, ignore_index: Optional[int] = None, validate_args: bool = True,
                 **kwargs: Any) -> None:
        super().__init__(num_classes, ignore_index, normalize=None, validate_args=validate_args, **kwargs)

    def compute(self) -> Tensor:
        return _matthews_corrcoef_reduce(self.confmat)


class MultilabelMatthewsCorrCoef(_Score01, MultilabelConfusionMatrix):
    """Reference matthews_corrcoef.py:260-368."""

    plot_lower_bound: float = 0.0  # as declared by the reference (the score itself can be negative)
    plot_legend_name: str = "Label"

    def __init__(self, num_labels: int, threshold: float = 0.5, ignore_index: Optional[int] = None,
                 validate_args: bool = True, **kwargs: Any) -> None:
        super().__init__(num_labels, threshold, ignore_index, normalize=None, validate_args=validate_args, **kwargs)

    def compute(self) -> Tensor:
        return _matthews_corrcoef_reduce(self.confmat)


class MatthewsCorrCoef(_ClassificationTaskWrapper):
    """Task wrapper (reference matthews_corrcoef.py:371-420)."""

    def __new__(cls, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,  # type: ignore[misc]
                num_classes: Optional[int] = None, num_labels: Optional[int] = None, ignore_index: Optional[int] = None,
                validate_args: bool = True, **kwargs: Any) -> Metric:
        task = ClassificationTask.from_str(task)
        kwargs.update({"ignore_index": ignore_index, "validate_args": validate_args})
        if task == ClassificationTask.BINARY:
            return BinaryMatthewsCorrCoef(threshold, **kwargs)
        if task == ClassificationTask.MULTICLASS:
            if not isinstance(num_classes, int):
                raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
            return MulticlassMatthewsCorrCoef(num_classes, **kwargs)
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return MultilabelMatthewsCorrCoef(num_labels, threshold, **kwargs)
