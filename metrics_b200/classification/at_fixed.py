"""RecallAtFixedPrecision / PrecisionAtFixedRecall / SensitivityAtSpecificity / SpecificityAtSensitivity metric classes
(reference: classification/{recall_fixed_precision,precision_fixed_recall,sensitivity_specificity,specificity_sensitivity}.py):
precision-recall-curve state holders whose `compute` picks one operating point per curve."""
from __future__ import annotations

from typing import Any, List, Optional, Union

from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification.base import _ClassificationTaskWrapper
from metrics_b200.classification.precision_recall_curve import (
    BinaryPrecisionRecallCurve,
    MulticlassPrecisionRecallCurve,
    MultilabelPrecisionRecallCurve,
)
from metrics_b200.functional.classification.at_fixed import (
    _publish_signature,
    _FAMILIES,
    _binary_at_fixed_compute,
    _floor_validation,
    _multiclass_at_fixed_compute,
    _multilabel_at_fixed_compute,
)
from metrics_b200.metric import Metric
from metrics_b200.utilities.enums import ClassificationTask

_STEM = {
    "recall_at_fixed_precision": "RecallAtFixedPrecision",
    "precision_at_fixed_recall": "PrecisionAtFixedRecall",
    "sensitivity_at_specificity": "SensitivityAtSpecificity",
    "specificity_at_sensitivity": "SpecificityAtSensitivity",
}
_Thr = Optional[Union[int, List[float], Tensor]]


def _family(kind: str):
    fam = _FAMILIES[kind]
    stem = _STEM[kind]
    attrs = {"is_differentiable": False, "higher_is_better": None, "full_state_update": False,
             "plot_lower_bound": 0.0, "plot_upper_bound": 1.0, "__module__": __name__}

    def named(floor: Optional[float], kwargs: dict) -> float:
        # the reference names this constructor argument per family (min_precision, min_recall, ...)
        if floor is None:
            if fam.arg not in kwargs:
                raise TypeError(f"missing required argument `{fam.arg}`")
            floor = kwargs.pop(fam.arg)
        return floor

    def b_init(self, floor: Optional[float] = None, thresholds: _Thr = None, ignore_index: Optional[int] = None,
               validate_args: bool = True, **kwargs: Any) -> None:
        floor = named(floor, kwargs)
        BinaryPrecisionRecallCurve.__init__(self, thresholds, ignore_index, validate_args=validate_args, **kwargs)
        if validate_args:
            _floor_validation(fam.arg, floor)
        setattr(self, fam.arg, floor)

    def b_compute(self):
        return _binary_at_fixed_compute(kind, self._state(), self.thresholds, getattr(self, fam.arg))

    def mc_init(self, num_classes: int, floor: Optional[float] = None, thresholds: _Thr = None,
                ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> None:
        floor = named(floor, kwargs)
        MulticlassPrecisionRecallCurve.__init__(self, num_classes=num_classes, thresholds=thresholds,
                                                ignore_index=ignore_index, validate_args=validate_args, **kwargs)
        if validate_args:
            _floor_validation(fam.arg, floor)
        setattr(self, fam.arg, floor)

    def mc_compute(self):
        return _multiclass_at_fixed_compute(kind, self._state(), self.num_classes, self.thresholds, getattr(self, fam.arg))

    def ml_init(self, num_labels: int, floor: Optional[float] = None, thresholds: _Thr = None,
                ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> None:
        floor = named(floor, kwargs)
        MultilabelPrecisionRecallCurve.__init__(self, num_labels=num_labels, thresholds=thresholds,
                                                ignore_index=ignore_index, validate_args=validate_args, **kwargs)
        if validate_args:
            _floor_validation(fam.arg, floor)
        setattr(self, fam.arg, floor)

    def ml_compute(self):
        return _multilabel_at_fixed_compute(kind, self._state(), self.num_labels, self.thresholds, self.ignore_index,
                                            getattr(self, fam.arg))

    doc = f"{kind.replace('_', ' ')} (reference classification/{fam.reference}); first positional argument after the task size is `{fam.arg}`."
    for init in (b_init, mc_init, ml_init):
        _publish_signature(init, fam.arg)
    b = type(f"Binary{stem}", (BinaryPrecisionRecallCurve,), {**attrs, "__init__": b_init, "compute": b_compute, "__doc__": "Binary " + doc})
    mc = type(f"Multiclass{stem}", (MulticlassPrecisionRecallCurve,),
              {**attrs, "plot_legend_name": "Class", "__init__": mc_init, "compute": mc_compute, "__doc__": "Multiclass " + doc})
    ml = type(f"Multilabel{stem}", (MultilabelPrecisionRecallCurve,),
              {**attrs, "plot_legend_name": "Label", "__init__": ml_init, "compute": ml_compute, "__doc__": "Multilabel " + doc})

    def __new__(cls, task: Literal["binary", "multiclass", "multilabel"], floor: Optional[float] = None, thresholds: _Thr = None,
                num_classes: Optional[int] = None, num_labels: Optional[int] = None, ignore_index: Optional[int] = None,
                validate_args: bool = True, **kwargs: Any) -> Metric:
        if floor is None:  # the reference names this argument per family (min_precision, min_recall, ...)
            floor = kwargs.pop(fam.arg)
        task = ClassificationTask.from_str(task)
        kwargs.update({"thresholds": thresholds, "ignore_index": ignore_index, "validate_args": validate_args})
        if task == ClassificationTask.BINARY:
            return b(floor, **kwargs)
        if task == ClassificationTask.MULTICLASS:
            if not isinstance(num_classes, int):
                raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
            return mc(num_classes, floor, **kwargs)
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return ml(num_labels, floor, **kwargs)

    _publish_signature(__new__, fam.arg)
    wrapper = type(stem, (_ClassificationTaskWrapper,), {"__new__": __new__, "__module__": __name__,
                                                         "__doc__": f"Task wrapper for {kind.replace('_', ' ')}."})
    return b, mc, ml, wrapper


(BinaryRecallAtFixedPrecision, MulticlassRecallAtFixedPrecision, MultilabelRecallAtFixedPrecision,
 RecallAtFixedPrecision) = _family("recall_at_fixed_precision")
(BinaryPrecisionAtFixedRecall, MulticlassPrecisionAtFixedRecall, MultilabelPrecisionAtFixedRecall,
 PrecisionAtFixedRecall) = _family("precision_at_fixed_recall")
(BinarySensitivityAtSpecificity, MulticlassSensitivityAtSpecificity, MultilabelSensitivityAtSpecificity,
 SensitivityAtSpecificity) = _family("sensitivity_at_specificity")
(BinarySpecificityAtSensitivity, MulticlassSpecificityAtSensitivity, MultilabelSpecificityAtSensitivity,
 SpecificityAtSensitivity) = _family("specificity_at_sensitivity")
