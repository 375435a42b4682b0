"""Precision / Recall / Specificity / NegativePredictiveValue / HammingDistance metric classes.

Reference: classification/{precision_recall,specificity,negative_predictive_value,hamming}.py.  Each class is a stat-scores
state holder (same `tp fp tn fn` states, same constructor) whose `compute` applies one ratio — built from the table in
functional/classification/ratio_metrics.py.
"""
from __future__ import annotations

from typing import Any, Optional

from torch import Tensor
from typing_extensions import Literal

from metrics_b200.classification.base import _ClassificationTaskWrapper
from metrics_b200.classification.stat_scores import BinaryStatScores, MulticlassStatScores, MultilabelStatScores, _dispatch
from metrics_b200.functional.classification.ratio_metrics import _RATIOS, _ratio_reduce
from metrics_b200.metric import Metric

_CLASS_STEM = {
    "precision": "Precision",
    "recall": "Recall",
    "specificity": "Specificity",
    "negative_predictive_value": "NegativePredictiveValue",
    "hamming_distance": "HammingDistance",
}
# which classes forward `zero_division` at compute time (reference: only precision / recall do)
_CLASS_ZERO_DIVISION = {"precision", "recall"}


def _family(kind: str):
    # `higher_is_better` as the reference declares it (None for specificity / NPV: specificity.py:99, negative_predictive_value.py:99)
    higher = {"hamming_distance": False, "specificity": None, "negative_predictive_value": None}.get(kind, True)
    stem = _CLASS_STEM[kind]
    attrs = {
        "is_differentiable": False,
        "higher_is_better": higher,
        "full_state_update": False,
        "plot_lower_bound": 0.0,
        "plot_upper_bound": 1.0,
    }

    def zd(self) -> float:
        return self.zero_division if kind in _CLASS_ZERO_DIVISION else 0

    def compute_binary(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _ratio_reduce(kind, tp, fp, tn, fn, "binary", self.multidim_average, zero_division=zd(self))

    def compute_multiclass(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _ratio_reduce(kind, tp, fp, tn, fn, self.average, self.multidim_average, top_k=self.top_k,
                             zero_division=zd(self))

    def compute_multilabel(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _ratio_reduce(kind, tp, fp, tn, fn, self.average, self.multidim_average, multilabel=True,
                             zero_division=zd(self))

    ref = _RATIOS[kind].reference
    b = type(f"Binary{stem}", (BinaryStatScores,), {**attrs, "compute": compute_binary, "__module__": __name__,
                                                   "__doc__": f"Binary {kind.replace('_', ' ')} (reduce: functional {ref})."})
    mc = type(f"Multiclass{stem}", (MulticlassStatScores,), {**attrs, "plot_legend_name": "Class", "compute": compute_multiclass,
                                                         "__module__": __name__,
                                                         "__doc__": f"Multiclass {kind.replace('_', ' ')} (reduce: functional {ref})."})
    ml = type(f"Multilabel{stem}", (MultilabelStatScores,), {**attrs, "plot_legend_name": "Label", "compute": compute_multilabel,
                                                         "__module__": __name__,
                                                         "__doc__": f"Multilabel {kind.replace('_', ' ')} (reduce: functional {ref})."})

    def __new__(cls, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,
                num_classes: Optional[int] = None, num_labels: Optional[int] = None,
                average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
                multidim_average: Optional[Literal["global", "samplewise"]] = "global", top_k: Optional[int] = 1,
                ignore_index: Optional[int] = None, validate_args: bool = True, **kwargs: Any) -> Metric:
        assert multidim_average is not None  # noqa: S101
        kwargs.update({"multidim_average": multidim_average, "ignore_index": ignore_index, "validate_args": validate_args})
        return _dispatch(b, mc, ml, task, threshold, num_classes, num_labels, average, top_k, kwargs)

    wrapper = type(stem, (_ClassificationTaskWrapper,), {"__new__": __new__, "__module__": __name__,
                                                         "__doc__": f"Task wrapper for {kind.replace('_', ' ')}."})
    return b, mc, ml, wrapper


BinaryPrecision, MulticlassPrecision, MultilabelPrecision, Precision = _family("precision")
BinaryRecall, MulticlassRecall, MultilabelRecall, Recall = _family("recall")
BinarySpecificity, MulticlassSpecificity, MultilabelSpecificity, Specificity = _family("specificity")
(BinaryNegativePredictiveValue, MulticlassNegativePredictiveValue, MultilabelNegativePredictiveValue,
 NegativePredictiveValue) = _family("negative_predictive_value")
BinaryHammingDistance, MulticlassHammingDistance, MultilabelHammingDistance, HammingDistance = _family("hamming_distance")
