"""Ratio metrics over the stat-scores counters: precision, recall, specificity, negative predictive value, Hamming distance.

Reference: functional/classification/{precision_recall,specificity,negative_predictive_value,hamming}.py.  All of them
are "one ratio of tp/fp/tn/fn counters + a class average" on top of the K1b / K2 kernels; they need no device code of
their own, so they are generated here from one table instead of five near-identical modules.
"""
from __future__ import annotations

from typing import Callable, Dict, NamedTuple, Optional

from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.stat_scores import (
    _binary_stat_scores_arg_validation,
    _binary_stat_scores_tensor_validation,
    _binary_stat_scores_update,
    _multiclass_stat_scores_arg_validation,
    _multiclass_stat_scores_states,
    _multiclass_stat_scores_tensor_validation,
    _multilabel_stat_scores_arg_validation,
    _multilabel_stat_scores_tensor_validation,
    _multilabel_stat_scores_update,
)
from metrics_b200.utilities.compute import _adjust_weights_safe_divide, _safe_divide


class _Ratio(NamedTuple):
    """score = post(num / den).  ``micro_sums``: which counters take part (they are summed over classes first)."""

    num: Callable[[Tensor, Tensor, Tensor, Tensor, bool], Tensor]  # (tp, fp, tn, fn, multilabel) -> numerator
    den: Callable[[Tensor, Tensor, Tensor, Tensor, bool], Tensor]
    complement: bool  # Hamming distance = 1 - accuracy-like ratio
    uses_zero_division: bool
    uses_top_k: bool
    reference: str


_RATIOS: Dict[str, _Ratio] = {
    # precision_recall.py:37-60
    "precision": _Ratio(lambda tp, fp, tn, fn, ml: tp, lambda tp, fp, tn, fn, ml: tp + fp, False, True, True,
                        "precision_recall.py:37-60"),
    "recall": _Ratio(lambda tp, fp, tn, fn, ml: tp, lambda tp, fp, tn, fn, ml: tp + fn, False, True, True,
                     "precision_recall.py:37-60"),
    # specificity.py:37-54
    "specificity": _Ratio(lambda tp, fp, tn, fn, ml: tn, lambda tp, fp, tn, fn, ml: tn + fp, False, False, False,
                          "specificity.py:37-54"),
    # negative_predictive_value.py:37-56
    "negative_predictive_value": _Ratio(lambda tp, fp, tn, fn, ml: tn, lambda tp, fp, tn, fn, ml: tn + fn, False, True, True,
                                        "negative_predictive_value.py:37-56"),
    # hamming.py:37-83: multiclass uses tp / (tp + fn), binary and multilabel (tp + tn) / all
    "hamming_distance": _Ratio(lambda tp, fp, tn, fn, ml: tp + tn if ml else tp,
                               lambda tp, fp, tn, fn, ml: tp + tn + fp + fn if ml else tp + fn, True, False, False,
                               "hamming.py:37-83"),
}


def _ratio_reduce(
    kind: str,
    tp: Tensor,
    fp: Tensor,
    tn: Tensor,
    fn: Tensor,
    average: Optional[str],
    multidim_average: str = "global",
    multilabel: bool = False,
    top_k: int = 1,
    zero_division: float = 0,
) -> Tensor:
    """Counters -> score.  ``average="binary"`` treats the (scalar / per-sample) counters as one problem."""
    spec = _RATIOS[kind]
    zd = zero_division if spec.uses_zero_division else 0
    elementwise_all = multilabel or average == "binary"  # Hamming: binary behaves like one multilabel label

    def score_of(a: Tensor, b: Tensor, c: Tensor, d: Tensor) -> Tensor:
        s = _safe_divide(spec.num(a, b, c, d, elementwise_all), spec.den(a, b, c, d, elementwise_all), zd)
        return 1 - s if spec.complement else s

    if average == "binary":
        return score_of(tp, fp, tn, fn)
    if average == "micro":
        axis = 0 if multidim_average == "global" else 1
        return score_of(tp.sum(dim=axis), fp.sum(dim=axis), tn.sum(dim=axis), fn.sum(dim=axis))
    score = score_of(tp, fp, tn, fn)
    if spec.uses_top_k:
        return _adjust_weights_safe_divide(score, average, multilabel, tp, fp, fn, top_k=top_k)
    return _adjust_weights_safe_divide(score, average, multilabel, tp, fp, fn)


def _publish(fn: Callable, spec: _Ratio) -> Callable:
    """Introspection parity: the families without a `zero_division` argument in the reference (specificity, Hamming
    distance) do not list the shared implementation's one in `inspect.signature`."""
    if not spec.uses_zero_division:
        import inspect

        sig = inspect.signature(fn)
        fn.__signature__ = sig.replace(parameters=[p for p in sig.parameters.values() if p.name != "zero_division"])
    return fn


def _make_binary(kind: str) -> Callable:
    spec = _RATIOS[kind]

    def fn_(preds: Tensor, target: Tensor, threshold: float = 0.5,
            multidim_average: Literal["global", "samplewise"] = "global", ignore_index: Optional[int] = None,
            validate_args: bool = True, zero_division: float = 0) -> Tensor:
        if validate_args:
            _binary_stat_scores_arg_validation(threshold, multidim_average, ignore_index, zero_division)
            _binary_stat_scores_tensor_validation(preds, target, multidim_average, ignore_index)
        tp, fp, tn, fn = _binary_stat_scores_update(preds, target, threshold, multidim_average, ignore_index, validate_args)
        return _ratio_reduce(kind, tp, fp, tn, fn, "binary", multidim_average, zero_division=zero_division)

    fn_.__name__ = fn_.__qualname__ = f"binary_{kind}"
    fn_.__doc__ = f"Binary {kind.replace('_', ' ')} from ONE counting kernel (reference functional/classification/{spec.reference})."
    return _publish(fn_, spec)


def _make_multiclass(kind: str) -> Callable:
    spec = _RATIOS[kind]

    def fn_(preds: Tensor, target: Tensor, num_classes: int,
            average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro", top_k: int = 1,
            multidim_average: Literal["global", "samplewise"] = "global", ignore_index: Optional[int] = None,
            validate_args: bool = True, zero_division: float = 0) -> Tensor:
        if validate_args:
            _multiclass_stat_scores_arg_validation(num_classes, top_k, average, multidim_average, ignore_index, zero_division)
            _multiclass_stat_scores_tensor_validation(preds, target, num_classes, multidim_average, ignore_index)
        tp, fp, tn, fn = _multiclass_stat_scores_states(
            preds, target, num_classes, top_k, average, multidim_average, ignore_index, validate_args)
        return _ratio_reduce(kind, tp, fp, tn, fn, average, multidim_average, top_k=top_k, zero_division=zero_division)

    fn_.__name__ = fn_.__qualname__ = f"multiclass_{kind}"
    fn_.__doc__ = f"Multiclass {kind.replace('_', ' ')} from the fused argmax + counters kernel (reference {spec.reference})."
    return _publish(fn_, spec)


def _make_multilabel(kind: str) -> Callable:
    spec = _RATIOS[kind]

    def fn_(preds: Tensor, target: Tensor, num_labels: int, threshold: float = 0.5,
            average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
            multidim_average: Literal["global", "samplewise"] = "global", ignore_index: Optional[int] = None,
            validate_args: bool = True, zero_division: float = 0) -> Tensor:
        if validate_args:
            _multilabel_stat_scores_arg_validation(num_labels, threshold, average, multidim_average, ignore_index, zero_division)
            _multilabel_stat_scores_tensor_validation(preds, target, num_labels, multidim_average, ignore_index)
        tp, fp, tn, fn = _multilabel_stat_scores_update(
            preds, target, num_labels, threshold, multidim_average, ignore_index, validate_args)
        return _ratio_reduce(kind, tp, fp, tn, fn, average, multidim_average, multilabel=True, zero_division=zero_division)

    fn_.__name__ = fn_.__qualname__ = f"multilabel_{kind}"
    fn_.__doc__ = f"Multilabel {kind.replace('_', ' ')} from ONE counting kernel (reference {spec.reference})."
    return _publish(fn_, spec)


def _make_task(kind: str, b: Callable, mc: Callable, ml: Callable) -> Callable:
    def fn_(preds: Tensor, target: Tensor, task: Literal["binary", "multiclass", "multilabel"], threshold: float = 0.5,
            num_classes: Optional[int] = None, num_labels: Optional[int] = None,
            average: Optional[Literal["micro", "macro", "weighted", "none"]] = "micro",
            multidim_average: Optional[Literal["global", "samplewise"]] = "global", top_k: Optional[int] = 1,
            ignore_index: Optional[int] = None, validate_args: bool = True, zero_division: float = 0) -> Tensor:
        from metrics_b200.utilities.enums import ClassificationTask

        task_ = ClassificationTask.from_str(task)
        if task_ == ClassificationTask.BINARY:
            return b(preds, target, threshold, multidim_average, ignore_index, validate_args, zero_division)
        if task_ == ClassificationTask.MULTICLASS:
            if not isinstance(num_classes, int):
                raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
            if not isinstance(top_k, int):
                raise ValueError(f"`top_k` is expected to be `int` but `{type(top_k)} was passed.`")
            return mc(preds, target, num_classes, average, top_k, multidim_average, ignore_index, validate_args, zero_division)
        if not isinstance(num_labels, int):
            raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
        return ml(preds, target, num_labels, threshold, average, multidim_average, ignore_index, validate_args, zero_division)

    fn_.__name__ = fn_.__qualname__ = kind
    fn_.__doc__ = f"Task wrapper for {kind.replace('_', ' ')}."
    return _publish(fn_, _RATIOS[kind])


binary_precision, multiclass_precision, multilabel_precision = (
    _make_binary("precision"), _make_multiclass("precision"), _make_multilabel("precision"))
binary_recall, multiclass_recall, multilabel_recall = _make_binary("recall"), _make_multiclass("recall"), _make_multilabel("recall")
binary_specificity, multiclass_specificity, multilabel_specificity = (
    _make_binary("specificity"), _make_multiclass("specificity"), _make_multilabel("specificity"))
binary_negative_predictive_value, multiclass_negative_predictive_value, multilabel_negative_predictive_value = (
    _make_binary("negative_predictive_value"), _make_multiclass("negative_predictive_value"),
    _make_multilabel("negative_predictive_value"))
binary_hamming_distance, multiclass_hamming_distance, multilabel_hamming_distance = (
    _make_binary("hamming_distance"), _make_multiclass("hamming_distance"), _make_multilabel("hamming_distance"))
precision = _make_task("precision", binary_precision, multiclass_precision, multilabel_precision)
recall = _make_task("recall", binary_recall, multiclass_recall, multilabel_recall)
specificity = _make_task("specificity", binary_specificity, multiclass_specificity, multilabel_specificity)
negative_predictive_value = _make_task("negative_predictive_value", binary_negative_predictive_value,
                                       multiclass_negative_predictive_value, multilabel_negative_predictive_value)
hamming_distance = _make_task("hamming_distance", binary_hamming_distance, multiclass_hamming_distance,
                              multilabel_hamming_distance)
