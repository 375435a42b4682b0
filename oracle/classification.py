"""Oracle for the confusion-matrix / stat-scores family (numpy).  TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

All inputs are numpy arrays (bf16/f16 tensors are widened to float32 by the caller: exact).
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def argmax_dim1(scores: np.ndarray) -> np.ndarray:
    """`preds.argmax(dim=1)` (functional/classification/confusion_matrix.py:310, stat_scores.py:341).

    torch.argmax semantics, restated explicitly: the FIRST index holding the maximum; a NaN counts as the maximum
    (first NaN wins); -0.0 == +0.0.
    """
    isnan = np.isnan(scores)
    has_nan = isnan.any(axis=1)
    first_nan = isnan.argmax(axis=1)
    with np.errstate(invalid="ignore"):
        plain = np.where(isnan, -np.inf, scores).argmax(axis=1)  # numpy argmax returns the first maximum
    return np.where(has_nan, first_nan, plain).astype(np.int64)


def multiclass_format(preds: np.ndarray, target: np.ndarray, ignore_index: Optional[int]):
    """_multiclass_confusion_matrix_format (confusion_matrix.py:297-321): argmax if preds has a class dim,
    flatten both, drop rows whose target == ignore_index."""
    if preds.ndim == target.ndim + 1:
        preds = argmax_dim1(preds)
    preds = preds.reshape(-1).astype(np.int64)
    target = target.reshape(-1).astype(np.int64)
    if ignore_index is not None:
        keep = target != ignore_index
        preds, target = preds[keep], target[keep]
    return preds, target


def multiclass_confusion_matrix(preds, target, num_classes: int, ignore_index: Optional[int] = None) -> np.ndarray:
    """_multiclass_confusion_matrix_update (confusion_matrix.py:324-328) via _bincount (utilities/data.py:199-206):
    bincount(target * C + preds, minlength=C*C).reshape(C, C)."""
    p, t = multiclass_format(preds, target, ignore_index)
    bins = np.bincount(t * num_classes + p, minlength=num_classes * num_classes)
    return bins.reshape(num_classes, num_classes).astype(np.int64)


def multiclass_stat_scores(preds, target, num_classes: int, average: Optional[str] = "macro",
                           ignore_index: Optional[int] = None):
    """_multiclass_stat_scores_update, top_k == 1 and multidim_average == "global" (stat_scores.py:424-448)."""
    p, t = multiclass_format(preds, target, ignore_index)
    if average == "micro":  # :424-434
        tp = np.int64((p == t).sum())
        fp = np.int64((p != t).sum())
        fn = np.int64((p != t).sum())
        tn = np.int64(num_classes * p.size - (fp + fn + tp))
        return tp, fp, tn, fn
    cm = np.bincount(t * num_classes + p, minlength=num_classes**2).reshape(num_classes, num_classes).astype(np.int64)
    tp = np.diag(cm).copy()  # :445-448
    fp = cm.sum(0) - tp
    fn = cm.sum(1) - tp
    tn = cm.sum() - (fp + fn + tp)
    return tp, fp, tn, fn


def _safe_divide_f32(num, denom, zero_division: float = 0.0) -> np.ndarray:
    """_safe_divide (utilities/compute.py:47-68): .float() then divide, zero_division where denom == 0."""
    num = np.asarray(num, dtype=np.float32)
    denom = np.asarray(denom, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = num / denom
    return np.where(denom != 0, out, np.float32(zero_division)).astype(np.float32)


def _adjust_weights_safe_divide(score, average, multilabel, tp, fp, fn, top_k=1) -> np.ndarray:
    """_adjust_weights_safe_divide (utilities/compute.py:71-82)."""
    if average is None or average == "none":
        return score
    if average == "weighted":
        weights = (tp + fn).astype(np.float32)  # int64 * f32 -> f32 in torch
    else:
        weights = np.ones_like(score, dtype=np.float32)
        if not multilabel:
            weights[(tp + fp + fn == 0) if top_k == 1 else (tp + fn == 0)] = 0.0
    prod = (weights * score).astype(np.float32)
    return _safe_divide_f32(prod, weights.sum(-1, keepdims=True, dtype=np.float32)).sum(-1, dtype=np.float32)


def accuracy_reduce(tp, fp, tn, fn, average) -> np.ndarray:
    """_accuracy_reduce, multiclass/global branches (functional/classification/accuracy.py:37-88)."""
    if average == "micro":
        tp, fn = np.sum(tp), np.sum(fn)
        return _safe_divide_f32(tp, tp + fn)
    score = _safe_divide_f32(tp, tp + fn)
    return _adjust_weights_safe_divide(score, average, False, tp, fp, fn)


def fbeta_reduce(tp, fp, tn, fn, beta: float, average, zero_division: float = 0.0) -> np.ndarray:
    """_fbeta_reduce, multiclass/global branches (functional/classification/f_beta.py:37-58).
    torch: (1 + beta2) * int64 tensor -> float32 tensor (python float scalar promotes int tensors to f32)."""
    beta2 = beta**2
    if average == "micro":
        tp, fn, fp = np.sum(tp), np.sum(fn), np.sum(fp)
    f = np.float32
    tpf, fnf, fpf = np.asarray(tp, dtype=f), np.asarray(fn, dtype=f), np.asarray(fp, dtype=f)
    num = f(1 + beta2) * tpf
    den = f(1 + beta2) * tpf + f(beta2) * fnf + fpf
    score = _safe_divide_f32(num, den, zero_division)
    if average == "micro":
        return score
    return _adjust_weights_safe_divide(score, average, False, tp, fp, fn)


# ---------------------------------------------------------------------------------------------------------
# binary / multilabel stat scores and confusion matrices
# ---------------------------------------------------------------------------------------------------------
def _sigmoid_if_logits(preds: np.ndarray) -> np.ndarray:
    """normalize_logits_if_needed(preds, "sigmoid") (utilities/compute.py:190-229), fp32 math."""
    with np.errstate(invalid="ignore"):
        cond = (preds < 0).any() or (preds > 1).any()
    if not cond:
        return preds
    x = preds.astype(np.float32)
    return (1.0 / (1.0 + np.exp(-x, dtype=np.float32))).astype(np.float32)


def binary_stat_scores(preds, target, threshold: float = 0.5, ignore_index: Optional[int] = None, samplewise: bool = False):
    """_binary_stat_scores_format + _update (stat_scores.py:95-134)."""
    if np.issubdtype(preds.dtype, np.floating):
        preds = _sigmoid_if_logits(preds) > threshold
    p = preds.reshape(preds.shape[0], -1).astype(np.int64)
    t = target.reshape(target.shape[0], -1).astype(np.int64).copy()
    if ignore_index is not None:
        t[t == ignore_index] = -1
    axis = (1,) if samplewise else (0, 1)
    tp = ((t == p) & (t == 1)).sum(axis)
    fn = ((t != p) & (t == 1)).sum(axis)
    fp = ((t != p) & (t == 0)).sum(axis)
    tn = ((t == p) & (t == 0)).sum(axis)
    return tp, fp, tn, fn


def multilabel_stat_scores(preds, target, num_labels: int, threshold: float = 0.5, ignore_index: Optional[int] = None,
                           samplewise: bool = False):
    """_multilabel_stat_scores_format + _update (stat_scores.py:681-714)."""
    if np.issubdtype(preds.dtype, np.floating):
        preds = _sigmoid_if_logits(preds) > threshold
    p = preds.reshape(preds.shape[0], preds.shape[1], -1).astype(np.int64)
    t = target.reshape(target.shape[0], target.shape[1], -1).astype(np.int64).copy()
    if ignore_index is not None:
        t[t == ignore_index] = -1
    axis = (2,) if samplewise else (0, 2)
    tp = ((t == p) & (t == 1)).sum(axis)
    fn = ((t != p) & (t == 1)).sum(axis)
    fp = ((t != p) & (t == 0)).sum(axis)
    tn = ((t == p) & (t == 0)).sum(axis)
    return tp, fp, tn, fn


def confmat_from_counts(tp, fp, tn, fn) -> np.ndarray:
    """[[tn, fp], [fn, tp]] — what bincount(2*target + preds, 4).reshape(2, 2) yields (confusion_matrix.py:148-152, 511-516)."""
    return np.stack([np.stack([tn, fp], -1), np.stack([fn, tp], -1)], -2).astype(np.int64)


def binary_groups_stat_scores(preds, target, groups, threshold: float = 0.5, ignore_index: Optional[int] = None) -> np.ndarray:
    """_binary_groups_stat_scores (group_fairness.py:52-83): sort the samples by group id, split at the id changes, binary
    stat scores per piece.  Returns ``[G, 4]`` (tp, fp, tn, fn) for the G ids present, ascending."""
    if np.issubdtype(preds.dtype, np.floating):
        preds = _sigmoid_if_logits(preds) > threshold  # the vote is taken over the WHOLE batch, before the split (:66)
    preds = preds.reshape(preds.shape[0], -1)
    target = target.reshape(target.shape[0], -1)
    ids = groups.reshape(-1)
    out = []
    for gid in np.unique(ids):
        sel = ids == gid
        out.append(np.array([int(np.sum(x)) for x in binary_stat_scores(preds[sel].astype(np.int64), target[sel], threshold, ignore_index)]))
    return np.stack(out)


def fairness_ratios(counts: np.ndarray) -> dict:
    """_compute_binary_demographic_parity / _equal_opportunity (group_fairness.py:164-174, :243-255) in float32 like the
    reference's `_safe_divide`: lowest over highest positive rate / true-positive rate, keyed by the two group indices."""
    tp, fp, tn, fn = (counts[:, i].astype(np.float32) for i in range(4))
    out = {}
    for tag, rates in (("DP", _safe_divide_f32(tp + fp, tp + fp + tn + fn)), ("EO", _safe_divide_f32(tp, tp + fn))):
        lo, hi = int(np.argmin(rates)), int(np.argmax(rates))
        out[f"{tag}_{lo}_{hi}"] = _safe_divide_f32(rates[lo], rates[hi])
    return out

