"""The reference's execution of the hot path, restated op for op on torch tensors.  TEST/BENCH INFRASTRUCTURE.

The reference is pure Python over stock ATen ops, so "the reference's CPU implementation" of the confusion-matrix
update IS this op chain; it cannot travel to the GPU box (/root/reference is absent there), hence this port.  The functions
are device-agnostic: on CPU tensors they are the reference's CPU path, on CUDA tensors the stock-ATen-on-B200 chain the
reference would execute there (SURVEY.md §2.2's bar, and the arbiter when the reference's CPU and CUDA results differ by
1-ulp sigmoid / softmax ties).  Used only by bench.py (`cpu_baseline`, `--impl reference`, `aten_gpu_baseline` when
baseline/_ref is absent) and by tests.  Each line cites what it restates.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor


def multiclass_confmat_update_cpu(confmat: Tensor, preds: Tensor, target: Tensor, num_classes: int,
                                  ignore_index: Optional[int] = None, validate_args: bool = False) -> None:
    """MulticlassConfusionMatrix.update on CPU tensors (classification/confusion_matrix.py:280-286)."""
    if validate_args:  # functional/classification/confusion_matrix.py:287-294 (content check = unique + len)
        check = num_classes if ignore_index is None else num_classes + 1
        if len(torch.unique(target)) > check:
            raise RuntimeError("Detected more unique values in `target` than expected.")
    if preds.ndim == target.ndim + 1:  # :309-310
        preds = preds.argmax(dim=1)
    preds = preds.flatten()  # :312
    target = target.flatten()  # :313
    if ignore_index is not None:  # :315-319
        keep = target != ignore_index
        preds, target = preds[keep], target[keep]
    unique_mapping = target.to(torch.long) * num_classes + preds.to(torch.long)  # :326
    bins = torch.bincount(unique_mapping, minlength=num_classes**2)  # utilities/data.py:206
    confmat += bins.reshape(num_classes, num_classes)  # :328 and classification/confusion_matrix.py:286


def binary_auroc_ap_compute_cpu(preds: Tensor, target: Tensor):
    """BinaryAUROC.compute + BinaryAveragePrecision.compute on CPU tensors, op for op: each metric runs its own
    `_binary_clf_curve` (functional/classification/precision_recall_curve.py:30-82) — two full sorts per collection
    compute (roc.py:53 and precision_recall_curve.py:275)."""
    import torch.nn.functional as F

    def clf_curve(p: Tensor, t: Tensor):
        idx = torch.argsort(p, descending=True)  # :60
        p, t = p[idx], t[idx]  # :62-63
        distinct = torch.where(p[1:] - p[:-1])[0]  # :70
        thr_idx = F.pad(distinct, [0, 1], value=t.size(0) - 1)  # :71
        t = (t == 1).to(torch.long)  # :72
        tps = torch.cumsum(t * 1.0, dim=0)[thr_idx]  # :73
        fps = 1 + thr_idx - tps  # :80
        return fps, tps, p[thr_idx]

    fps, tps, _ = clf_curve(preds, target)  # roc.py:53-78
    tps = torch.cat([torch.zeros(1, dtype=tps.dtype, device=tps.device), tps])
    fps = torch.cat([torch.zeros(1, dtype=fps.dtype, device=fps.device), fps])
    fpr, tpr = fps / fps[-1], tps / tps[-1]
    auroc = torch.trapz(tpr, fpr)  # utilities/compute.py:101-109
    fps, tps, _ = clf_curve(preds, target)  # precision_recall_curve.py:275-290
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    precision = torch.cat([precision.flip(0), torch.ones(1, device=preds.device)])
    recall = torch.cat([recall.flip(0), torch.zeros(1, device=preds.device)])
    ap = -torch.sum((recall[1:] - recall[:-1]) * precision[:-1])  # average_precision.py:74-75
    return auroc, ap


def multiclass_stat_scores_update_cpu(tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor, preds: Tensor, target: Tensor,
                                      num_classes: int) -> None:
    """MulticlassStatScores.update on CPU tensors, top_k=1 / global / macro-style states (functional/classification/
    stat_scores.py:328-344 format, :435-448 bincount path; classification/stat_scores.py:69-80 state add)."""
    if preds.ndim == target.ndim + 1:  # :341
        preds = preds.argmax(dim=1)
    preds, target = preds.flatten(), target.flatten()  # :342-343
    unique_mapping = target.to(torch.long) * num_classes + preds.to(torch.long)  # :441
    confmat = torch.bincount(unique_mapping, minlength=num_classes**2).reshape(num_classes, num_classes)  # :442-443
    d_tp = confmat.diag()  # :444
    d_fp = confmat.sum(0) - d_tp  # :445
    d_fn = confmat.sum(1) - d_tp  # :446
    d_tn = confmat.sum() - (d_fp + d_fn + d_tp)  # :447
    tp += d_tp
    fp += d_fp
    tn += d_tn
    fn += d_fn


def macro_accuracy_cpu(tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor) -> Tensor:
    """_accuracy_reduce(average="macro") (functional/classification/accuracy.py:84-88 + utilities/compute.py:71-82)."""
    score = torch.where(tp + fn != 0, tp.float() / (tp + fn).float(), torch.zeros((), device=tp.device))
    weights = torch.ones_like(score)
    weights[tp + fp + fn == 0] = 0.0
    return (weights * score / weights.sum()).sum()


def multiclass_auroc_compute_cpu(preds: Tensor, target: Tensor, num_classes: int) -> Tensor:
    """MulticlassAUROC.compute(average="macro") on CPU tensors: the reference's Python loop over classes, one
    `_binary_clf_curve` (argsort + cumsum) per class (roc.py:176-181, auroc.py:193-205)."""
    import torch.nn.functional as F

    aucs = []
    for c in range(num_classes):
        p, t = preds[:, c], target
        idx = torch.argsort(p, descending=True)
        p, t = p[idx], t[idx]
        distinct = torch.where(p[1:] - p[:-1])[0]
        thr_idx = F.pad(distinct, [0, 1], value=t.size(0) - 1)
        t = (t == c).to(torch.long)
        tps = torch.cumsum(t * 1.0, dim=0)[thr_idx]
        fps = 1 + thr_idx - tps
        tps = torch.cat([torch.zeros(1, dtype=tps.dtype, device=tps.device), tps])
        fps = torch.cat([torch.zeros(1, dtype=fps.dtype, device=fps.device), fps])
        fpr = fps / fps[-1] if fps[-1] > 0 else torch.zeros_like(fps)
        tpr = tps / tps[-1] if tps[-1] > 0 else torch.zeros_like(tps)
        aucs.append(torch.trapz(tpr, fpr))
    return torch.stack(aucs).mean()


# ----------------------------------------------------------------------------------------------------------------------
# exact-mode curve functionals (thresholds=None), format + compute, any device
# ----------------------------------------------------------------------------------------------------------------------
def normalize_logits_if_needed_chain(t: Tensor, normalization: str) -> Tensor:
    """utilities/compute.py:190-229: host-synchronising branch on CPU (:218-221), `torch.where` branch on device (:223-229)."""
    if t.device.type == "cpu":
        if not torch.all((t >= 0) * (t <= 1)):
            t = t.sigmoid() if normalization == "sigmoid" else torch.softmax(t, dim=1)
        return t
    condition = ((t < 0) | (t > 1)).any()
    return torch.where(condition, torch.sigmoid(t) if normalization == "sigmoid" else torch.softmax(t, dim=1), t)


def binary_clf_curve_chain(preds: Tensor, target: Tensor, pos_label: int = 1):
    """_binary_clf_curve without sample weights (functional/classification/precision_recall_curve.py:30-82)."""
    import torch.nn.functional as F

    idx = torch.argsort(preds, descending=True)  # :60
    preds, target = preds[idx], target[idx]  # :62-63
    distinct = torch.where(preds[1:] - preds[:-1])[0]  # :70
    thr_idx = F.pad(distinct, [0, 1], value=target.size(0) - 1)  # :71
    target = (target == pos_label).to(torch.long)  # :72
    tps = torch.cumsum(target * 1.0, dim=0)[thr_idx]  # :73
    fps = 1 + thr_idx - tps  # :80
    return fps, tps, preds[thr_idx]


def binary_roc_compute_chain(preds: Tensor, target: Tensor, pos_label: int = 1):
    """_binary_roc_compute, exact mode (roc.py:53-78)."""
    fps, tps, thres = binary_clf_curve_chain(preds, target, pos_label)
    tps = torch.cat([torch.zeros(1, dtype=tps.dtype, device=tps.device), tps])
    fps = torch.cat([torch.zeros(1, dtype=fps.dtype, device=fps.device), fps])
    thres = torch.cat([torch.ones(1, dtype=thres.dtype, device=thres.device), thres])
    fpr = torch.zeros_like(thres) if fps[-1] <= 0 else fps / fps[-1]
    tpr = torch.zeros_like(thres) if tps[-1] <= 0 else tps / tps[-1]
    return fpr, tpr, thres


def binary_prc_compute_chain(preds: Tensor, target: Tensor, pos_label: int = 1):
    """_binary_precision_recall_curve_compute, exact mode (precision_recall_curve.py:275-290)."""
    fps, tps, thresholds = binary_clf_curve_chain(preds, target, pos_label)
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    if (target == 0).all():
        recall = torch.ones_like(recall)
    precision = torch.cat([precision.flip(0), torch.ones(1, dtype=precision.dtype, device=precision.device)])
    recall = torch.cat([recall.flip(0), torch.zeros(1, dtype=recall.dtype, device=recall.device)])
    return precision, recall, thresholds.flip(0).detach().clone()


def _binary_format_chain(preds: Tensor, target: Tensor, ignore_index: Optional[int]):
    """_binary_precision_recall_curve_format (precision_recall_curve.py:153-178)."""
    preds, target = preds.flatten(), target.flatten()
    if ignore_index is not None:
        keep = target != ignore_index
        preds, target = preds[keep], target[keep]
    return normalize_logits_if_needed_chain(preds, "sigmoid"), target


def _multiclass_format_chain(preds: Tensor, target: Tensor, num_classes: int, ignore_index: Optional[int]):
    """_multiclass_precision_recall_curve_format, average=None (precision_recall_curve.py:446-462)."""
    preds = preds.transpose(0, 1).reshape(num_classes, -1).T
    target = target.flatten()
    if ignore_index is not None:
        keep = target != ignore_index
        preds, target = preds[keep], target[keep]
    return normalize_logits_if_needed_chain(preds, "softmax"), target


def exact_curve_functional_chain(fn: str, preds: Tensor, target: Tensor, num_classes: Optional[int] = None,
                                 ignore_index: Optional[int] = None):
    """binary_roc / binary_precision_recall_curve / multiclass_roc / multiclass_precision_recall_curve with thresholds=None,
    average=None: the reference's format + per-class compute loop (roc.py:176-181, precision_recall_curve.py:565-569)."""
    if fn in ("binary_roc", "binary_precision_recall_curve"):
        p, t = _binary_format_chain(preds, target, ignore_index)
        return binary_roc_compute_chain(p, t) if fn == "binary_roc" else binary_prc_compute_chain(p, t)
    if fn in ("multiclass_roc", "multiclass_precision_recall_curve"):
        p, t = _multiclass_format_chain(preds, target, num_classes, ignore_index)
        one = binary_roc_compute_chain if fn == "multiclass_roc" else binary_prc_compute_chain
        cols = [one(p[:, i], t, pos_label=i) for i in range(num_classes)]
        return [c[0] for c in cols], [c[1] for c in cols], [c[2] for c in cols]
    raise KeyError(fn)
