"""The reference's CPU execution of the hot path, restated op for op on torch CPU tensors.  TEST/BENCH INFRASTRUCTURE.

The reference is pure Python over stock ATen ops, so "the reference's CPU implementation" of the confusion-matrix
update IS this op chain; it cannot travel to the GPU box (/root/reference is absent there), hence this port.
Used only by bench.py (`cpu_baseline` leg and `--impl reference`) and by tests that cross-check it against the numpy
oracle.  Each line cites what it restates.
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
    tps = torch.cat([torch.zeros(1, dtype=tps.dtype), tps])
    fps = torch.cat([torch.zeros(1, dtype=fps.dtype), fps])
    fpr, tpr = fps / fps[-1], tps / tps[-1]
    auroc = torch.trapz(tpr, fpr)  # utilities/compute.py:101-109
    fps, tps, _ = clf_curve(preds, target)  # precision_recall_curve.py:275-290
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    precision = torch.cat([precision.flip(0), torch.ones(1)])
    recall = torch.cat([recall.flip(0), torch.zeros(1)])
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
    score = torch.where(tp + fn != 0, tp.float() / (tp + fn).float(), torch.zeros(()))
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
        tps = torch.cat([torch.zeros(1, dtype=tps.dtype), tps])
        fps = torch.cat([torch.zeros(1, dtype=fps.dtype), fps])
        fpr = fps / fps[-1] if fps[-1] > 0 else torch.zeros_like(fps)
        tpr = tps / tps[-1] if tps[-1] > 0 else torch.zeros_like(tps)
        aucs.append(torch.trapz(tpr, fpr))
    return torch.stack(aucs).mean()
