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
