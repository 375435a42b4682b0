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
