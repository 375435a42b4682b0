"""Confusion-matrix metric classes (reference: classification/confusion_matrix.py)."""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.confusion_matrix import (
    _multiclass_confusion_matrix_arg_validation,
    _multiclass_confusion_matrix_compute,
    _multiclass_confusion_matrix_tensor_validation,
    _multiclass_confusion_matrix_update_,
)
from metrics_b200.metric import Metric


class MulticlassConfusionMatrix(Metric):
    """``[C, C]`` int64 confusion matrix accumulated over batches (reference :191-290).

    State: ``confmat`` (``dist_reduce_fx="sum"``).  ``update`` launches one fused kernel that reads the batch once
    and adds into ``confmat`` in place; with ``validate_args=False`` it never synchronises the host.
    """

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = None
    full_state_update: bool = False
    confmat: Tensor

    def __init__(
        self,
        num_classes: int,
        ignore_index: Optional[int] = None,
        normalize: Optional[Literal["none", "true", "pred", "all"]] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        super().__init__(**kwargs)
        if validate_args:
            _multiclass_confusion_matrix_arg_validation(num_classes, ignore_index, normalize)
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.normalize = normalize
        self.validate_args = validate_args
        self.add_state("confmat", torch.zeros(num_classes, num_classes, dtype=torch.long), dist_reduce_fx="sum")

    def update(self, preds: Tensor, target: Tensor) -> None:
        if self.validate_args:
            _multiclass_confusion_matrix_tensor_validation(preds, target, self.num_classes, self.ignore_index)
        _multiclass_confusion_matrix_update_(
            self.confmat, preds, target, self.num_classes, self.ignore_index, self.validate_args
        )

    def compute(self) -> Tensor:
        return _multiclass_confusion_matrix_compute(self.confmat, self.normalize)
