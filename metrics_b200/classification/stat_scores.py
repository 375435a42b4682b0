"""Stat-scores metric classes (reference: classification/stat_scores.py)."""
from __future__ import annotations

from typing import Any, List, Optional, Union

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification.stat_scores import (
    _multiclass_stat_scores_arg_validation,
    _multiclass_stat_scores_compute,
    _multiclass_stat_scores_tensor_validation,
    _multiclass_stat_scores_update_,
    _require_kernel_mode,
    stat_scores_workspace,
)
from metrics_b200.metric import Metric
from metrics_b200.utilities.data import dim_zero_cat


class _AbstractStatScores(Metric):
    """Holds the four counters ``tp, fp, tn, fn`` (reference :43-88): int64 tensors with ``sum`` reduction, or
    lists with ``cat`` reduction when ``multidim_average="samplewise"``."""

    tp: Union[List[Tensor], Tensor]
    fp: Union[List[Tensor], Tensor]
    tn: Union[List[Tensor], Tensor]
    fn: Union[List[Tensor], Tensor]

    def _create_state(self, size: int, multidim_average: str = "global") -> None:
        for name in ("tp", "fp", "tn", "fn"):
            if multidim_average == "samplewise":
                self.add_state(name, [], dist_reduce_fx="cat")
            else:
                self.add_state(name, torch.zeros(size, dtype=torch.long), dist_reduce_fx="sum")
        self._scratch: Optional[Tensor] = None

    def _workspace(self, n_slots: int, device: torch.device) -> Tensor:
        """Per-instance kernel scratch (NOT a metric state: never synced, saved or compared)."""
        ws = self._scratch
        if ws is None or ws.device != device or ws.numel() != 3 * n_slots + 2:
            ws = stat_scores_workspace(n_slots, device)
            self._scratch = ws
        return ws

    def _update_state(self, tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor) -> None:
        if self.multidim_average == "samplewise":
            self.tp.append(tp)
            self.fp.append(fp)
            self.tn.append(tn)
            self.fn.append(fn)
        else:
            self.tp += tp
            self.fp += fp
            self.tn += tn
            self.fn += fn

    def _final_state(self) -> tuple[Tensor, Tensor, Tensor, Tensor]:
        return dim_zero_cat(self.tp), dim_zero_cat(self.fp), dim_zero_cat(self.tn), dim_zero_cat(self.fn)


class MulticlassStatScores(_AbstractStatScores):
    """tp / fp / tn / fn / support for multiclass tasks (reference :198-352)."""

    is_differentiable: bool = False
    higher_is_better: Optional[bool] = None
    full_state_update: bool = False

    def __init__(
        self,
        num_classes: Optional[int] = None,
        top_k: int = 1,
        average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
        multidim_average: Literal["global", "samplewise"] = "global",
        ignore_index: Optional[int] = None,
        validate_args: bool = True,
        **kwargs: Any,
    ) -> None:
        zero_division = kwargs.pop("zero_division", 0)
        super(_AbstractStatScores, self).__init__(**kwargs)
        if validate_args:
            _multiclass_stat_scores_arg_validation(num_classes, top_k, average, multidim_average, ignore_index, zero_division)
        _require_kernel_mode(top_k, multidim_average)
        self.num_classes = num_classes
        self.top_k = top_k
        self.average = average
        self.multidim_average = multidim_average
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self.zero_division = zero_division
        self._create_state(
            size=1 if (average == "micro" and top_k == 1) else (num_classes or 1), multidim_average=multidim_average
        )

    def update(self, preds: Tensor, target: Tensor) -> None:
        if self.validate_args:
            _multiclass_stat_scores_tensor_validation(
                preds, target, self.num_classes, self.multidim_average, self.ignore_index
            )
        num_classes = self.num_classes if self.num_classes is not None else 1
        _multiclass_stat_scores_update_(
            self.tp, self.fp, self.tn, self.fn, self._workspace(num_classes, self.tp.device), preds, target,
            num_classes, self.top_k, self.average, self.multidim_average, self.ignore_index, self.validate_args,
        )

    def compute(self) -> Tensor:
        tp, fp, tn, fn = self._final_state()
        return _multiclass_stat_scores_compute(tp, fp, tn, fn, self.average, self.multidim_average)
