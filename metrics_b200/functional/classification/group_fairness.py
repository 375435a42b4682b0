"""Group fairness for binary classification: per-group tp / fp / tn / fn, demographic parity and equal opportunity
(reference: functional/classification/group_fairness.py).

The reference sorts the batch by group id, splits it on the host (`.cpu().tolist()` of the group sizes, :71-81) and runs
the binary stat-scores op chain once per group.  Here ONE launch of the counting kernel K2 (csrc/binary.cu) produces the
``[G, 4]`` table: the batch is presented as a multilabel problem with one label per group, where sample ``n`` carries its
target under label ``groups[n]`` and the kernel's ignore value everywhere else.  The score → {0, 1} format step (sigmoid
vote, threshold, ``ignore_index``) is the kernel's own.  Counters are indexed by GROUP ID; the reference indexes them by
the rank of the id among the ids present in the batch (:71-83, classification/group_fairness.py:51-57), which is the same
thing whenever every group occurs in every batch and mis-files the counts otherwise.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200.functional.classification import _binary_counts as _bc
from metrics_b200.functional.classification.stat_scores import (
    _binary_stat_scores_arg_validation,
    _binary_stat_scores_tensor_validation,
)
from metrics_b200.utilities.compute import _safe_divide
from metrics_b200.utilities.prints import rank_zero_warn

_TASKS = ("demographic_parity", "equal_opportunity", "all")
GroupStats = Tuple[Tensor, Tensor, Tensor, Tensor]


def _groups_validation(groups: Tensor, num_groups: int) -> None:
    """Ids are ``0 .. num_groups-1`` in an int64 tensor (reference :30-44)."""
    if groups.numel() and int(groups.max()) > num_groups:
        raise ValueError(
            f"The largest number in the groups tensor is {int(groups.max())}, which is larger than the specified"
            f" number of groups {num_groups}. The group identifiers should be ``0, 1, ..., (num_groups - 1)``."
        )
    if groups.dtype != torch.long:
        raise ValueError(f"Expected dtype of argument groups to be long, not {groups.dtype}.")


def _groups_format(groups: Tensor) -> Tensor:
    """One id per sample, shaped ``[N, 1]`` (reference :47-49)."""
    return groups.reshape(groups.shape[0], -1)


def _check_task(task: str) -> None:
    if task not in _TASKS:
        raise ValueError(
            f"Expected argument `task` to either be ``demographic_parity``,``equal_opportunity`` or ``all`` but got {task}."
        )


def _group_counts(preds: Tensor, target: Tensor, groups: Tensor, num_groups: int, threshold: float = 0.5,
                  ignore_index: Optional[int] = None, validate_args: bool = True) -> Tensor:
    """``[num_groups, 4]`` int64 ``(tp, fp, tn, fn)`` per group id from ONE counting-kernel launch."""
    if validate_args:
        _binary_stat_scores_arg_validation(threshold, "global", ignore_index)
        _binary_stat_scores_tensor_validation(preds, target, "global", ignore_index)
        if groups.dtype != torch.long:
            raise ValueError(f"Expected dtype of argument groups to be long, not {groups.dtype}.")
    n = preds.shape[0]
    scores, labels = preds.reshape(n, 1, -1), target.reshape(n, 1, -1)
    if labels.is_floating_point():
        labels = labels.long()
    skip = -1 if ignore_index is None else ignore_index  # other groups' slots look like ignored targets to the kernel
    ids = torch.arange(num_groups, device=preds.device).reshape(1, num_groups, 1)
    per_group_target = torch.where(_groups_format(groups).reshape(n, 1, 1) == ids, labels, skip)
    per_group_scores = scores.expand(n, num_groups, scores.shape[2])
    return _bc.counts(per_group_scores, per_group_target, num_groups, threshold, skip, False, validate_args)


def _binary_groups_stat_scores(preds: Tensor, target: Tensor, groups: Tensor, num_groups: int, threshold: float = 0.5,
                               ignore_index: Optional[int] = None, validate_args: bool = True) -> List[GroupStats]:
    """Reference seam (:52-83): one ``(tp, fp, tn, fn)`` tuple of 0-d tensors per group PRESENT in the batch, in
    ascending id order — the reference's sort-and-split yields exactly that list, whatever ``num_groups`` says.  The ids
    are therefore replaced by their rank among the ids present (one `unique`, one host read of how many there are)."""
    if validate_args:
        _groups_validation(groups, num_groups)
    present, rank = torch.unique(groups, return_inverse=True)
    counts = _group_counts(preds, target, rank, int(present.numel()), threshold, ignore_index, validate_args)
    return [tuple(row.unbind(0)) for row in counts]


def _groups_reduce(group_stats: List[GroupStats]) -> Dict[str, Tensor]:
    """Rates: each group's four counters divided by their sum (reference :86-90)."""
    table = [torch.stack(stats) for stats in group_stats]
    return {f"group_{i}": row / row.sum() for i, row in enumerate(table)}


def _groups_stat_transform(group_stats: List[GroupStats]) -> Dict[str, Tensor]:
    """List of per-group tuples -> one ``[G]`` tensor per statistic (reference :93-102)."""
    return {name: torch.stack([stats[i] for stats in group_stats]) for i, name in enumerate(("tp", "fp", "tn", "fn"))}


def binary_groups_stat_rates(preds: Tensor, target: Tensor, groups: Tensor, num_groups: int, threshold: float = 0.5,
                             ignore_index: Optional[int] = None, validate_args: bool = True) -> Dict[str, Tensor]:
    """``{"group_i": [tp, fp, tn, fn] / total}`` (reference :105-161)."""
    return _groups_reduce(_binary_groups_stat_scores(preds, target, groups, num_groups, threshold, ignore_index, validate_args))


def _extreme_ratio(rates: Tensor, tag: str) -> Dict[str, Tensor]:
    """``{tag_<argmin>_<argmax>: min / max}`` — the key needs the two indices on the host (one sync, like the reference)."""
    lo, hi = int(torch.argmin(rates)), int(torch.argmax(rates))
    return {f"{tag}_{lo}_{hi}": _safe_divide(rates[lo], rates[hi])}


def _compute_binary_demographic_parity(tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor) -> Dict[str, Tensor]:
    """Lowest over highest positive-prediction rate (reference :164-174)."""
    return _extreme_ratio(_safe_divide(tp + fp, tp + fp + tn + fn), "DP")


def _compute_binary_equal_opportunity(tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor) -> Dict[str, Tensor]:
    """Lowest over highest true-positive rate (reference :243-255)."""
    return _extreme_ratio(_safe_divide(tp, tp + fn), "EO")


def _present_group_count(groups: Tensor) -> int:
    return int(torch.unique(groups).shape[0])


def demographic_parity(preds: Tensor, groups: Tensor, threshold: float = 0.5, ignore_index: Optional[int] = None,
                       validate_args: bool = True) -> Dict[str, Tensor]:
    """Reference :177-240.  No target is needed: positives are counted against an all-zero one."""
    target = torch.zeros(preds.shape, dtype=torch.long, device=preds.device)
    stats = _binary_groups_stat_scores(preds, target, groups, _present_group_count(groups), threshold, ignore_index, validate_args)
    return _compute_binary_demographic_parity(**_groups_stat_transform(stats))


def equal_opportunity(preds: Tensor, target: Tensor, groups: Tensor, threshold: float = 0.5,
                      ignore_index: Optional[int] = None, validate_args: bool = True) -> Dict[str, Tensor]:
    """Reference :258-323."""
    stats = _binary_groups_stat_scores(preds, target, groups, _present_group_count(groups), threshold, ignore_index, validate_args)
    return _compute_binary_equal_opportunity(**_groups_stat_transform(stats))


def binary_fairness(preds: Tensor, target: Tensor, groups: Tensor,
                    task: Literal["demographic_parity", "equal_opportunity", "all"] = "all", threshold: float = 0.5,
                    ignore_index: Optional[int] = None, validate_args: bool = True) -> Dict[str, Tensor]:
    """Demographic parity, equal opportunity or both (reference :326-382)."""
    _check_task(task)
    if task == "demographic_parity":
        if target is not None:
            rank_zero_warn("The task demographic_parity does not require a target.", UserWarning)
        target = torch.zeros(preds.shape, dtype=torch.long, device=preds.device)
    stats = _groups_stat_transform(
        _binary_groups_stat_scores(preds, target, groups, _present_group_count(groups), threshold, ignore_index, validate_args)
    )
    out: Dict[str, Tensor] = {}
    if task in ("demographic_parity", "all"):
        out.update(_compute_binary_demographic_parity(**stats))
    if task in ("equal_opportunity", "all"):
        out.update(_compute_binary_equal_opportunity(**stats))
    return out
