"""Stat-scores (tp / fp / tn / fn) functionals (reference: functional/classification/stat_scores.py).

Multiclass, ``top_k == 1``, ``multidim_average == "global"`` runs in ONE kernel
(`mb200_multiclass_stat_scores_update`, csrc/confmat.cu): row argmax + per-class counters, never materialising
the ``C*C`` bincount the reference derives tp/fp/fn/tn from (stat_scores.py:435-448).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from typing_extensions import Literal

from metrics_b200 import _native
from metrics_b200.functional.classification._validation import check_multiclass_shapes, labels_as_int, new_flag, raise_if_flagged

_AVERAGES = ("micro", "macro", "weighted", "none", None)
_MULTIDIM = ("global", "samplewise")


def _multiclass_stat_scores_arg_validation(
    num_classes: Optional[int],
    top_k: int = 1,
    average: Optional[str] = "macro",
    multidim_average: str = "global",
    ignore_index: Optional[int] = None,
    zero_division: float = 0,
) -> None:
    """Non-tensor argument rules (reference :218-262)."""
    if num_classes is None and average != "micro":
        raise ValueError(
            f"Argument `num_classes` can only be `None` for `average='micro'`, but got `average={average}`."
        )
    if num_classes is not None and (not isinstance(num_classes, int) or num_classes < 2):
        raise ValueError(f"Expected argument `num_classes` to be an integer larger than 1, but got {num_classes}")
    if not isinstance(top_k, int) or top_k < 1:
        raise ValueError(f"Expected argument `top_k` to be an integer larger than or equal to 1, but got {top_k}")
    if top_k > (num_classes if num_classes is not None else 1):
        raise ValueError(
            f"Expected argument `top_k` to be smaller or equal to `num_classes` but got {top_k} and {num_classes}"
        )
    if average not in _AVERAGES:
        raise ValueError(f"Expected argument `average` to be one of {_AVERAGES}, but got {average}")
    if multidim_average not in _MULTIDIM:
        raise ValueError(
            f"Expected argument `multidim_average` to be one of {_MULTIDIM}, but got {multidim_average}"
        )
    if ignore_index is not None and not isinstance(ignore_index, int):
        raise ValueError(f"Expected argument `ignore_index` to either be `None` or an integer, but got {ignore_index}")
    if zero_division not in [0, 1]:
        raise ValueError(f"Expected argument `zero_division` to be 0 or 1, but got {zero_division}.")


def _multiclass_stat_scores_tensor_validation(
    preds: Tensor,
    target: Tensor,
    num_classes: Optional[int],
    multidim_average: str = "global",
    ignore_index: Optional[int] = None,
) -> None:
    """Host-side shape rules (reference :264-314); label values are range-checked inside the kernel."""
    check_multiclass_shapes(preds, target, num_classes)
    if multidim_average != "global":
        if preds.ndim == target.ndim + 1 and preds.ndim < 3:
            raise ValueError(
                "If `preds` have one dimension more than `target`, the shape of `preds` should be"
                " at least 3D when multidim_average is set to `samplewise`"
            )
        if preds.ndim == target.ndim and preds.ndim < 2:
            raise ValueError(
                "When `preds` and `target` have the same shape, the shape of `preds` should be"
                " at least 2D when multidim_average is set to `samplewise`"
            )


def _require_kernel_mode(top_k: int, multidim_average: str) -> None:
    if top_k != 1 and multidim_average != "global":
        raise NotImplementedError(
            "metrics_b200: multiclass stat scores with `top_k > 1` are implemented for `multidim_average='global'` "
            f"only (got top_k={top_k}, multidim_average={multidim_average!r})."
        )


def _multiclass_stat_scores_update_(
    tp: Tensor,
    fp: Tensor,
    tn: Tensor,
    fn: Tensor,
    workspace: Tensor,
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    top_k: int = 1,
    average: Optional[str] = "macro",
    multidim_average: str = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = False,
) -> None:
    """FUSED format+update: add this batch's tp/fp/tn/fn to the four int64 state tensors in place
    (``multidim_average="global"``; ``top_k > 1`` uses the top-k refined prediction kernel, per-class states)."""
    _require_kernel_mode(top_k, multidim_average)
    if multidim_average != "global":
        raise ValueError("use `_multiclass_stat_scores_update` for samplewise statistics")
    flag = new_flag(tp.device) if validate_args else None
    # validate_args: count into scratch states first and fold them in only after the kernel's error word came back clean, so
    # that a caught validation error leaves the states untouched (the reference validates before it counts, :287-326)
    states = (tp, fp, tn, fn) if flag is None else tuple(torch.zeros_like(s) for s in (tp, fp, tn, fn))
    if top_k > 1:
        _native.multiclass_stat_scores_topk_update_(*states, workspace, preds, target, num_classes, top_k, ignore_index, flag)
    else:
        _native.multiclass_stat_scores_update_(
            *states, workspace, labels_as_int(preds, target), target, num_classes, ignore_index, average == "micro", flag
        )
    if flag is not None:
        raise_if_flagged(flag, num_classes, ignore_index)
        for state, delta in zip((tp, fp, tn, fn), states):
            state += delta


def _class_count_bound(preds: Tensor, target: Tensor) -> int:
    """``num_classes=None`` (legal with ``average="micro"`` only, reference :238-241): the micro counters need no class count
    — except ``tn``, where the reference substitutes 1 (:343, :434; accuracy.py:269), i.e. ``tn = -fp`` — but the kernel's
    range check needs a bound.  It is read from the batch: one device sync, on this corner only."""
    if preds.is_floating_point() and preds.ndim == target.ndim + 1:
        return max(int(preds.shape[1]), 2)
    if preds.numel() == 0:
        return 2
    return max(int(torch.maximum(preds.max(), target.max()).item()) + 1, 2)


def _multiclass_micro_update_unknown_classes_(tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor, preds: Tensor, target: Tensor,
                                             ignore_index: Optional[int], workspace: Optional[Tensor] = None) -> None:
    """Global top-1 micro counters for inputs whose class count was not given; `tn` is rewritten to the reference's `-fp`."""
    bound = _class_count_bound(preds, target)
    if workspace is None or workspace.numel() != 3 * bound + 2:
        workspace = stat_scores_workspace(bound, tp.device)
    _multiclass_stat_scores_update_(tp, fp, tn, fn, workspace, preds, target, bound, 1, "micro", "global", ignore_index, False)
    tn.copy_(-fp)


def stat_scores_workspace(num_classes: int, device: torch.device) -> Tensor:
    """Zeroed, self-cleaning scratch required by the stat-scores kernel (see include/metrics_b200.h)."""
    return torch.zeros(3 * num_classes + 2, dtype=torch.int64, device=device)


def _multiclass_stat_scores_update(
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    top_k: int = 1,
    average: Optional[str] = "macro",
    multidim_average: str = "global",
    ignore_index: Optional[int] = None,
) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    """Functional seam (reference :371-449): fresh tp, fp, tn, fn for one batch (``[N, C]`` when samplewise)."""
    _require_kernel_mode(top_k, multidim_average)
    if multidim_average == "samplewise":
        return _native.multiclass_stat_scores_samplewise(labels_as_int(preds, target), target, num_classes, ignore_index)
    if top_k > 1:
        average = "macro"  # per-class counters: the reference keeps [C]-sized states for top-k, also for micro
    size = () if average == "micro" else (num_classes,)
    states = [torch.zeros(size if size else (1,), dtype=torch.int64, device=preds.device) for _ in range(4)]
    ws = stat_scores_workspace(num_classes, preds.device)
    _multiclass_stat_scores_update_(*states, ws, preds, target, num_classes, top_k, average, multidim_average, ignore_index)
    if average == "micro":
        states = [s.reshape(()) for s in states]
    return states[0], states[1], states[2], states[3]


def _multiclass_stat_scores_compute(
    tp: Tensor,
    fp: Tensor,
    tn: Tensor,
    fn: Tensor,
    average: Optional[str] = "macro",
    multidim_average: str = "global",
) -> Tensor:
    """Stack ``[tp, fp, tn, fn, support]`` and apply the averaging strategy (reference :452-478)."""
    res = torch.stack([tp, fp, tn, fn, tp + fn], dim=-1)
    sum_dim = 0 if multidim_average == "global" else 1
    if average == "micro":
        return res.sum(sum_dim) if res.ndim > 1 else res
    if average == "macro":
        return res.float().mean(sum_dim)
    if average == "weighted":
        weight = tp + fn
        if multidim_average == "global":
            return (res * (weight / weight.sum()).reshape(*weight.shape, 1)).sum(sum_dim)
        return (res * (weight / weight.sum(-1, keepdim=True)).reshape(*weight.shape, 1)).sum(sum_dim)
    if average is None or average == "none":
        return res
    return None


def _multiclass_stat_scores_states(
    preds: Tensor, target: Tensor, num_classes: int, top_k: int, average: Optional[str], multidim_average: str,
    ignore_index: Optional[int], validate_args: bool,
) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    """tp, fp, tn, fn of one call in the layout the reference's functional path produces: 0-d (micro, top-1, global),
    ``[C]`` (global) or ``[N, C]`` (samplewise)."""
    _require_kernel_mode(top_k, multidim_average)
    if multidim_average == "samplewise":
        flag = new_flag(preds.device) if validate_args else None
        out = _native.multiclass_stat_scores_samplewise(labels_as_int(preds, target), target, num_classes, ignore_index, flag)
        if flag is not None:
            raise_if_flagged(flag, num_classes, ignore_index)
        return out
    micro = average == "micro" and top_k == 1
    states = [torch.zeros(1 if micro else num_classes, dtype=torch.int64, device=preds.device) for _ in range(4)]
    ws = stat_scores_workspace(num_classes, preds.device)
    _multiclass_stat_scores_update_(
        *states, ws, preds, target, num_classes, top_k, average, multidim_average, ignore_index, validate_args
    )
    if micro:
        states = [s.reshape(()) for s in states]
    return states[0], states[1], states[2], states[3]


def multiclass_stat_scores(
    preds: Tensor,
    target: Tensor,
    num_classes: int,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    top_k: int = 1,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """tp / fp / tn / fn / support for multiclass inputs (reference :481-600)."""
    if validate_args:
        _multiclass_stat_scores_arg_validation(num_classes, top_k, average, multidim_average, ignore_index)
        _multiclass_stat_scores_tensor_validation(preds, target, num_classes, multidim_average, ignore_index)
    states = _multiclass_stat_scores_states(preds, target, num_classes, top_k, average, multidim_average, ignore_index, validate_args)
    return _multiclass_stat_scores_compute(*states, average, multidim_average)


# =========================================================================================================
# binary / multilabel (kernel K2, csrc/binary.cu)
# =========================================================================================================
from metrics_b200.functional.classification import _binary_counts as _bc  # noqa: E402


def _binary_stat_scores_arg_validation(
    threshold: float = 0.5, multidim_average: str = "global", ignore_index: Optional[int] = None, zero_division: float = 0
) -> None:
    _bc.check_threshold(threshold)
    _bc.check_common(multidim_average, ignore_index, zero_division)


def _binary_stat_scores_tensor_validation(
    preds: Tensor, target: Tensor, multidim_average: str = "global", ignore_index: Optional[int] = None
) -> None:
    """Shape rules (reference :49-88); values are range-checked inside the counting kernel."""
    _bc.binary_shape_validation(preds, target, multidim_average)


def _binary_stat_scores_update(
    preds: Tensor,
    target: Tensor,
    threshold: float = 0.5,
    multidim_average: str = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = False,
) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    """FUSED format+update (reference :95-134): sigmoid-if-logits, threshold, ignore mask and the four masked sums in
    one counting kernel.  Returns 0-d tensors (global) or ``[N]`` tensors (samplewise)."""
    samplewise = multidim_average == "samplewise"
    c = _bc.counts(preds, target, 1, threshold, ignore_index, samplewise, validate_args)
    tp, fp, tn, fn = c.unbind(-1)
    if not samplewise:
        return tp.reshape(()), fp.reshape(()), tn.reshape(()), fn.reshape(())
    return tp.squeeze(), fp.squeeze(), tn.squeeze(), fn.squeeze()


def _binary_stat_scores_compute(tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor, multidim_average: str = "global") -> Tensor:
    return torch.stack([tp, fp, tn, fn, tp + fn], dim=0 if multidim_average == "global" else 1).squeeze()


def binary_stat_scores(
    preds: Tensor,
    target: Tensor,
    threshold: float = 0.5,
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """tp, fp, tn, fn, support for binary inputs (reference :137-213)."""
    if validate_args:
        _binary_stat_scores_arg_validation(threshold, multidim_average, ignore_index)
        _binary_stat_scores_tensor_validation(preds, target, multidim_average, ignore_index)
    tp, fp, tn, fn = _binary_stat_scores_update(preds, target, threshold, multidim_average, ignore_index, validate_args)
    return _binary_stat_scores_compute(tp, fp, tn, fn, multidim_average)


def _multilabel_stat_scores_arg_validation(
    num_labels: int,
    threshold: float = 0.5,
    average: Optional[str] = "macro",
    multidim_average: str = "global",
    ignore_index: Optional[int] = None,
    zero_division: float = 0,
) -> None:
    if not isinstance(num_labels, int) or num_labels < 2:
        raise ValueError(f"Expected argument `num_labels` to be an integer larger than 1, but got {num_labels}")
    if not (isinstance(threshold, float) and (0 <= threshold <= 1)):
        raise ValueError(f"Expected argument `threshold` to be a float, but got {threshold}.")
    if average not in _AVERAGES:
        raise ValueError(f"Expected argument `average` to be one of {_AVERAGES}, but got {average}")
    _bc.check_common(multidim_average, ignore_index, zero_division)


def _multilabel_stat_scores_tensor_validation(
    preds: Tensor, target: Tensor, num_labels: int, multidim_average: str, ignore_index: Optional[int] = None
) -> None:
    _bc.multilabel_shape_validation(preds, target, num_labels, multidim_average)


def _multilabel_stat_scores_update(
    preds: Tensor,
    target: Tensor,
    num_labels: int,
    threshold: float = 0.5,
    multidim_average: str = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = False,
) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    """FUSED format+update (reference :681-714): ``[L]`` (global) or ``[N, L]`` (samplewise) counters."""
    samplewise = multidim_average == "samplewise"
    c = _bc.counts(preds, target, num_labels, threshold, ignore_index, samplewise, validate_args)
    if samplewise:
        c = c.reshape(preds.shape[0], num_labels, 4)
    tp, fp, tn, fn = c.unbind(-1)
    # the reference squeezes the counters (:709-712): a single sample in samplewise mode collapses to [L]
    return tp.squeeze(), fp.squeeze(), tn.squeeze(), fn.squeeze()


def _multilabel_stat_scores_compute(
    tp: Tensor, fp: Tensor, tn: Tensor, fn: Tensor, average: Optional[str] = "macro", multidim_average: str = "global"
) -> Tensor:
    res = torch.stack([tp, fp, tn, fn, tp + fn], dim=-1)
    sum_dim = 0 if multidim_average == "global" else 1
    if average == "micro":
        return res.sum(sum_dim)
    if average == "macro":
        return res.float().mean(sum_dim)
    if average == "weighted":
        w = tp + fn
        return (res * (w / w.sum()).reshape(*w.shape, 1)).sum(sum_dim)
    if average is None or average == "none":
        return res
    return None


def multilabel_stat_scores(
    preds: Tensor,
    target: Tensor,
    num_labels: int,
    threshold: float = 0.5,
    average: Optional[Literal["micro", "macro", "weighted", "none"]] = "macro",
    multidim_average: Literal["global", "samplewise"] = "global",
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """tp, fp, tn, fn, support per label (reference :748-860)."""
    if validate_args:
        _multilabel_stat_scores_arg_validation(num_labels, threshold, average, multidim_average, ignore_index)
        _multilabel_stat_scores_tensor_validation(preds, target, num_labels, multidim_average, ignore_index)
    tp, fp, tn, fn = _multilabel_stat_scores_update(
        preds, target, num_labels, threshold, multidim_average, ignore_index, validate_args
    )
    return _multilabel_stat_scores_compute(tp, fp, tn, fn, average, multidim_average)


def stat_scores(
    preds: Tensor,
    target: Tensor,
    task: Literal["binary", "multiclass", "multilabel"],
    threshold: float = 0.5,
    num_classes: Optional[int] = None,
    num_labels: Optional[int] = None,
    average: Optional[str] = "micro",
    multidim_average: Optional[str] = "global",
    top_k: Optional[int] = 1,
    ignore_index: Optional[int] = None,
    validate_args: bool = True,
) -> Tensor:
    """Task dispatcher (reference :1100-1162)."""
    from metrics_b200.utilities.enums import ClassificationTask

    task = ClassificationTask.from_str(task)
    if task == ClassificationTask.BINARY:
        return binary_stat_scores(preds, target, threshold, multidim_average, ignore_index, validate_args)
    if task == ClassificationTask.MULTICLASS:
        if not isinstance(num_classes, int):
            raise ValueError(f"`num_classes` is expected to be `int` but `{type(num_classes)} was passed.`")
        if not isinstance(top_k, int):
            raise ValueError(f"`top_k` is expected to be `int` but `{type(top_k)} was passed.`")
        return multiclass_stat_scores(preds, target, num_classes, average, top_k, multidim_average, ignore_index, validate_args)
    if not isinstance(num_labels, int):
        raise ValueError(f"`num_labels` is expected to be `int` but `{type(num_labels)} was passed.`")
    return multilabel_stat_scores(preds, target, num_labels, threshold, average, multidim_average, ignore_index, validate_args)


def _refine_preds_oh(preds: Tensor, preds_oh: Tensor, target: Tensor, top_k: int) -> Tensor:
    """Host-side statement of the top-k refinement the kernel fuses (operator seam of reference :347-368): the one-hot
    prediction becomes the target's one-hot when the target is among the ``top_k`` best scores, else the best score's."""
    scores, labels = preds.squeeze(), target.squeeze()
    best = torch.topk(scores, k=top_k, dim=1).indices
    chosen = torch.where((best == labels.unsqueeze(1)).any(dim=1), labels, best[:, 0])
    return torch.zeros_like(preds_oh, dtype=torch.int32).scatter_(-1, chosen.unsqueeze(1).unsqueeze(1), 1)
