"""Cross-rank gathering of metric states.

``gather_all_tensors`` keeps the reference plug-in contract (utilities/distributed.py:100-153):
``fn(tensor, group=None) -> list[Tensor]`` of length world_size, element i = rank i's tensor, ragged shapes
allowed, own-rank entry is the input object.  The implementation differs: no barrier, ONE size exchange and
ONE padded-free data exchange (all_gather_into_tensor on a flat byte-exact buffer when shapes are equal,
otherwise an all_gather of max-numel flats sliced back — never `F.pad` on every dim).

The collection-level fast path (one bucketed all-reduce for every "sum" state, one size exchange for every
"cat" state) lives in ``metrics_b200.parallel_sync`` and is used by ``Metric.sync`` when no custom
``dist_sync_fn`` is supplied.
"""
from __future__ import annotations

from typing import Any, List, Optional

import torch
from torch import Tensor


def _world(group: Optional[Any]) -> int:
    return torch.distributed.get_world_size(group)


def gather_all_tensors(result: Tensor, group: Optional[Any] = None) -> List[Tensor]:
    """All-gather ``result`` from every rank of ``group``; shapes may differ between ranks."""
    if group is None:
        group = torch.distributed.group.WORLD
    result = result.contiguous()
    world = _world(group)
    rank = torch.distributed.get_rank(group)
    if result.ndim == 0:
        gathered = [torch.zeros_like(result) for _ in range(world)]
        torch.distributed.all_gather(gathered, result, group=group)
        gathered[rank] = result
        return gathered

    # one exchange of shapes (ndim is identical across ranks by contract)
    local_shape = torch.tensor(result.shape, device=result.device, dtype=torch.int64)
    shapes = [torch.zeros_like(local_shape) for _ in range(world)]
    torch.distributed.all_gather(shapes, local_shape, group=group)
    shapes_host = torch.stack(shapes).cpu().tolist()

    if all(s == shapes_host[0] for s in shapes_host):
        gathered = [torch.zeros_like(result) for _ in range(world)]
        torch.distributed.all_gather(gathered, result, group=group)
        gathered[rank] = result
        return gathered

    numels = [int(torch.Size(s).numel()) for s in shapes_host]
    max_numel = max(numels)
    flat = result.reshape(-1)
    if flat.numel() < max_numel:
        padded = flat.new_zeros(max_numel)
        padded[: flat.numel()] = flat
    else:
        padded = flat
    buf = [torch.empty_like(padded) for _ in range(world)]
    torch.distributed.all_gather(buf, padded, group=group)
    out = [b[:n].reshape(s) for b, n, s in zip(buf, numels, shapes_host)]
    out[rank] = result
    return out


def reduce(x: Tensor, reduction: Optional[str]) -> Tensor:
    """``"elementwise_mean"`` | ``"sum"`` | ``"none"`` / ``None`` (reference :22-42)."""
    if reduction == "elementwise_mean":
        return torch.mean(x)
    if reduction == "sum":
        return torch.sum(x)
    if reduction is None or reduction == "none":
        return x
    raise ValueError("Reduction parameter unknown.")


def class_reduce(num: Tensor, denom: Tensor, weights: Tensor, class_reduction: Optional[str] = "none") -> Tensor:
    """``num / denom`` per class with 0 where the quotient is NaN, then micro / macro / weighted / none (reference :45-88)."""
    valid_reduction = ("micro", "macro", "weighted", "none", None)
    if class_reduction not in valid_reduction:
        raise ValueError(f"Reduction parameter {class_reduction} unknown. Choose between one of these: {valid_reduction}")
    fraction = num.sum() / denom.sum() if class_reduction == "micro" else num / denom
    fraction = torch.where(torch.isnan(fraction), torch.zeros_like(fraction), fraction)
    if class_reduction == "macro":
        return fraction.mean()
    if class_reduction == "weighted":
        return (fraction * (weights.float() / weights.sum())).sum()
    return fraction
