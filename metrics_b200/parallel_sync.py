"""Bucketed cross-rank state exchange used by ``Metric.sync`` when no custom ``dist_sync_fn`` is supplied.

Replaces the per-state ``barrier + all_gather(shape) + all_gather(data)`` of the reference
(utilities/distributed.py:100-153, called once per state from metric.py:518-523) by:

* ONE ``all_reduce`` per (dtype, op) bucket holding every integer "sum"/"max"/"min" tensor state (bit-exact:
  integer addition / max / min are associative and commutative);
* floating "sum"/"mean"/"max"/"min" tensor states are gathered and reduced in rank order, exactly like the
  reference (``stack`` then ``sum(dim=0)``), so floating results are bit-identical to gather-then-reduce;
* ONE shape exchange for all "cat" states together, then one all-gather per cat state into pre-sized buffers
  (no pad on every dim, no barrier);
* states with ``dist_reduce_fx=None`` or a custom callable use ``gather_all_tensors`` per tensor.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import torch
from torch import Tensor

from metrics_b200.utilities.data import _flatten, dim_zero_cat, dim_zero_max, dim_zero_mean, dim_zero_min, dim_zero_sum
from metrics_b200.utilities.distributed import gather_all_tensors

_MAX_DIMS = 8
_DTYPE_CODES = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.int16,
                torch.int8, torch.uint8, torch.bool]
_INT_DTYPES = (torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8)


def _reduce_op(fn: Any) -> Optional[Any]:
    ops = torch.distributed.ReduceOp
    if fn is dim_zero_sum:
        return ops.SUM
    if fn is dim_zero_max:
        return ops.MAX
    if fn is dim_zero_min:
        return ops.MIN
    return None


def _gather_equal(t: Tensor, group: Any, world: int) -> Tensor:
    """Stack of the same-shaped tensor from every rank: [world, *t.shape]."""
    flat = t.contiguous().reshape(-1)
    if flat.numel() == 0:
        return torch.empty((world, *t.shape), dtype=t.dtype, device=t.device)
    out = torch.empty(world * flat.numel(), dtype=t.dtype, device=t.device)
    torch.distributed.all_gather_into_tensor(out, flat, group=group)
    return out.reshape(world, *t.shape)


_PEER_MIN_BYTES = 256 * 1024  # below this NCCL's small-message latency (~30 us) is as good as two signal barriers


def _peer_all_reduce(parts: List[Tensor], dtype: torch.dtype, op: Any, group: Any) -> Optional[Tensor]:
    """int64 bucket all-reduce over NVLink peer memory (csrc/peer.cu `mb200_peer_reduce_put_i64`): every rank stores its
    bucket in the group's symmetric workspace, reduces ITS slice of all ranks' buckets with peer loads and stores the
    reduced slice into every rank with peer stores — bit-exact like the NCCL all-reduce it replaces.  None = not applicable
    (small bucket, other dtype, CPU / non-NCCL group, peer memory unavailable): the caller uses NCCL."""
    if dtype != torch.int64 or not parts[0].is_cuda or os.environ.get("MB200_PEER_ALLREDUCE", "0") != "1":
        # opt-in: measured on 2 x B200 (profiles/r02_sync_2gpu.json) the 8 MB bucket takes 80 us this way (three launches,
        # two signal barriers, 63 us of host time) against 58 us for NCCL's all-reduce — a plain reduction has no compute to
        # fuse with, so NCCL stays the default; the fused pack + put of the curve exchange is where peer stores pay.
        return None
    n = sum(p.numel() for p in parts)
    if n * 8 < _PEER_MIN_BYTES:
        return None
    from metrics_b200 import peer

    ops = torch.distributed.ReduceOp
    code = {ops.SUM: 0, ops.MAX: 1, ops.MIN: 2}.get(op)
    if code is None:
        return None
    region = (n * 8 + 255) // 256 * 256
    ws = peer.get(group, parts[0].device, 2 * region)
    if ws is None:
        return None
    src = ws.view(0, (n,), torch.int64)
    # No leading barrier: the copy-in only writes THIS rank's block, and remote stores into other ranks' blocks are issued
    # after the first barrier below, which a rank enters behind everything it enqueued earlier on this stream.
    offset = 0
    for p in parts:
        src[offset: offset + p.numel()].copy_(p.reshape(-1))
        offset += p.numel()
    ws.barrier()  # every rank's input is in place
    ws.reduce_put_i64(0, region, n, code)
    ws.barrier()  # every slice of the result has landed here
    return ws.view(region, (n,), torch.int64).clone()


def sync_states_bucketed(metric: Any, group: Optional[Any]) -> bool:
    """Synchronise all states of ``metric`` in place.  Returns False if the fast path does not apply."""
    if group is None:
        group = torch.distributed.group.WORLD
    world = torch.distributed.get_world_size(group)
    rank = torch.distributed.get_rank(group)

    int_buckets: Dict[Any, List[str]] = {}
    float_reduce: List[str] = []
    cat_states: List[str] = []
    generic: List[str] = []
    for name, fn in metric._reductions.items():
        value = getattr(metric, name)
        if isinstance(value, Tensor):
            op = _reduce_op(fn)
            if op is not None and value.dtype in _INT_DTYPES:
                int_buckets.setdefault((value.dtype, op), []).append(name)
            elif op is not None or fn is dim_zero_mean:
                float_reduce.append(name)
            else:
                generic.append(name)
        elif isinstance(value, list) and fn is dim_zero_cat:
            cat_states.append(name)
        else:
            generic.append(name)

    # ---- integer reductions: one collective per (dtype, op) -------------------------------------------------
    for (dtype, op), names in int_buckets.items():
        parts = [getattr(metric, n).reshape(-1) for n in names]
        bucket = _peer_all_reduce(parts, dtype, op, group)
        if bucket is None:
            bucket = torch.cat(parts) if len(parts) > 1 else parts[0].clone()
            torch.distributed.all_reduce(bucket, op=op, group=group)
        offset = 0
        for n in names:
            ref = getattr(metric, n)
            setattr(metric, n, bucket[offset: offset + ref.numel()].reshape(ref.shape))
            offset += ref.numel()

    # ---- floating reductions: gather then reduce in rank order (bit-identical to the reference) --------------
    for n in float_reduce:
        stacked = _gather_equal(getattr(metric, n), group, world)
        setattr(metric, n, metric._reductions[n](stacked))

    # ---- cat states: one shape exchange for all of them, then one gather each ---------------------------------
    if cat_states:
        locals_: List[Tensor] = []
        for n in cat_states:
            value = getattr(metric, n)
            if len(value) == 0:
                locals_.append(torch.tensor([], device=metric.device, dtype=metric.dtype))
            else:
                locals_.append(dim_zero_cat(value).contiguous())
        dev = locals_[0].device
        desc = torch.zeros((len(cat_states), 2 + _MAX_DIMS), dtype=torch.int64)
        for i, t in enumerate(locals_):
            if t.ndim > _MAX_DIMS or t.dtype not in _DTYPE_CODES:
                # not expressible in the descriptor: say so IN the descriptor, so that every rank — also one that holds an
                # empty placeholder of another dtype — routes this state through the generic gather together (returning False
                # here would be a rank-local decision taken after the integer buckets above were already reduced)
                desc[i, 0] = -1
                continue
            desc[i, 0] = t.ndim
            desc[i, 1] = _DTYPE_CODES.index(t.dtype)
            for d, s in enumerate(t.shape):
                desc[i, 2 + d] = s
        desc = desc.to(dev)
        all_desc = _gather_equal(desc, group, world).cpu()  # [world, n_states, 2 + MAX_DIMS]; the one host sync
        for i, n in enumerate(cat_states):
            t = locals_[i]
            if any(int(all_desc[r, i, 0]) < 0 for r in range(world)):
                pieces = gather_all_tensors(t, group=group)
                setattr(metric, n, dim_zero_cat([p for p in pieces if p.numel() > 0] or [t]))
                continue
            shapes = [tuple(int(x) for x in all_desc[r, i, 2: 2 + int(all_desc[r, i, 0])]) for r in range(world)]
            numels = [int(torch.Size(s).numel()) for s in shapes]
            nonempty = [r for r in range(world) if numels[r] > 0]
            if not nonempty:
                setattr(metric, n, dim_zero_cat([t]) if t.numel() == 0 else t)
                continue
            # a rank that saw no data contributes an empty placeholder: it adopts the dtype the data-holding ranks use
            dtype = _DTYPE_CODES[int(all_desc[nonempty[0], i, 1])]
            if t.numel() == 0 and t.dtype != dtype:
                t = t.to(dtype)
            trailing = shapes[nonempty[0]][1:]
            if all(s == shapes[0] for s in shapes):
                gathered = _gather_equal(t, group, world)  # [world, n, *trailing]
                setattr(metric, n, gathered.reshape(-1, *trailing) if t.ndim >= 1 else gathered)
                continue
            if all(shapes[r][1:] == trailing for r in nonempty):
                # ragged only along dim 0: gather max-size flats, then take each rank's valid prefix
                max_numel = max(numels)
                flat = t.reshape(-1)
                if flat.numel() < max_numel:
                    padded = torch.zeros(max_numel, dtype=dtype, device=dev)
                    padded[: flat.numel()] = flat
                else:
                    padded = flat
                buf = _gather_equal(padded, group, world)  # [world, max_numel]
                pieces = [buf[r, : numels[r]].reshape(shapes[r]) for r in nonempty]
                pieces = [t if r == rank else p for r, p in zip(nonempty, pieces)]
                setattr(metric, n, torch.cat(pieces, dim=0))
                continue
            pieces = gather_all_tensors(t, group=group)
            setattr(metric, n, dim_zero_cat([p for p in pieces if p.numel() > 0]))

    # ---- everything else: the generic per-tensor gather -------------------------------------------------------
    for n in generic:
        value = getattr(metric, n)
        fn = metric._reductions[n]
        if isinstance(value, Tensor):
            out: Any = torch.stack(gather_all_tensors(value, group=group))
        else:
            if len(value) == 0 and fn is not None:
                setattr(metric, n, [])
                continue
            gathered = [gather_all_tensors(v, group=group) for v in value]  # list (per element) of per-rank lists
            out = _flatten(gathered)
            if len(out) == 0:
                setattr(metric, n, [])
                continue
        setattr(metric, n, fn(out) if fn is not None else out)
    return True
