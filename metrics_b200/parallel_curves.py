"""Class-sharded evaluation of one-vs-rest curve metrics across ranks (BASELINE cfg5).

The generic sync all-gathers every rank's `[N_r, C]` score matrix to every rank ((W-1) * S bytes received per rank) and
then every rank sorts all C classes.  The one-vs-rest curves are independent per class, so instead:

  1. every rank packs its scores into class-major sort keys `[C_pad, N_r]` (kernel, csrc/curve.cu `pack_keys_kernel`);
  2. ONE `all_to_all_single` gives rank r the key rows of ITS classes `[r*cpr, (r+1)*cpr)` from every rank
     ((W-1)/W * S bytes per rank), targets are all-gathered (8 B/sample);
  3. every rank sorts + scans only its C/W classes over all N samples (`mb200_curve_evaluate_keys`);
  4. the per-class AUROC / AP / counts (a few bytes per class) are all-gathered.

Used by `MulticlassAUROC` / `MulticlassAveragePrecision.compute()` when the default sync would run over an NCCL group;
`sync()` / `unsync()` called explicitly keep their full-gather semantics.
"""
from __future__ import annotations

import os
from typing import Any, Optional, Tuple

import torch
from torch import Tensor

from metrics_b200 import _native
from metrics_b200.parallel_sync import _gather_equal


def sharded_applicable(metric: Any) -> bool:
    """Default sync over an initialised NCCL group with more than one rank, exact mode, CUDA states."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return False
    if os.environ.get("MB200_SHARDED_CURVES", "1") == "0":  # A/B switch: fall back to gather-everything sync
        return False
    if metric.dist_sync_fn is not None or not metric._to_sync or metric.thresholds is not None:
        return False
    available = getattr(metric, "distributed_available_fn", None)
    if available is not None and not available():  # the user switched syncing off for this metric (metric.py sync())
        return False
    group = metric.process_group or torch.distributed.group.WORLD
    if torch.distributed.get_world_size(group) < 2:
        return False
    try:
        if torch.distributed.get_backend(group) != "nccl":
            return False
    except Exception:
        return False
    return metric.device.type == "cuda"


def ovr_curve_scalars_sharded(preds: Optional[Tensor], target: Optional[Tensor], num_classes: int, group: Any,
                              device: torch.device, cached: Optional[tuple] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """(auroc [C], ap [C], counts [C, 3]) over the union of all ranks' samples; identical on every rank.

    ``cached`` is this rank's memoised result for the same state, if any.  Whether it may be used is a collective
    decision (it rides on the sample-count all-reduce): a rank whose MetricCollection never formed compute groups — e.g.
    one that saw no batch — has no shared cache, and skipping the exchange on the other ranks would deadlock it.
    """
    dist = torch.distributed
    group = group or dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if preds is None:  # this rank saw no data
        preds = torch.zeros((0, num_classes), dtype=torch.float32, device=device)
        target = torch.zeros((0,), dtype=torch.int64, device=device)
    if preds.dtype == torch.float64:
        raise NotImplementedError("metrics_b200: float64 scores are not supported by the exact curve kernels")
    target = target.to(torch.int64)
    n_local = preds.shape[0]
    cpr = (num_classes + world - 1) // world  # classes per rank
    c_pad = cpr * world

    # sample counts of every rank (one tiny collective + host read: output sizes depend on it)
    n_all_t = torch.zeros(world + 1, dtype=torch.int64, device=device)
    n_all_t[rank] = n_local
    n_all_t[world] = 1 if cached is not None else 0
    dist.all_reduce(n_all_t, group=group)
    n_all = [int(x) for x in n_all_t.tolist()]
    if n_all.pop() == world:
        return cached
    n_total = sum(n_all)
    if n_total == 0:
        raise IndexError("metrics_b200: cannot evaluate a curve metric without samples")

    packed_rows = world * cpr
    ws = _peer_workspace(group, device, cpr, n_total, packed_rows)
    if ws is not None:
        return _exchange_over_peer_memory(ws, preds, target, num_classes, cpr, n_all, rank, world)

    # ---- NCCL path (no peer memory: other backends' groups never get here, see sharded_applicable) -------------------------
    # targets of all ranks, in rank order (ragged ranks: pad to the longest, gather, compact)
    n_max = max(n_all)
    if n_max == n_local and len(set(n_all)) == 1:
        tgt_all = torch.empty(n_total, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(tgt_all, target.contiguous(), group=group)
    else:
        padded = torch.zeros(n_max, dtype=torch.int64, device=device)
        padded[:n_local] = target
        slab = _gather_equal(padded, group, world)  # [world, n_max]
        tgt_all = torch.cat([slab[r, :n] for r, n in enumerate(n_all)])

    # class-major keys [c_pad, n_local]; rows [d*cpr, (d+1)*cpr) go to rank d
    keys_local = _native.curve_pack_keys(preds, rows_out=c_pad)
    recv = torch.empty(cpr * n_total, dtype=torch.int32, device=device)
    dist.all_to_all_single(recv, keys_local.reshape(-1), output_split_sizes=[cpr * n for n in n_all],
                           input_split_sizes=[cpr * n_local] * world, group=group)
    # [source rank][cpr][n_s] -> [cpr][n_total] with the samples in rank order (matches tgt_all)
    keys_mine = torch.empty((cpr, n_total), dtype=torch.int32, device=device)
    off_e, off_n = 0, 0
    for n in n_all:
        if n:
            keys_mine[:, off_n:off_n + n] = recv[off_e:off_e + cpr * n].view(cpr, n)
        off_e += cpr * n
        off_n += n

    auroc, ap, counts = _native.curve_evaluate_keys(keys_mine, tgt_all, first_class=rank * cpr, nonneg=True)  # metric states: post-format scores
    # per-class results of every rank
    packed = torch.cat([auroc.double().unsqueeze(1), ap.double().unsqueeze(1), counts.double()], dim=1).contiguous()  # [cpr, 5]
    gathered = _gather_equal(packed, group, world).reshape(world * cpr, 5)[:num_classes]
    return gathered[:, 0].float(), gathered[:, 1].float(), gathered[:, 2:].round().to(torch.int64)


def _layout(cpr: int, n_total: int, packed_rows: int) -> Tuple[int, int, int, int]:
    """Byte offsets of (keys u32 [cpr][n_total], targets i64 [n_total], results f64 [packed_rows][5]) and the total."""
    a = 256
    keys_off = 0
    tgt_off = (keys_off + cpr * n_total * 4 + a - 1) // a * a
    res_off = (tgt_off + n_total * 8 + a - 1) // a * a
    total = (res_off + packed_rows * 5 * 8 + a - 1) // a * a
    return keys_off, tgt_off, res_off, total


def _peer_workspace(group: Any, device: torch.device, cpr: int, n_total: int, packed_rows: int):
    from metrics_b200 import peer

    return peer.get(group, device, _layout(cpr, n_total, packed_rows)[3])


def _exchange_over_peer_memory(ws: Any, preds: Tensor, target: Tensor, num_classes: int, cpr: int, n_all: list, rank: int,
                               world: int) -> Tuple[Tensor, Tensor, Tensor]:
    """The exchange as stores into the owners' memory (csrc/peer.cu): ONE fused pack + put kernel for the scores, one put
    for the targets, a barrier; the owner sorts its key matrix where it landed; one put + barrier for the per-class scalars.
    No staging buffers, no all-to-all, no reorder pass, no host synchronisation beyond the sample-count read above.

    The leading barrier makes region reuse safe whatever used the workspace before (a rank may still be reading the result
    region of the previous exchange, or the output of a peer all-reduce, when a faster rank starts storing the next keys)."""
    n_total = sum(n_all)
    col_off = sum(n_all[:rank])
    keys_off, tgt_off, res_off, _ = _layout(cpr, n_total, world * cpr)
    ws.barrier()
    ws.pack_keys_put(preds, cpr, n_total, col_off, keys_off)
    if preds.shape[0]:
        ws.put_all(target, tgt_off + col_off * 8)
    ws.barrier()
    keys_mine = ws.view(keys_off, (cpr, n_total), torch.int32)
    tgt_all = ws.view(tgt_off, (n_total,), torch.int64)
    auroc, ap, counts = _native.curve_evaluate_keys(keys_mine, tgt_all, first_class=rank * cpr, nonneg=True)  # metric states: post-format scores
    packed = torch.cat([auroc.double().unsqueeze(1), ap.double().unsqueeze(1), counts.double()], dim=1).contiguous()  # [cpr, 5]
    ws.put_all(packed, res_off + rank * cpr * 5 * 8)
    ws.barrier()
    gathered = ws.view(res_off, (world * cpr, 5), torch.float64)[:num_classes]
    return gathered[:, 0].float(), gathered[:, 1].float(), gathered[:, 2:].round().to(torch.int64)
