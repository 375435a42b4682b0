"""Peer-memory workspace for the cross-rank state exchange (csrc/peer.cu, K10).

One process per GPU.  A workspace is ONE symmetric allocation per (process group, device): the same number of bytes on
every rank, each rank's block mapped into every other rank's address space over NVLink
(`torch.distributed._symmetric_memory`: allocation + handle exchange + signal pads — plumbing; the kernels that move the
data are this package's).  Exchange pattern, all on the caller's stream:

    barrier            every rank is done READING the regions about to be overwritten (previous exchange)
    put / pack+put     kernels store straight into the owners' memory (csrc/peer.cu)
    barrier            every rank's stores have landed: the local block can be consumed by ordinary kernels

Creating or growing the workspace is a collective (every rank must ask for the same size at the same point), so sizes are
always derived from quantities all ranks agree on.  Everything here fails soft: `get()` returns None when symmetric memory
cannot be brought up (no NVLink peer access, a non-NCCL group, `MB200_PEER_EXCHANGE=0`) and the callers keep their NCCL
collectives — on every rank alike, because the decision is taken collectively.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional, Tuple

import torch
from torch import Tensor

from metrics_b200 import _native

_ALIGN = 256
_workspaces: Dict[Tuple[int, int], "PeerWorkspace"] = {}
_disabled: Dict[int, str] = {}


def _round_up(n: int, a: int = _ALIGN) -> int:
    return (n + a - 1) // a * a


class PeerWorkspace:
    """`nbytes` of symmetric memory on every rank of `group` plus the device-resident table of all ranks' base pointers."""

    def __init__(self, group: Any, device: torch.device, nbytes: int) -> None:
        import torch.distributed._symmetric_memory as symm

        self.group = group
        self.device = device
        self.nbytes = nbytes
        self.world = torch.distributed.get_world_size(group)
        self.rank = torch.distributed.get_rank(group)
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.table = int(self.hdl.buffer_ptrs_dev)  # device array [world] of base pointers
        self.base = self.buf.data_ptr()
        if int(self.hdl.buffer_ptrs[self.rank]) != self.base:
            raise RuntimeError("symmetric allocation: the local block is not this rank's entry of the pointer table")

    def barrier(self) -> None:
        """Signal-pad barrier over all ranks, enqueued on the current stream (system-scope release / acquire)."""
        self.hdl.barrier(channel=0)

    def view(self, offset: int, shape: Tuple[int, ...], dtype: torch.dtype) -> Tensor:
        """This rank's block, reinterpreted: a plain CUDA tensor aliasing workspace bytes [offset, offset + size)."""
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if offset % 16 or offset + nbytes > self.nbytes:
            raise ValueError("workspace view out of range or misaligned")
        return self.buf[offset: offset + nbytes].view(dtype).view(shape)

    # ---- kernels ---------------------------------------------------------------------------------------------------
    def put_all(self, src: Tensor, dst_offset: int) -> None:
        """Store `src` (contiguous) at byte `dst_offset` of EVERY rank's block (`mb200_peer_put_all`)."""
        src = src.contiguous()
        nbytes = src.numel() * src.element_size()
        if dst_offset + nbytes > self.nbytes:
            raise ValueError("put_all beyond the workspace")
        with _native.on_device(self.device):
            rc = _native.lib().mb200_peer_put_all(src.data_ptr() if nbytes else None, nbytes, self.table, dst_offset,
                                                  self.world, _native.stream_handle(self.device))
        _native.check(rc, "peer_put_all")

    def pack_keys_put(self, preds: Tensor, classes_per_rank: int, n_total: int, col_offset: int, keys_offset: int) -> None:
        """Fused key packing + class-sharded exchange (`mb200_peer_pack_keys_put`): class-major sort keys of this rank's
        [n, C] scores, each class row stored into its owner's key matrix [classes_per_rank][n_total] at this rank's columns."""
        _native.require_cuda(preds)
        preds = preds.contiguous()
        n, c = preds.shape
        if keys_offset + classes_per_rank * n_total * 4 > self.nbytes:
            raise ValueError("key matrix beyond the workspace")
        with _native.on_device(self.device):
            rc = _native.lib().mb200_peer_pack_keys_put(preds.data_ptr() if n else None, _native.tag(preds), n, c,
                                                        classes_per_rank, self.world, n_total, col_offset, self.table,
                                                        keys_offset, _native.stream_handle(self.device))
        _native.check(rc, "peer_pack_keys_put")

    def reduce_put_i64(self, in_offset: int, out_offset: int, n: int, op: int) -> None:
        """All ranks hold int64 [n] at `in_offset`; afterwards every rank holds the reduction at `out_offset`
        (`mb200_peer_reduce_put_i64`; op 0 sum / 1 max / 2 min).  Needs a barrier before (inputs stored) and after."""
        with _native.on_device(self.device):
            rc = _native.lib().mb200_peer_reduce_put_i64(self.table, in_offset, out_offset, n, self.rank, self.world, op,
                                                         _native.stream_handle(self.device))
        _native.check(rc, "peer_reduce_put_i64")


def _group_key(group: Any) -> Any:
    """c10d's unique name of the group (an `id()` could be handed to a later group after this one is destroyed)."""
    return getattr(group, "group_name", None) or id(group)


def enabled() -> bool:
    return os.environ.get("MB200_PEER_EXCHANGE", "1") != "0"


def get(group: Any, device: torch.device, nbytes: int) -> Optional[PeerWorkspace]:
    """The group's workspace with at least `nbytes` bytes, created / grown collectively; None when peer memory is not
    available.  Every rank must call this with the same `nbytes` at the same point of its collective sequence."""
    if not enabled() or device.type != "cuda":
        return None
    group = group or torch.distributed.group.WORLD
    key = (_group_key(group), device.index if device.index is not None else torch.cuda.current_device())
    if key[0] in _disabled:
        return None
    ws = _workspaces.get(key)
    if ws is not None and ws.group is not group:  # a re-initialised process group reusing the name: never talk to the old one
        _workspaces.pop(key)
        ws = None
    if ws is not None and ws.nbytes >= nbytes:
        return ws
    want = _round_up(max(nbytes + nbytes // 4, 1 << 20), 1 << 20)  # headroom: growing is a collective re-allocation
    ok = 1
    new_ws = None
    try:
        if torch.distributed.get_backend(group) != "nccl":
            raise RuntimeError("peer exchange needs an NCCL group")
        new_ws = PeerWorkspace(group, device, want)
    except Exception as err:  # noqa: BLE001 - any failure means "use the NCCL collectives"
        ok = 0
        _disabled[key[0]] = repr(err)
    # collective agreement: one failing rank switches the path off everywhere (otherwise the ranks would diverge)
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        _disabled.setdefault(key[0], "another rank could not create the symmetric allocation")
        return None
    _workspaces[key] = new_ws
    return new_ws


def why_disabled(group: Any = None) -> Optional[str]:
    group = group or torch.distributed.group.WORLD
    return _disabled.get(_group_key(group))
