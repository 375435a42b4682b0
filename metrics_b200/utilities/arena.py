"""Arena-backed list states.

An exact-mode curve metric keeps every batch's formatted scores and targets in list states (``dist_reduce_fx="cat"``,
reference classification/precision_recall_curve.py:151-160) and concatenates them at ``compute()`` — for cfg3 two
``torch.cat`` calls over 1000 entries, 1.1 ms of a 1.7 ms compute.  An `ArenaList` is a ``list`` whose entries are
CONSECUTIVE VIEWS of one growing device buffer: the format kernel writes each batch where it will finally live, the list
looks and behaves like the reference's (entries are ordinary tensors; ``state_dict``, ``sync``, compute groups, ``forward``
snapshots see a list of tensors), and `packed()` hands ``compute()`` the concatenation without copying.

Anything that bypasses `reserve` / `commit` (an outside ``append``, an element assignment, a device move that rebuilds the
list) simply makes `packed()` answer ``None`` and the caller concatenates as before.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor


class ArenaList(list):
    """``list`` of tensors that are consecutive 1-D views of ``buffer[:used]`` (see module docstring)."""

    def __init__(self) -> None:
        super().__init__()
        self.buffer: Optional[Tensor] = None
        self.used = 0
        self.sizes: List[int] = []

    def clear(self) -> None:  # `Metric.reset`: forget the buffer too — views handed out earlier must stay intact
        super().clear()
        self.buffer, self.used, self.sizes = None, 0, []

    def consistent(self) -> bool:
        n = len(self.sizes)
        if len(self) != n:
            return False
        if n == 0:
            return True
        last = self[-1]
        return (self.buffer is not None and isinstance(last, Tensor) and last.dtype == self.buffer.dtype
                and last.data_ptr() == self.buffer.data_ptr() + (self.used - self.sizes[-1]) * self.buffer.element_size())

    def grow(self, need: int, dtype: torch.dtype, device: torch.device) -> None:
        """Make room for ``need`` elements in total (doubling); existing entries are re-pointed at the new buffer."""
        old = self.buffer
        cap = max(need, 2 * (old.numel() if old is not None else 0), 1 << 16)
        grown = torch.empty(cap, dtype=dtype, device=device)
        if self.used:
            grown[: self.used].copy_(old[: self.used])
            super().clear()
            super().extend(torch.split(grown[: self.used], self.sizes))
        self.buffer = grown

    def packed(self) -> Optional[Tensor]:
        """All entries as ONE tensor without a copy, or ``None`` when the list was changed behind the arena's back."""
        if len(self.sizes) == 0 or not self.consistent():
            return None
        return self.buffer[: self.used]
