"""COCO run-length codes of instance masks, host side (numpy) — only for the json formats on either side of the metric
(`MeanAveragePrecision.tm_to_coco` / `coco_to_tm` with ``iou_type="segm"``); the evaluation itself never builds a run-length
code (masks are bit-packed on the device, csrc/maskiou.cu).

The format (cocodataset.org/#format-data, pycocotools `maskApi.c`): a mask of height h and width w is read in COLUMN-major
order; ``counts`` are the lengths of the alternating runs of 0s and 1s, starting with 0s (so the first count may be 0).
"Uncompressed" codes carry ``counts`` as a list of ints; "compressed" ones as the ASCII string of `maskApi.c:rleToString`:
every count (from the fourth on: its difference to the count two places back) is written as little-endian groups of 5 bits,
bit 5 of a character says that more groups follow, the sign is carried by bit 4 of the last group, characters are offset by 48.
The string codec is restated from that published algorithm; pycocotools is not installed in this project's containers, so it
is checked by round trips and hand-computed cases only (tests/test_rle.py).  Polygon segmentations are not supported.
"""
from __future__ import annotations

from typing import Any, List, Sequence, Union

import numpy as np


def mask_to_counts(mask: np.ndarray) -> List[int]:
    """bool ``[h, w]`` -> run lengths in column-major order, starting with the run of zeros."""
    flat = np.asarray(mask, dtype=bool).T.reshape(-1)  # column-major
    if flat.size == 0:
        return []
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    edges = np.concatenate([[0], change, [flat.size]])
    runs = np.diff(edges).tolist()
    return ([0] + runs) if flat[0] else runs


def counts_to_mask(counts: Sequence[int], h: int, w: int) -> np.ndarray:
    """Run lengths (column-major, zeros first) -> bool ``[h, w]``."""
    counts = np.asarray(list(counts), dtype=np.int64)
    if int(counts.sum()) != h * w:
        raise ValueError(f"run-length code covers {int(counts.sum())} pixels, the mask has {h} x {w}")
    values = np.zeros(counts.size, dtype=bool)
    values[1::2] = True
    return np.repeat(values, counts).reshape(w, h).T


def counts_to_string(counts: Sequence[int]) -> str:
    """`maskApi.c:rleToString`."""
    out = []
    counts = [int(c) for c in counts]
    for i, c in enumerate(counts):
        x = c - counts[i - 2] if i > 2 else c
        while True:
            group = x & 0x1F
            x >>= 5  # arithmetic shift: -1 stays -1
            more = (x != -1) if (group & 0x10) else (x != 0)
            out.append(chr((group | 0x20 if more else group) + 48))
            if not more:
                break
    return "".join(out)


def string_to_counts(code: Union[str, bytes]) -> List[int]:
    """`maskApi.c:rleFrString`."""
    if isinstance(code, bytes):
        code = code.decode("ascii")
    counts: List[int] = []
    pos = 0
    while pos < len(code):
        x, k, more = 0, 0, True
        while more:
            c = ord(code[pos]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            pos += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def segmentation_to_mask(segmentation: Any, h: int = 0, w: int = 0) -> np.ndarray:
    """The `segmentation` field of a COCO annotation -> uint8 ``[h, w]`` (pycocotools `annToMask`), for run-length codes."""
    if isinstance(segmentation, dict) and "counts" in segmentation:
        hh, ww = (int(x) for x in segmentation["size"])
        counts = segmentation["counts"]
        if isinstance(counts, (str, bytes)):
            counts = string_to_counts(counts)
        return counts_to_mask(counts, hh, ww).astype(np.uint8)
    raise NotImplementedError("metrics_b200: only run-length coded segmentations are supported (polygons need pycocotools' rasteriser)")


def entry_to_masks(entry: np.ndarray) -> np.ndarray:
    """A `MeanAveragePrecision` mask state entry (int32 ``[n, H, W, areas.., bit rows..]``, see `_mask_state`) -> bool
    ``[n, H, W]``."""
    entry = np.ascontiguousarray(entry, dtype=np.int32)
    n, h, w = (int(x) for x in entry[:3])
    words = (h * w + 31) // 32
    rows = entry[3 + n: 3 + n + n * words].reshape(n, words)
    bits = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, : h * w]
    return bits.reshape(n, h, w).astype(bool)
