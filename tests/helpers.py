"""Shared test helpers: seeded input recipes (identical to tests/golden/make_golden.py) and digests."""
import hashlib

import numpy as np
import torch


def sha(t: torch.Tensor) -> str:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        t = t.view(torch.int16)
    return hashlib.sha256(t.numpy().tobytes()).hexdigest()


def to_np(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu()
    if t.dtype in (torch.bfloat16, torch.float16):
        t = t.float()
    return t.numpy()


def cfg1_inputs():
    g = torch.Generator().manual_seed(0)
    preds = torch.randn(100, 1024, 5, generator=g)
    target = torch.randint(0, 5, (100, 1024), generator=g)
    return preds, target


def cfg2_inputs():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(65536, 1000, generator=g).bfloat16()
    target = torch.randint(0, 1000, (65536,), generator=g)
    return logits, target


def stats_inputs(C: int, N: int):
    """Replays the generator stream of make_golden.py section E up to the requested case."""
    g = torch.Generator().manual_seed(11)
    for c, n in ((5, 300), (1000, 4096)):
        logits = torch.randn(n, c, generator=g)
        target = torch.randint(0, c, (n,), generator=g)
        if c == 5:
            target[target == 3] = 1
        if (c, n) == (C, N):
            return logits, target
    raise KeyError((C, N))


TORCH_DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16, "f64": torch.float64}


def cfg3_inputs():
    g = torch.Generator().manual_seed(0)
    preds = torch.rand(1000, 10000, generator=g)
    target = torch.randint(0, 2, (1000, 10000), generator=g)
    return preds, target


MC_CASES = ((5, 400, "probs"), (5, 400, "logits"), (37, 1500, "logits"), (1000, 2048, "logits"))


def mc_inputs(C: int, N: int, kind: str):
    """Replays the generator stream of make_golden.py (curves, multiclass section)."""
    g = torch.Generator().manual_seed(33)
    for c, n, k in MC_CASES:
        logits = torch.randn(n, c, generator=g)
        tgt = torch.randint(0, c, (n,), generator=g)
        if c == 5:
            tgt[tgt == 4] = 2
        p = torch.softmax(logits, 1) if k == "probs" else logits
        if (c, n, k) == (C, N, kind):
            return p, tgt
    raise KeyError((C, N, kind))


def cfg5_rank_batches(rank: int, n_batches: int):
    torch.manual_seed(rank)
    out = []
    for _ in range(n_batches):
        lg = torch.randn(4096, 1000)
        tg = torch.randint(0, 1000, (4096,))
        out.append((lg, tg))
    return out
