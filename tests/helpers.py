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


def synth_detection(seed: int, n_img: int, n_gt: int, n_det: int, n_cls: int, crowd_frac: float = 0.0, dup_scores: bool = False):
    """Scaled-down BASELINE cfg4 recipe: 640x480 images, jittered-gt detections + random ones."""
    g = torch.Generator().manual_seed(seed)
    preds, target = [], []
    for _ in range(n_img):
        x1 = torch.rand(n_gt, generator=g) * 540
        y1 = torch.rand(n_gt, generator=g) * 380
        w = 8 + torch.rand(n_gt, generator=g) * 192
        h = 8 + torch.rand(n_gt, generator=g) * 192
        gt = torch.stack([x1, y1, (x1 + w).clamp(max=640), (y1 + h).clamp(max=480)], 1)
        gl = torch.randint(0, n_cls, (n_gt,), generator=g)
        crowd = (torch.rand(n_gt, generator=g) < crowd_frac).long()
        n_jit = min(n_gt, n_det)
        jit = gt[:n_jit] + torch.randn(n_jit, 4, generator=g) * 0.1 * torch.stack([w, h, w, h], 1)[:n_jit]
        jl = torch.where(torch.rand(n_jit, generator=g) < 0.9, gl[:n_jit], torch.randint(0, n_cls, (n_jit,), generator=g))
        n_rand = n_det - n_jit
        rx = torch.rand(n_rand, generator=g) * 540
        ry = torch.rand(n_rand, generator=g) * 380
        rnd = torch.stack([rx, ry, rx + 8 + torch.rand(n_rand, generator=g) * 192, ry + 8 + torch.rand(n_rand, generator=g) * 192], 1)
        boxes = torch.cat([jit, rnd])
        boxes = torch.stack([boxes[:, 0].clamp(0, 639), boxes[:, 1].clamp(0, 479), boxes[:, 2], boxes[:, 3]], 1)
        boxes[:, 2] = torch.maximum(boxes[:, 2], boxes[:, 0] + 1)
        boxes[:, 3] = torch.maximum(boxes[:, 3], boxes[:, 1] + 1)
        labels = torch.cat([jl, torch.randint(0, n_cls, (n_rand,), generator=g)])
        scores = torch.rand(n_det, generator=g)
        if dup_scores:
            scores = (scores * 20).floor() / 20
        preds.append({"boxes": boxes, "scores": scores, "labels": labels})
        t = {"boxes": gt, "labels": gl}
        if crowd_frac > 0:
            t["iscrowd"] = crowd
        target.append(t)
    return preds, target


LEGACY_MAP_CASES = {
    "small": dict(seed=5, n_img=12, n_gt=6, n_det=20, n_cls=4),
    "mid": dict(seed=6, n_img=60, n_gt=10, n_det=40, n_cls=8),
    "dup": dict(seed=7, n_img=30, n_gt=8, n_det=30, n_cls=5, dup_scores=True),
}


def det_to_numpy(preds, target):
    """list-of-dict torch inputs -> the per-image numpy lists oracle.coco_map.coco_evaluate takes"""
    kw = dict(
        det_boxes=[p["boxes"].numpy() for p in preds], det_scores=[p["scores"].numpy() for p in preds],
        det_labels=[p["labels"].numpy() for p in preds], gt_boxes=[t["boxes"].numpy() for t in target],
        gt_labels=[t["labels"].numpy() for t in target],
    )
    if any("iscrowd" in t for t in target):
        kw["gt_crowds"] = [t.get("iscrowd", torch.zeros_like(t["labels"])).numpy() for t in target]
    if any("area" in t for t in target):
        kw["gt_areas"] = [t.get("area", torch.zeros_like(t["labels"])).numpy() for t in target]
    return kw
