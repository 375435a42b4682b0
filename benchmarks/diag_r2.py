#!/usr/bin/env python
"""Round-2 diagnostics (one B200): where do the 20-step windows of bench.py lose time, and what does the stock ATen op
chain the reference would execute ON THE SAME B200 cost (SURVEY.md §2.2's bar) for cfg2 / cfg3 / the sort alone.

    python benchmarks/diag_r2.py [--out gpurun_out/r2_diag.json]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def ev_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return {"min_ms": min(out), "median_ms": statistics.median(out)}


def windows(dev):
    """Distribution of K-step windows exactly as bench.py times them, plus the host time of the same K calls."""
    from metrics_b200.classification import MulticlassConfusionMatrix

    n, c, nrot = 65536, 1000, 16
    batches = []
    for i in range(nrot):
        g = torch.Generator(device=dev).manual_seed(i)
        batches.append((torch.randn(n, c, generator=g, device=dev).bfloat16(),
                        torch.randint(0, c, (n,), generator=g, device=dev)))
    m = MulticlassConfusionMatrix(num_classes=c, validate_args=False).to(dev)
    for i in range(200):
        m.update(*batches[i % nrot])
    torch.cuda.synchronize()
    res = {}
    for k in (20, 100, 2000):
        dev_ms, host_us = [], []
        for rep in range(15 if k <= 100 else 3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            for i in range(k):
                m.update(*batches[i % nrot])
            t1 = time.perf_counter()
            e1.record()
            torch.cuda.synchronize()
            dev_ms.append(e0.elapsed_time(e1) / k * 1e3)
            host_us.append((t1 - t0) / k * 1e6)
        res[f"k{k}"] = {"dev_us_per_step": [round(x, 2) for x in dev_ms], "host_us_per_call": [round(x, 2) for x in host_us]}
    # per-step events inside one 20-step window: when does each kernel finish relative to the first event?
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    evs[0].record()
    for i in range(20):
        m.update(*batches[i % nrot])
        evs[i + 1].record()
    torch.cuda.synchronize()
    res["per_step_finish_us"] = [round(evs[0].elapsed_time(evs[i + 1]) * 1e3, 1) for i in range(20)]
    return res


def aten_cfg2(dev):
    """The reference's op chain for MulticlassConfusionMatrix.update on CUDA tensors (confusion_matrix.py:297-328)."""
    n, c = 65536, 1000
    g = torch.Generator(device=dev).manual_seed(0)
    batches = [(torch.randn(n, c, generator=g, device=dev).bfloat16(), torch.randint(0, c, (n,), generator=g, device=dev))
               for _ in range(4)]
    confmat = torch.zeros(c, c, dtype=torch.long, device=dev)
    state = {"i": 0}

    def step():
        lg, tg = batches[state["i"] % 4]
        state["i"] += 1
        p = lg.argmax(dim=1).flatten()
        t = tg.flatten()
        um = t.to(torch.long) * c + p.to(torch.long)
        bins = torch.bincount(um, minlength=c * c)
        confmat.add_(bins.reshape(c, c))

    def steps16():
        for _ in range(16):
            step()

    r = ev_ms(steps16, reps=8, warm=2)
    return {"us_per_update": r["min_ms"] / 16 * 1e3, "median_us_per_update": r["median_ms"] / 16 * 1e3,
            "what": "argmax -> t*C+p -> bincount(minlength=C^2) -> confmat += (stock ATen on cuda, validate_args=False)"}


def aten_cfg3(dev):
    import torch.nn.functional as F

    g = torch.Generator(device=dev).manual_seed(0)
    p = torch.rand(10_000_000, generator=g, device=dev)
    t = torch.randint(0, 2, (10_000_000,), generator=g, device=dev)

    def clf_curve():
        idx = torch.argsort(p, descending=True)
        ps, ts = p[idx], t[idx]
        distinct = torch.where(ps[1:] - ps[:-1])[0]
        thr_idx = F.pad(distinct, [0, 1], value=ts.size(0) - 1)
        ts = (ts == 1).to(torch.long)
        tps = torch.cumsum(ts * 1.0, dim=0)[thr_idx]
        fps = 1 + thr_idx - tps
        return fps, tps, ps[thr_idx]

    def both():
        fps, tps, _ = clf_curve()
        tps2 = torch.cat([torch.zeros(1, dtype=tps.dtype, device=dev), tps])
        fps2 = torch.cat([torch.zeros(1, dtype=fps.dtype, device=dev), fps])
        auroc = torch.trapz(tps2 / tps2[-1], fps2 / fps2[-1])
        fps, tps, _ = clf_curve()
        precision = tps / (tps + fps)
        recall = tps / tps[-1]
        precision = torch.cat([precision.flip(0), torch.ones(1, device=dev)])
        recall = torch.cat([recall.flip(0), torch.zeros(1, device=dev)])
        ap = -torch.sum((recall[1:] - recall[:-1]) * precision[:-1])
        return auroc, ap

    out = {"sort_f32_1e7": ev_ms(lambda: torch.sort(p, descending=True), reps=10, warm=3),
           "argsort_f32_1e7": ev_ms(lambda: torch.argsort(p, descending=True), reps=10, warm=3),
           "clf_curve_once": ev_ms(clf_curve, reps=5, warm=2),
           "auroc_plus_ap_compute": ev_ms(both, reps=5, warm=2)}
    # our evaluation on the same tensors
    from metrics_b200 import _native

    out["ours_curve_evaluate_1e7"] = ev_ms(lambda: _native.curve_evaluate(p, t), reps=10, warm=3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {"gpu": torch.cuda.get_device_name(dev), "host_cpus": os.cpu_count()}
    out["windows"] = windows(dev)
    out["aten_cfg2"] = aten_cfg2(dev)
    out["aten_cfg3"] = aten_cfg3(dev)
    try:
        import torch.distributed._symmetric_memory as symm  # noqa: F401

        out["symmetric_memory_importable"] = True
    except Exception as err:  # pragma: no cover
        out["symmetric_memory_importable"] = repr(err)
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        open(args.out, "w").write(text)


if __name__ == "__main__":
    main()
