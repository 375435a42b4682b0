#!/usr/bin/env python
"""Timings of the BASELINE.json configs other than the bench.py headline (cfg2): cfg1, cfg3, cfg4, cfg5.

    python benchmarks/run_configs.py [--out gpurun_out/configs.json]            # 1 GPU
    torchrun --nproc-per-node N ... benchmarks/run_configs.py --only cfg5      # cfg5 sync at N ranks

Every GPU number is device time between CUDA events (after warm-up); CPU legs use the reference's op chain restated
in oracle/torch_cpu_chain.py where one exists.  Parity is asserted against goldens/oracle in tests/, not here.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def ev_time(fn, reps=5, warm=2):
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
    return min(out), sorted(out)[len(out) // 2]


def cfg1(dev):
    from metrics_b200.classification import MulticlassAccuracy
    from tests.helpers import cfg1_inputs

    preds, target = cfg1_inputs()
    preds, target = preds.to(dev), target.to(dev)
    res = {}
    for validate in (True, False):
        m = MulticlassAccuracy(num_classes=5, validate_args=validate).to(dev)

        def run():
            m.reset()
            for i in range(100):
                m.update(preds[i], target[i])
            return m.compute()

        run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        val = run()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        res[f"validate_{validate}"] = {"wall_us_per_update": wall / 100 * 1e6, "value": float(val),
                                       "units_per_s": 512000 / wall}
    # the reference's CPU op chain on the same tensors, timed beside it (validate_args=False equivalent)
    from oracle.torch_cpu_chain import macro_accuracy_cpu, multiclass_stat_scores_update_cpu

    cp, ct = preds.cpu(), target.cpu()
    torch.set_num_threads(min(os.cpu_count() or 1, 8))

    def cpu_run():
        st = [torch.zeros(5, dtype=torch.long) for _ in range(4)]
        for i in range(100):
            multiclass_stat_scores_update_cpu(*st, cp[i], ct[i], 5)
        return macro_accuracy_cpu(*st)

    cpu_run()
    t0 = time.perf_counter()
    cval = cpu_run()
    wall = time.perf_counter() - t0
    res["cpu_chain"] = {"wall_us_per_update": wall / 100 * 1e6, "value": float(cval), "threads": torch.get_num_threads(),
                        "what": "reference op chain (argmax, bincount C^2, diag/row/col sums) without the Metric wrapper"}
    return res


def cfg3(dev):
    from metrics_b200 import MetricCollection, _native
    from metrics_b200.classification import BinaryAUROC, BinaryAveragePrecision
    from tests.helpers import cfg3_inputs

    preds, target = cfg3_inputs()
    dp, dt = preds.to(dev), target.to(dev)
    mc = MetricCollection([BinaryAUROC(validate_args=False), BinaryAveragePrecision(validate_args=False)]).to(dev)

    def updates():
        mc.reset()
        for i in range(1000):
            mc.update(dp[i], dt[i])

    upd_min, upd_med = ev_time(updates, reps=3, warm=1)
    comp_min, comp_med = ev_time(lambda: (mc.__setattr__("_dummy", None), [setattr(m, "_computed", None) for m in mc.values(copy_state=False)], mc.compute()), reps=5, warm=1)
    flat_p, flat_t = dp.reshape(-1), dt.reshape(-1)
    k_min, _ = ev_time(lambda: _native.curve_evaluate(flat_p, flat_t), reps=10, warm=3)
    res = mc.compute()
    out = {"update_phase_ms": upd_min, "us_per_update": upd_min, "compute_ms": comp_min, "curve_evaluate_1e7_ms": k_min,
           "auroc": float(res["BinaryAUROC"]), "ap": float(res["BinaryAveragePrecision"]),
           "samples_per_s_end_to_end": 1e7 / ((upd_min + comp_min) * 1e-3),
           "roofline_compute_only": {"algorithmic_bytes": 150e6, "achieved_gbs": 150e6 / (k_min * 1e-3) / 1e9}}
    # CPU: the reference's compute chain on the concatenated 1e7 samples (one repetition: it takes seconds)
    from oracle.torch_cpu_chain import binary_auroc_ap_compute_cpu

    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    fp, ft = preds.reshape(-1), target.reshape(-1)
    t0 = time.perf_counter()
    a, p = binary_auroc_ap_compute_cpu(fp, ft)
    out["cpu_compute_chain_ms"] = (time.perf_counter() - t0) * 1e3
    out["cpu_threads"] = torch.get_num_threads()
    out["cpu_auroc"], out["cpu_ap"] = float(a), float(p)
    return out


def cfg4(dev):
    from metrics_b200.detection import MeanAveragePrecision
    from tests.helpers import synth_detection

    preds, target = synth_detection(seed=0, n_img=5000, n_gt=20, n_det=100, n_cls=80, crowd_frac=0.02)
    to = lambda items: [{k: v.to(dev) for k, v in d.items()} for d in items]  # noqa: E731
    preds, target = to(preds), to(target)
    m = MeanAveragePrecision().to(dev)
    m.warn_on_many_detections = False

    def updates():
        m.reset()
        for i in range(0, 5000, 100):
            m.update(preds[i:i + 100], target[i:i + 100])

    updates()  # untimed warm-up: the first pass pays one-off costs (first use of the cat / stack kernels, allocator growth)
    torch.cuda.synchronize()
    upd_walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        updates()
        torch.cuda.synchronize()
        upd_walls.append(time.perf_counter() - t0)
    upd_wall = min(upd_walls)
    vals = []
    walls = []
    for _ in range(3):
        m._computed = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = m.compute()
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
        vals.append(float(r["map"]))
    return {"update_phase_s_wall": upd_wall, "update_phase_s_wall_all": upd_walls, "compute_s_wall_min": min(walls), "map": vals[-1],
            "detections_per_s_end_to_end": 500000 / (upd_wall + min(walls)), "images_per_s_end_to_end": 5000 / (upd_wall + min(walls)),
            "note": "reference CPU path not runnable anywhere (pycocotools absent); oracle is a Python restatement, far too slow to time at this size"}


def cfg5(dev, rank, world):
    import bench

    return bench.leg_cfg5(dev, rank, world)


# --------------------------------------------------------------------------------------------------------------
# bench.py --config cfg3|cfg4|cfg5: the same line format as the cfg2 headline
# --------------------------------------------------------------------------------------------------------------
_LINES = {
    "cfg3": ("samples/sec (BinaryAUROC + BinaryAveragePrecision: 1000 update() calls of 10,000 samples, then compute())", "samples/s", 10_000_000),
    "cfg4": ("detections/sec (MeanAveragePrecision bbox: 50 update() calls of 100 images x 100 detections, then compute())", "detections/s", 500_000),
    "cfg5": ("metric-updates/sec (MetricCollection([MulticlassF1Score, MulticlassAUROC], C=1000): 4 update() calls of [4096,1000] per rank, then compute())", "updates/s", 4 * 4096 * 1000),
}


def leg_cfg3(dev, steps: int = 3):
    """cfg3 through the public API, `steps` full passes; device time of updates and compute separately, and end to end from
    pinned host batches (H2D per update, scalar results read back)."""
    from metrics_b200 import MetricCollection, _native
    from metrics_b200.classification import BinaryAUROC, BinaryAveragePrecision

    g = torch.Generator().manual_seed(0)
    preds = torch.rand(1000, 10000, generator=g)
    target = torch.randint(0, 2, (1000, 10000), generator=g)
    dp, dt = preds.to(dev), target.to(dev)
    mc = MetricCollection([BinaryAUROC(validate_args=False), BinaryAveragePrecision(validate_args=False)]).to(dev)

    def updates():
        mc.reset()
        for i in range(1000):
            mc.update(dp[i], dt[i])

    def compute():
        for m in mc.values(copy_state=False):
            m._computed = None
            if hasattr(m, "_group_cache"):
                m._group_cache.clear()
        return mc.compute()

    upd_min, upd_med = ev_time(updates, reps=steps, warm=1)
    cmp_min, cmp_med = ev_time(compute, reps=max(steps, 3), warm=1)
    flat_p, flat_t = dp.reshape(-1), dt.reshape(-1)
    k_min, _ = ev_time(lambda: _native.curve_evaluate(flat_p, flat_t), reps=10, warm=3)
    res = compute()
    # end to end: pinned host batches -> device, per update; two scalars back
    pp, pt = preds.pin_memory(), target.pin_memory()

    def e2e():
        mc.reset()
        for i in range(1000):
            mc.update(pp[i].to(dev, non_blocking=True), pt[i].to(dev, non_blocking=True))
        r = compute()
        return float(r["BinaryAUROC"]), float(r["BinaryAveragePrecision"])

    e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e()
    e2e_s = time.perf_counter() - t0
    return {"update_phase_ms": upd_min, "update_phase_ms_median": upd_med, "us_per_update": upd_min, "compute_ms": cmp_min,
            "compute_ms_median": cmp_med, "curve_evaluate_1e7_ms": k_min, "auroc": float(res["BinaryAUROC"]),
            "ap": float(res["BinaryAveragePrecision"]), "samples_per_s_device_resident": 1e7 / ((upd_min + cmp_min) * 1e-3),
            "e2e_samples_per_s": 1e7 / e2e_s, "e2e_h2d_bytes": 1e7 * 12, "e2e_d2h_bytes": 8,
            "roofline_compute_only": {"algorithmic_bytes": 150e6, "achieved_gbs": 150e6 / (k_min * 1e-3) / 1e9}}


def aten_cfg3(dev):
    """The reference's compute chain for cfg3 on CUDA tensors (two `_binary_clf_curve` sorts, roc.py:53 and
    precision_recall_curve.py:275), via baseline/_ref when present, else the device-agnostic restatement."""
    import bench

    g = torch.Generator().manual_seed(0)
    p = torch.rand(10_000_000, generator=g).to(dev)
    t = torch.randint(0, 2, (10_000_000,), generator=g).to(dev)
    if bench.have_reference():
        tm = bench.import_reference()
        import torchmetrics.functional.classification as RF

        fn = lambda: (RF.binary_auroc(p, t, validate_args=False), RF.binary_average_precision(p, t, validate_args=False))  # noqa: E731
        kind = f"the unmodified reference (baseline/_ref, TorchMetrics {tm.__version__}) binary_auroc + binary_average_precision on CUDA tensors"
    else:
        from oracle.torch_cpu_chain import binary_auroc_ap_compute_cpu

        fn = lambda: binary_auroc_ap_compute_cpu(p, t)  # noqa: E731
        kind = "reference op chain (oracle/torch_cpu_chain.py) on CUDA tensors"
    mn, med = ev_time(fn, reps=5, warm=2)
    srt, _ = ev_time(lambda: torch.sort(p, descending=True), reps=10, warm=3)
    a, b = fn()
    return {"compute_ms": mn, "compute_ms_median": med, "torch_sort_1e7_ms": srt, "auroc": float(a), "ap": float(b), "kind": kind}


def leg_cfg4(dev):
    r = cfg4(dev)
    r["detections_per_s"] = r["detections_per_s_end_to_end"]
    return r


def bench_line(args, dev, rank, world, sampler):
    """One JSON line for --config cfg3|cfg4|cfg5 (bench.py prints it)."""
    metric, unit, units = _LINES[args.config]
    sampler.wait_ready()
    w0 = time.time()
    if args.config == "cfg3":
        leg = leg_cfg3(dev, steps=max(1, min(args.steps, 5)))
        step_ms = leg["update_phase_ms"] + leg["compute_ms"]
        value, e2e = units / (step_ms * 1e-3), {"value": leg["e2e_samples_per_s"], "unit": unit,
                                                "h2d_bytes_per_step": leg["e2e_h2d_bytes"], "d2h_bytes_per_step": leg["e2e_d2h_bytes"]}
        leg["aten_gpu_baseline"] = aten_cfg3(dev)
    elif args.config == "cfg4":
        leg = leg_cfg4(dev)
        step_ms = (leg["update_phase_s_wall"] + leg["compute_s_wall_min"]) * 1e3
        value = units / (step_ms * 1e-3)
        e2e = {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 12 * 4,
               "note": "wall clock of the public API with device-resident per-image dicts (the reference's input format)"}
    else:
        import bench

        leg = bench.leg_cfg5(dev, rank, world)
        step_ms = leg["update_ms_4_batches"] + leg["compute_ms"]
        value = world * units / (step_ms * 1e-3)
        e2e = {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8}
    sampler.window("value", w0, time.time())
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": f"BASELINE.json {args.config}", **leg},
            "e2e": e2e, "clocks": sampler.stop()}


def reference_line(args):
    """--impl reference for cfg3/cfg5: the reference's CPU implementation (baseline/_ref when present, else the op-chain
    port) on the host cores; cfg4 has no runnable reference (pycocotools absent)."""
    import bench

    metric, unit, units = _LINES[args.config]
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    kind = "reference" if bench.have_reference() else "port"
    if args.config == "cfg3":
        g = torch.Generator().manual_seed(0)
        preds = torch.rand(1000, 10000, generator=g)
        target = torch.randint(0, 2, (1000, 10000), generator=g)
        if kind == "reference":
            tm = bench.import_reference()
            mc = tm.MetricCollection([tm.classification.BinaryAUROC(validate_args=False),
                                      tm.classification.BinaryAveragePrecision(validate_args=False)])
            t0 = time.perf_counter()
            for i in range(1000):
                mc.update(preds[i], target[i])
            mc.compute()
            dt = time.perf_counter() - t0
        else:
            from oracle.torch_cpu_chain import binary_auroc_ap_compute_cpu

            t0 = time.perf_counter()
            binary_auroc_ap_compute_cpu(preds.reshape(-1), target.reshape(-1))
            dt = time.perf_counter() - t0
        sample = "one full pass: 1000 updates of 10,000 samples + compute()"
    elif args.config == "cfg5":
        from oracle.torch_cpu_chain import multiclass_auroc_compute_cpu, multiclass_stat_scores_update_cpu

        torch.manual_seed(0)
        cb = [(torch.randn(4096, 1000), torch.randint(0, 1000, (4096,))) for _ in range(4)]
        kind = "port"
        t0 = time.perf_counter()
        st = [torch.zeros(1000, dtype=torch.long) for _ in range(4)]
        probs = []
        for lg, tg in cb:
            multiclass_stat_scores_update_cpu(*st, lg, tg, 1000)
            probs.append(torch.softmax(lg, 1))
        multiclass_auroc_compute_cpu(torch.cat(probs), torch.cat([tg for _, tg in cb]), 1000)
        dt = time.perf_counter() - t0
        sample = "one rank's share: 4 updates of [4096,1000] + compute() (1000 one-vs-rest sorts)"
    else:
        return {"impl": "reference", "unavailable": "cfg4: the reference needs pycocotools / faster-coco-eval (absent, no network)"}
    value = units / dt
    return {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": f"BASELINE.json {args.config}", "device": "cpu"},
            "cpu_baseline": {"value": value, "unit": unit, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    want = [w for w in args.only.split(",") if w] or (["cfg1", "cfg3", "cfg4", "cfg5"] if world == 1 else ["cfg5"])
    out = {"host_cpus": os.cpu_count(), "gpu": torch.cuda.get_device_name(dev), "world": world}
    if "cfg1" in want:
        out["cfg1"] = cfg1(dev)
    if "cfg3" in want:
        out["cfg3"] = cfg3(dev)
    if "cfg4" in want:
        out["cfg4"] = cfg4(dev)
    if "cfg5" in want:
        out["cfg5"] = cfg5(dev, rank, world)
    if rank == 0:
        text = json.dumps(out, indent=1)
        print(text)
        if args.out:
            open(args.out, "w").write(text)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
