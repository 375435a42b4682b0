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

    t0 = time.perf_counter()
    updates()
    torch.cuda.synchronize()
    upd_wall = time.perf_counter() - t0
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
    return {"update_phase_s_wall": upd_wall, "compute_s_wall_min": min(walls), "map": vals[-1],
            "detections_per_s_end_to_end": 500000 / (upd_wall + min(walls)), "images_per_s_end_to_end": 5000 / (upd_wall + min(walls)),
            "note": "reference CPU path not runnable anywhere (pycocotools absent); oracle is a Python restatement, far too slow to time at this size"}


def cfg5(dev, rank, world):
    from metrics_b200 import MetricCollection
    from metrics_b200.classification import MulticlassAUROC, MulticlassF1Score
    from tests.helpers import cfg5_rank_batches

    batches = [(lg.to(dev), tg.to(dev)) for lg, tg in cfg5_rank_batches(rank, 4)]
    mc = MetricCollection([MulticlassF1Score(num_classes=1000, validate_args=False),
                           MulticlassAUROC(num_classes=1000, validate_args=False)]).to(dev)

    def updates():
        mc.reset()
        for lg, tg in batches:
            mc.update(lg, tg)

    upd_min, _ = ev_time(updates, reps=5, warm=2)

    def compute():
        for m in mc.values(copy_state=False):
            m._computed = None
            if hasattr(m, "_group_cache"):
                m._group_cache.clear()
        return mc.compute()

    comp_min, comp_med = ev_time(compute, reps=5, warm=2)
    comp_gather_min = None
    if world > 1:  # A/B: the reference's gather-everything sync followed by an all-class evaluation on every rank
        os.environ["MB200_SHARDED_CURVES"] = "0"
        comp_gather_min, _ = ev_time(compute, reps=5, warm=2)
        os.environ["MB200_SHARDED_CURVES"] = "1"
    # sync only
    def sync_only():
        for m in mc.values(copy_state=False):
            m.sync()
            m.unsync()

    sync_min = None
    if world > 1:
        sync_min, _ = ev_time(sync_only, reps=5, warm=2)
    res = compute()
    cpu = None
    if rank == 0 and world == 1:  # the reference's CPU chain for one rank's share, timed beside it
        from oracle.torch_cpu_chain import multiclass_auroc_compute_cpu, multiclass_stat_scores_update_cpu

        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        cb = [(lg.cpu(), tg.cpu()) for lg, tg in batches]
        t0 = time.perf_counter()
        st = [torch.zeros(1000, dtype=torch.long) for _ in range(4)]
        probs = []
        for lg, tg in cb:
            multiclass_stat_scores_update_cpu(*st, lg, tg, 1000)
            probs.append(torch.softmax(lg, 1))  # normalize_logits_if_needed on logits
        t_upd = time.perf_counter() - t0
        t0 = time.perf_counter()
        auc = multiclass_auroc_compute_cpu(torch.cat(probs), torch.cat([tg for _, tg in cb]), 1000)
        t_cmp = time.perf_counter() - t0
        cpu = {"update_ms_4_batches": t_upd * 1e3, "compute_ms": t_cmp * 1e3, "auroc": float(auc),
               "threads": torch.get_num_threads()}
    t = torch.tensor([upd_min, comp_min, sync_min or 0.0, comp_gather_min or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return {"world": world, "update_ms_4_batches": float(t[0]), "update_units_per_s_per_gpu": 4 * 4096 * 1000 / (float(t[0]) * 1e-3),
            "compute_ms": float(t[1]), "sync_only_ms": float(t[2]) if world > 1 else None,
            "compute_ms_gather_everything": float(t[3]) if world > 1 else None,
            "compute_path": "class-sharded all_to_all (metrics_b200/parallel_curves.py)" if world > 1 else "local",
            "f1": float(res["MulticlassF1Score"]), "auroc": float(res["MulticlassAUROC"]), "cpu_chain_one_rank": cpu,
            "sync_bytes_per_rank": 16384 * 1000 * 4 + 16384 * 8 + 4 * 1000 * 8}


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
