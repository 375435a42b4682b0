#!/usr/bin/env python
"""Round-2 scaling diagnostic (torchrun, N ranks): 20-step windows of the cfg2 update loop exactly as bench.py times them,
repeated; every rank's own device time and host time per window, gathered — is the N-GPU slowdown systematic (every rank,
every window) or the maximum over occasionally late ranks?"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from metrics_b200.classification import MulticlassConfusionMatrix

    n, c, nrot = 65536, 1000, 16
    batches = []
    for i in range(nrot):
        g = torch.Generator(device=dev).manual_seed(1000 * rank + i)
        batches.append((torch.randn(n, c, generator=g, device=dev).bfloat16(), torch.randint(0, c, (n,), generator=g, device=dev)))
    m = MulticlassConfusionMatrix(num_classes=c, validate_args=False).to(dev)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.3:
        for i in range(64):
            m.update(*batches[i % nrot])
        torch.cuda.synchronize()
    order = [batches[i % nrot] for i in range(20)]
    upd = m.update
    reps = 25
    devms, hostus = [], []
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        with torch.no_grad():
            for lg, tg in order:
                upd(lg, tg)
        h1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        devms.append(e0.elapsed_time(e1) / 20 * 1e3)
        hostus.append((h1 - h0) / 20 * 1e6)
    mine = torch.tensor([devms, hostus], dtype=torch.float64, device=dev)
    slab = torch.empty((world, 2, reps), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(slab, mine)
    if rank == 0:
        s = slab.cpu()
        out = {"world": world, "per_rank_dev_us_median": [round(float(s[r, 0].median()), 2) for r in range(world)],
               "per_rank_dev_us_min": [round(float(s[r, 0].min()), 2) for r in range(world)],
               "per_rank_host_us_median": [round(float(s[r, 1].median()), 2) for r in range(world)],
               "max_over_ranks_per_window_us": [round(float(x), 2) for x in s[:, 0].max(dim=0).values],
               "median_of_max_over_ranks_us": round(float(s[:, 0].max(dim=0).values.median()), 2)}
        print(json.dumps(out, indent=1))
        open(os.path.join(ROOT, "gpurun_out", f"r2_diag_scale_{world}gpu.json"), "w").write(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
