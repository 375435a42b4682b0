#!/usr/bin/env python
"""Round-2 sync diagnostics, run under torchrun (N >= 2): what do the bare NCCL collectives cost next to the metric-level
`compute()` that wraps them, and does torch's symmetric memory (peer pointers over NVLink) come up on this box?

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 benchmarks/diag_sync_r2.py
"""
from __future__ import annotations

import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def ev(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        fn()
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1), (t1 - t0) * 1e3))
    return {"dev_min_ms": min(o[0] for o in out), "dev_med_ms": statistics.median(o[0] for o in out),
            "host_min_ms": min(o[1] for o in out)}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    # bare collectives
    x = torch.ones(1000 * 1000, dtype=torch.int64, device=dev)
    out["nccl_all_reduce_8MB_i64"] = ev(lambda: dist.all_reduce(x))
    xf = torch.ones(2 * 1000 * 1000, dtype=torch.float32, device=dev)
    out["nccl_all_reduce_8MB_f32"] = ev(lambda: dist.all_reduce(xf))
    small = torch.ones(16, dtype=torch.int64, device=dev)
    out["nccl_all_reduce_128B"] = ev(lambda: dist.all_reduce(small))
    n_keys = 1000 * 16384
    send = torch.zeros(n_keys, dtype=torch.int32, device=dev)
    recv = torch.empty(n_keys, dtype=torch.int32, device=dev)
    out["nccl_all_to_all_65MB"] = ev(lambda: dist.all_to_all_single(recv, send))
    big = torch.empty(world * n_keys, dtype=torch.int32, device=dev)
    out["nccl_all_gather_65MB_per_rank"] = ev(lambda: dist.all_gather_into_tensor(big, send))

    # metric-level compute (cfg2 state) and its pieces
    from metrics_b200.classification import MulticlassConfusionMatrix

    m = MulticlassConfusionMatrix(num_classes=1000, validate_args=False).to(dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    lg = torch.randn(65536, 1000, generator=g, device=dev).bfloat16()
    tg = torch.randint(0, 1000, (65536,), generator=g, device=dev)
    m.update(lg, tg)

    def compute():
        m._computed = None
        return m.compute()

    out["confmat_compute"] = ev(compute)

    def sync_only():
        m.sync()
        m.unsync()

    out["confmat_sync_unsync"] = ev(sync_only)

    # symmetric memory
    try:
        import torch.distributed._symmetric_memory as symm

        t = symm.empty(1000 * 1000, dtype=torch.int64, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        t.fill_(rank + 1)
        hdl.barrier()
        peer = (rank + 1) % world
        pt = hdl.get_buffer(peer, (1000 * 1000,), torch.int64)
        ok = bool((pt == peer + 1).all())
        out["symm"] = {"ok": ok, "ptrs": [hex(p) for p in hdl.buffer_ptrs], "multicast": hex(hdl.multicast_ptr or 0),
                       "signal_pad_size": hdl.signal_pad_size, "barrier": ev(lambda: hdl.barrier())}
        dst = torch.empty_like(t)
        out["symm"]["peer_read_8MB"] = ev(lambda: dst.copy_(pt))
        out["symm"]["peer_write_8MB"] = ev(lambda: pt.copy_(dst))
        from metrics_b200 import parallel_sync

        src = torch.full((1000 * 1000,), rank + 1, dtype=torch.int64, device=dev)
        res = parallel_sync._peer_all_reduce([src], torch.int64, dist.ReduceOp.SUM, dist.group.WORLD)
        torch.cuda.synchronize()
        out["symm"]["own_allreduce_ok"] = bool(res is not None and (res == world * (world + 1) // 2).all())
        out["symm"]["own_allreduce_8MB"] = ev(lambda: parallel_sync._peer_all_reduce([src], torch.int64, dist.ReduceOp.SUM, dist.group.WORLD))
        out["confmat_compute_peer"] = ev(compute)
        os.environ["MB200_PEER_EXCHANGE"] = "0"
        out["confmat_compute_nccl"] = ev(compute)
        os.environ["MB200_PEER_EXCHANGE"] = "1"
        # cfg5 compute: peer-memory exchange vs NCCL all_to_all exchange vs gather-everything
        import bench

        leg = {}
        for name, env in (("peer", {"MB200_PEER_EXCHANGE": "1"}), ("nccl_all_to_all", {"MB200_PEER_EXCHANGE": "0"})):
            os.environ.update(env)
            leg[name] = bench.leg_cfg5(dev, rank, world)
        os.environ["MB200_PEER_EXCHANGE"] = "1"
        out["cfg5"] = {k: {kk: v[kk] for kk in ("update_ms_4_batches", "compute_ms", "compute_ms_median", "compute_ms_gather_everything", "auroc", "parity")} for k, v in leg.items()}
    except Exception as err:  # pragma: no cover
        import traceback

        out["symm"] = {"error": repr(err), "tb": traceback.format_exc()[-1500:]}

    if rank == 0:
        print(json.dumps(out, indent=1))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", f"r2_diag_sync_{world}gpu.json"), "w").write(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
