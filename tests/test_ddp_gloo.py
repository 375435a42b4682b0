"""CPU, world_size 2 over gloo: cross-rank state sync (reference behaviours of tests/unittests/bases/test_ddp.py).

One spawn runs every scenario in both ranks (process start-up dominates the cost).  Covers the bucketed fast path of
`Metric.sync` (integer all-reduce buckets, rank-ordered float gather, cat states incl. ragged / empty ranks, None
reductions), the plug-in `gather_all_tensors` contract, sync/unsync semantics and `MetricCollection` under DDP.
"""
import os
import socket
import traceback

import pytest
import torch
import torch.multiprocessing as mp

WORLD = 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scenarios(rank: int) -> None:
    import torch.distributed as dist

    from metrics_b200 import MetricCollection
    from metrics_b200.utilities.distributed import gather_all_tensors
    from metrics_b200.utilities.exceptions import TorchMetricsUserError
    from tests.dummies import DummyCat, DummyIntStates, DummyMean, DummyNone, DummySum

    # ---- gather_all_tensors: equal shapes, ragged 1-D, ragged multi-dim, scalar, own-rank identity ----------------
    t = torch.ones(3) * (rank + 1)
    out = gather_all_tensors(t)
    assert [o.tolist() for o in out] == [[1.0] * 3, [2.0] * 3] and out[rank] is t
    rag = torch.arange(rank + 2, dtype=torch.float32)
    out = gather_all_tensors(rag)
    assert [o.shape[0] for o in out] == [2, 3] and out[1].tolist() == [0.0, 1.0, 2.0]
    rag2 = torch.full((rank + 1, 2, rank + 2), float(rank))
    out = gather_all_tensors(rag2)
    assert [tuple(o.shape) for o in out] == [(1, 2, 2), (2, 2, 3)] and float(out[1].sum()) == 12.0
    sc = gather_all_tensors(torch.tensor(float(rank)))
    assert [float(s) for s in sc] == [0.0, 1.0]

    # ---- sum (float) + unsync restores the local value ------------------------------------------------------------------
    m = DummySum()
    m.update(float(rank + 1))
    assert float(m.compute()) == 3.0 and float(m.x) == rank + 1 and not m._is_synced

    # ---- integer buckets: sum / max / min in two collectives ------------------------------------------------------------
    mi = DummyIntStates(n=3)
    mi.update([1 + rank, 2, 3 - rank])
    res = mi.compute()
    assert res[0].tolist() == [3, 4, 5] and res[3].tolist() == [12, 16, 20]
    assert res[4].tolist() == [2, 2, 3] and res[5].tolist() == [1, 2, 2]
    assert mi.tp.tolist() == [1 + rank, 2, 3 - rank]

    # ---- mean ---------------------------------------------------------------------------------------------------------------
    mm = DummyMean()
    mm.update(2.0 * (rank + 1))
    assert float(mm.compute()) == 3.0

    # ---- cat: equal lengths, ragged lengths, empty list on one rank, multi-dim -------------------------------------------
    c = DummyCat()
    c.update(torch.tensor([1.0 + rank, 2.0 + rank]))
    vals, ids = c.compute()
    assert vals.tolist() == [1.0, 2.0, 2.0, 3.0] and ids.tolist() == [0, 1, 0, 1]
    assert isinstance(c.vals, list) and len(c.vals) == 1  # restored
    c2 = DummyCat()
    c2.update(torch.arange(3 + 2 * rank, dtype=torch.float32))
    c2.update(torch.tensor([9.0]))
    vals, _ = c2.compute()
    assert vals.tolist() == [0, 1, 2, 9, 0, 1, 2, 3, 4, 9]
    c3 = DummyCat()
    if rank == 0:
        c3.update(torch.tensor([5.0, 6.0]))
    c3._update_count = 1
    vals, _ = c3.compute()
    assert vals.tolist() == [5.0, 6.0]
    c4 = DummyCat()
    c4.update(torch.full((2 + rank, 3), float(rank)), ids=torch.arange(2 + rank))
    vals, _ = c4.compute()
    assert tuple(vals.shape) == (5, 3) and vals[:2].sum() == 0 and vals[2:].sum() == 9

    # ---- dist_reduce_fx=None: stacked tensor / flattened list -------------------------------------------------------------
    n = DummyNone()
    n.update([1.0 * rank, 2.0])
    tt, ll = n.compute()
    assert tuple(tt.shape) == (2, 2) and tt[:, 0].tolist() == [0.0, 1.0] and len(ll) == 2

    # ---- sync / unsync protocol, state_dict while synced --------------------------------------------------------------------
    s = DummySum()
    s.persistent(True)
    s.update(float(rank + 1))
    s.sync()
    assert s._is_synced and float(s.x) == 3.0 and float(s.state_dict()["x"]) == 3.0
    with pytest.raises(TorchMetricsUserError):
        s.sync()
    s.unsync()
    assert float(s.x) == rank + 1
    with pytest.raises(TorchMetricsUserError):
        s.unsync()
    with s.sync_context():
        assert float(s.x) == 3.0
    assert float(s.x) == rank + 1
    ns = DummySum(sync_on_compute=False)
    ns.update(float(rank + 1))
    assert float(ns.compute()) == rank + 1

    # ---- custom dist_sync_fn keeps its one-call-per-state contract under real DDP ------------------------------------------
    calls = []

    def fn(tensor, group=None):
        calls.append(tuple(tensor.shape))
        return gather_all_tensors(tensor, group)

    cu = DummyIntStates(n=2, dist_sync_fn=fn)
    cu.update([rank, 1])
    assert cu.compute()[0].tolist() == [1, 2] and len(calls) == 6

    # ---- dist_sync_on_step forward ------------------------------------------------------------------------------------------
    f = DummySum(dist_sync_on_step=True)
    assert float(f(float(rank + 1))) == 3.0
    assert float(f.x) == rank + 1

    # ---- collection under DDP -----------------------------------------------------------------------------------------------
    mc = MetricCollection({"a": DummySum(), "b": DummyMean()})
    mc.update(float(rank + 1))
    res = mc.compute()
    assert float(res["a"]) == 3.0 and float(res["b"]) == 1.5
    # ---- mAP: per-image list states, ragged image counts per rank, one packed exchange ---------------------------------------
    from metrics_b200.detection import MeanAveragePrecision
    from tests.helpers import synth_detection

    n_mine = 3 if rank == 0 else 2
    shards = [synth_detection(seed=40 + r, n_img=3 if r == 0 else 2, n_gt=3, n_det=5, n_cls=4, crowd_frac=0.3) for r in range(WORLD)]
    preds, target = shards[rank]
    target[0]["area"] = torch.tensor([10.0, 20.0, 30.0])  # explicit area on one image, default elsewhere
    mp_ = MeanAveragePrecision()
    mp_.update(preds[:2], target[:2])
    mp_.update(preds[2:], target[2:])
    local_boxes = [b.clone() for b in mp_.detection_box]
    mp_.sync()
    assert len(mp_.detection_box) == 5 == len(mp_.groundtruth_area) == len(mp_.detection_scores)
    order = [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2)]  # (rank, image): interleaved like the reference's per-image gathers
    for pos, (r, i) in enumerate(order):
        exp_p, exp_t = shards[r]
        assert torch.equal(mp_.detection_scores[pos], exp_p[i]["scores"])
        assert torch.equal(mp_.detection_labels[pos], exp_p[i]["labels"])
        assert torch.equal(mp_.groundtruth_labels[pos], exp_t[i]["labels"])
        assert torch.equal(mp_.groundtruth_crowds[pos], exp_t[i]["iscrowd"].to(torch.int64))
        assert tuple(mp_.detection_box[pos].shape) == (5, 4) and tuple(mp_.groundtruth_box[pos].shape) == (3, 4)
    assert mp_.groundtruth_area[0].tolist() == [10.0, 20.0, 30.0] and mp_.groundtruth_area[1].tolist() == [10.0, 20.0, 30.0]
    assert mp_.groundtruth_area[2].tolist() == [0.0, 0.0, 0.0]
    mp_.unsync()
    assert len(mp_.detection_box) == n_mine and all(torch.equal(a, b) for a, b in zip(mp_.detection_box, local_boxes))
    dist.barrier()


def _worker(rank: int, port: int, errq) -> None:
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.distributed.init_process_group("gloo", rank=rank, world_size=WORLD)
        _scenarios(rank)
        torch.distributed.destroy_process_group()
    except Exception:  # noqa: BLE001
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


@pytest.mark.timeout(240)
def test_sync_over_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, errq)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(200)
    errors = []
    while not errq.empty():
        errors.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            errors.append("worker hung")
    assert not errors and all(p.exitcode == 0 for p in procs), "\n".join(errors)
