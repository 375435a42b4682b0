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

    # ---- a cat state the descriptor cannot express (9 dims) next to integer sum states: it goes through the generic gather —
    # a decision carried IN the descriptor exchange, so all ranks take it together — and the integer buckets reduced before it
    # must not be reduced twice ---------------------------------------------------------------------------------------------
    class Mixed(DummyIntStates):
        def __init__(self, **kw):
            super().__init__(n=2, **kw)
            self.add_state("z", [], dist_reduce_fx="cat")

        def update(self, v):
            super().update(v)
            self.z.append(torch.full((1 + rank, 1, 1, 1, 1, 1, 1, 1, 2), float(rank)))

        def compute(self):
            from metrics_b200.utilities.data import dim_zero_cat

            return self.tp, dim_zero_cat(self.z)

    mx = Mixed()
    mx.update([1 + rank, 5])
    tp, z = mx.compute()
    assert tp.tolist() == [3, 10], tp  # 1 + 2 and 5 + 5: reduced exactly once
    assert tuple(z.shape) == (3, 1, 1, 1, 1, 1, 1, 1, 2) and z.flatten().tolist() == [0.0, 0.0, 1.0, 1.0, 1.0, 1.0]

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
    # ---- mAP with instance masks: the bit-packed per-image mask entries ride in the same packed exchange ---------------------
    import metrics_b200._native as native
    from tests.reference_runtime import cpu_kernels

    real_pack = native.mask_pack_entry
    native.mask_pack_entry = cpu_kernels.mask_pack_entry  # the kernel's stand-in: this scenario is about the exchange
    try:
        g = torch.Generator().manual_seed(70 + rank)
        sizes = [(9, 11), (33, 40), (5, 5)][: n_mine]
        seg_p = [dict(masks=torch.rand(2 + i, h, w, generator=g) > 0.5, scores=torch.rand(2 + i, generator=g),
                      labels=torch.randint(0, 3, (2 + i,), generator=g)) for i, (h, w) in enumerate(sizes)]
        seg_t = [dict(masks=torch.rand(i, h, w, generator=g) > 0.5, labels=torch.randint(0, 3, (i,), generator=g))
                 for i, (h, w) in enumerate(sizes)]  # image 0 has no ground truth at all
        ms = MeanAveragePrecision(iou_type="segm")
        ms.update(seg_p, seg_t)
        mine = [m.clone() for m in ms.detection_mask]
        assert [m[:3].tolist() for m in mine] == [[2 + i, h, w] for i, (h, w) in enumerate(sizes)]
        assert [int(m.numel()) for m in ms.groundtruth_mask] == [3 + i + i * ((h * w + 31) // 32) for i, (h, w) in enumerate(sizes)]
        for m, p_ in zip(mine, seg_p):  # areas follow the header
            assert m[3:3 + p_["masks"].shape[0]].tolist() == p_["masks"].reshape(p_["masks"].shape[0], -1).sum(1).tolist()
        ms.sync()
        assert len(ms.detection_mask) == 5 == len(ms.groundtruth_mask) == len(ms.detection_scores) and len(ms.detection_box) == 0
        heads = [m[:3].tolist() for m in ms.detection_mask]
        assert heads == [[2, 9, 11], [2, 9, 11], [3, 33, 40], [3, 33, 40], [4, 5, 5]]  # images interleaved rank by rank
        for pos, (r, i) in enumerate(order):
            if r == rank:
                assert torch.equal(ms.detection_mask[pos], mine[i]) and torch.equal(ms.detection_scores[pos], seg_p[i]["scores"])
            assert ms.detection_mask[pos].dtype == torch.int32 and ms.groundtruth_mask[pos][0] == i
        ms.unsync()
        assert len(ms.detection_mask) == n_mine and all(torch.equal(a, b) for a, b in zip(ms.detection_mask, mine))
    finally:
        native.mask_pack_entry = real_pack
    dist.barrier()
    _sharded_curve_choreography(rank)
    dist.barrier()


def _sharded_curve_choreography(rank: int) -> None:
    """Class-sharded AUROC/AP exchange (metrics_b200/parallel_curves.py) over gloo: the collectives, the ragged / empty
    rank handling and the collective cache decision run for real; the two kernels are replaced by numpy stand-ins built
    on the oracle (the real kernels are covered on 2 GPUs by tests/test_sharded_curves_gpu.py)."""
    import numpy as np

    import metrics_b200._native as native
    import metrics_b200.parallel_curves as pc
    from oracle import curves as oc

    def to_key(x: torch.Tensor) -> torch.Tensor:  # descending-order key of an f32 score, as the pack kernel emits it
        b = x.contiguous().view(torch.int32).numpy().view(np.uint32)
        ok = np.where(b >> 31, ~b, b | np.uint32(0x80000000))
        return torch.from_numpy((~ok).view(np.int32).copy())

    def from_key(k: np.ndarray) -> np.ndarray:
        ok = ~k.view(np.uint32)
        b = np.where(ok >> 31, ok & np.uint32(0x7FFFFFFF), ~ok)
        return b.view(np.float32)

    def pack_keys(preds, rows_out=None):
        n, c = preds.shape
        keys = torch.zeros((rows_out or c, n), dtype=torch.int32)
        keys[:c] = to_key(preds.float().T.contiguous())
        return keys

    def evaluate_keys(keys, target, first_class, nonneg=False):
        s, n = keys.shape
        au, ap, cnt = np.zeros(s, np.float32), np.zeros(s, np.float32), np.zeros((s, 3), np.int64)
        for j in range(s):
            score = from_key(keys[j].numpy())
            lab = (target.numpy() == first_class + j).astype(np.int64)
            au[j], ap[j] = oc.binary_auroc_exact(score, lab), oc.binary_average_precision_exact(score, lab)
            cnt[j] = [lab.sum(), n - lab.sum(), np.unique(score).size]
        return torch.from_numpy(au), torch.from_numpy(ap), torch.from_numpy(cnt)

    native.curve_pack_keys, native.curve_evaluate_keys = pack_keys, evaluate_keys
    group = torch.distributed.group.WORLD
    C = 5
    for case, ns in enumerate(([40, 27], [0, 33], [16, 16])):  # ragged, one empty rank, equal
        g = torch.Generator().manual_seed(100 * case)
        alld = []
        for r in range(WORLD):
            p = torch.softmax(torch.randn(ns[r], C, generator=g), 1)
            alld.append((p, torch.randint(0, C, (ns[r],), generator=g)))
        mine = alld[rank] if ns[rank] else (None, None)
        auroc, ap, counts = pc.ovr_curve_scalars_sharded(mine[0], mine[1], C, group, torch.device("cpu"))
        allp, allt = torch.cat([a[0] for a in alld]).numpy(), torch.cat([a[1] for a in alld]).numpy()
        np.testing.assert_allclose(auroc.numpy(), oc.multiclass_auroc_exact(allp, allt, C), rtol=1e-6)
        exp_ap = [oc.binary_average_precision_exact(allp[:, c], (allt == c).astype(np.int64)) for c in range(C)]
        np.testing.assert_allclose(ap.numpy(), exp_ap, rtol=1e-6)
        assert counts[:, 0].tolist() == [int((allt == c).sum()) for c in range(C)]
        # collective cache decision: a hit on one rank only must NOT skip the exchange (the other rank would dead-lock)
        again = pc.ovr_curve_scalars_sharded(mine[0], mine[1], C, group, torch.device("cpu"),
                                             cached=(auroc, ap, counts) if rank == 0 else None)
        assert torch.equal(again[0], auroc)
        both = pc.ovr_curve_scalars_sharded(mine[0], mine[1], C, group, torch.device("cpu"), cached=("sentinel",) * 3)
        assert both == ("sentinel",) * 3  # every rank had a hit: the memoised value is returned after one tiny all-reduce


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
