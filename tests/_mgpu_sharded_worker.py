"""torchrun worker for tests/test_sharded_curves_gpu.py: class-sharded AUROC/AP over NCCL vs (a) the gather-everything sync
and (b) a single-GPU evaluation of the concatenated data.  Prints 'SHARDED_OK' on rank 0 when everything agrees."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rank_data(rank, n, c, seed):
    g = torch.Generator().manual_seed(1000 * seed + rank)
    lg = torch.randn(n, c, generator=g)
    lg = (lg * 4).round() / 4 if seed % 2 else lg  # ties across ranks on odd seeds
    return lg, torch.randint(0, c, (n,), generator=g)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.distributed.init_process_group("nccl", device_id=dev)
    from metrics_b200 import MetricCollection
    from metrics_b200.classification import MulticlassAUROC, MulticlassAveragePrecision

    warnings.simplefilter("ignore")
    cases = [  # (C, samples per rank, ignore_index, average)
        (5, [300 + 17 * r for r in range(world)], None, "macro"),
        (7, [0 if r == 0 else 129 for r in range(world)], None, "none"),  # rank 0 never updates
        (1000, [4096] * world, None, "macro"),
        (12, [257 * (r + 1) for r in range(world)], 3, "weighted"),
        (3, [64] * world, None, "none"),
    ]
    for seed, (c, ns, ignore, average) in enumerate(cases):
        kw = dict(num_classes=c, average=average, ignore_index=ignore, validate_args=False)
        mine = rank_data(rank, ns[rank], c, seed)

        def build():
            mc = MetricCollection([MulticlassAUROC(**kw), MulticlassAveragePrecision(**kw)]).to(dev)
            if ns[rank] > 0:
                half = ns[rank] // 2
                mc.update(mine[0][:half].to(dev), mine[1][:half].to(dev))
                mc.update(mine[0][half:].to(dev), mine[1][half:].to(dev))
            return mc

        os.environ["MB200_SHARDED_CURVES"] = "1"
        sharded = build().compute()
        os.environ["MB200_SHARDED_CURVES"] = "0"
        gathered = build().compute()
        os.environ["MB200_SHARDED_CURVES"] = "1"
        # single-GPU truth on the union, in rank order
        alld = [rank_data(r, ns[r], c, seed) for r in range(world)]
        lg, tg = torch.cat([a[0] for a in alld]).to(dev), torch.cat([a[1] for a in alld]).to(dev)
        one = MetricCollection([MulticlassAUROC(**kw, sync_on_compute=False), MulticlassAveragePrecision(**kw, sync_on_compute=False)]).to(dev)
        one.update(lg, tg)
        truth = one.compute()
        for k in truth:
            a, b, t = sharded[k].cpu(), gathered[k].cpu(), truth[k].cpu()
            assert torch.equal(a.isnan(), t.isnan()), (k, seed, a, t)
            # same keys, same integer scan, same fp64 reduction order per class: bit-identical to both baselines
            assert torch.equal(a.nan_to_num(7.0), t.nan_to_num(7.0)), (k, seed, a, t)
            assert torch.equal(b.nan_to_num(7.0), t.nan_to_num(7.0)), (k, seed, b, t)
    # ---- mAP under DDP: packed ragged exchange of the per-image list states, then the device evaluation ----------------------
    from metrics_b200.detection import MeanAveragePrecision
    from tests.helpers import synth_detection

    n_imgs = [6 + 2 * r for r in range(world)]
    shards = [synth_detection(seed=70 + r, n_img=n_imgs[r], n_gt=6, n_det=12, n_cls=5, crowd_frac=0.1, dup_scores=True)
              for r in range(world)]

    def to_dev(items):
        return [{k: v.to(dev) for k, v in d.items()} for d in items]

    m = MeanAveragePrecision(class_metrics=True).to(dev)
    m.update(to_dev(shards[rank][0]), to_dev(shards[rank][1]))
    got = m.compute()
    one = MeanAveragePrecision(class_metrics=True, sync_on_compute=False).to(dev)
    for i in range(max(n_imgs)):  # the interleaved image order of the synced state
        for r in range(world):
            if i < n_imgs[r]:
                one.update(to_dev(shards[r][0][i:i + 1]), to_dev(shards[r][1][i:i + 1]))
    truth = one.compute()
    for k in truth:
        assert torch.equal(got[k].cpu(), truth[k].cpu()), (k, got[k], truth[k])
    assert len(m.detection_box) == n_imgs[rank]  # unsynced again
    # the default path above is the evaluation SHARDED over ranks (own images matched, own classes accumulated); the
    # reference-shaped "gather every image everywhere" sync must give the very same numbers, and so must an empty rank
    os.environ["MB200_SHARDED_MAP"] = "0"
    m._computed = None
    gathered = m.compute()
    os.environ["MB200_SHARDED_MAP"] = "1"
    for k in truth:
        assert torch.equal(gathered[k].cpu(), truth[k].cpu()), (k, gathered[k], truth[k])
    lonely = MeanAveragePrecision(class_metrics=True).to(dev)
    if rank != 0:  # rank 0 holds no image at all
        lonely.update(to_dev(shards[rank][0]), to_dev(shards[rank][1]))
    got2 = lonely.compute()
    one2 = MeanAveragePrecision(class_metrics=True, sync_on_compute=False).to(dev)
    for i in range(max(n_imgs)):
        for r in range(1, world):
            if i < n_imgs[r]:
                one2.update(to_dev(shards[r][0][i:i + 1]), to_dev(shards[r][1][i:i + 1]))
    truth2 = one2.compute()
    for k in truth2:
        assert torch.equal(got2[k].cpu(), truth2[k].cpu()), (k, got2[k], truth2[k])
    # ---- the same three-way agreement with instance masks (("bbox", "segm")): sharded (no mask leaves its rank), gathered
    # (bit-packed masks in the packed exchange), one GPU fed the interleaved images -------------------------------------------
    from tests.test_map_segm_gpu import synth_masks

    seg = [synth_masks(seed=90 + r, n_img=4 + 3 * r, n_gt=4, n_det=9, n_cls=4, crowd_frac=0.2, dup_scores=True, with_boxes=True)
           for r in range(world)]
    n_seg = [4 + 3 * r for r in range(world)]
    ms = MeanAveragePrecision(iou_type=("bbox", "segm"), class_metrics=True).to(dev)
    ms.update(to_dev(seg[rank][0]), to_dev(seg[rank][1]))
    got_s = ms.compute()
    one_s = MeanAveragePrecision(iou_type=("bbox", "segm"), class_metrics=True, sync_on_compute=False).to(dev)
    for i in range(max(n_seg)):
        for r in range(world):
            if i < n_seg[r]:
                one_s.update(to_dev(seg[r][0][i:i + 1]), to_dev(seg[r][1][i:i + 1]))
    truth_s = one_s.compute()
    os.environ["MB200_SHARDED_MAP"] = "0"
    ms._computed = None
    gathered_s = ms.compute()
    os.environ["MB200_SHARDED_MAP"] = "1"
    assert "segm_map" in truth_s and "bbox_map_per_class" in truth_s
    for k in truth_s:
        assert torch.equal(got_s[k].cpu(), truth_s[k].cpu()), (k, got_s[k], truth_s[k])
        assert torch.equal(gathered_s[k].cpu(), truth_s[k].cpu()), (k, gathered_s[k], truth_s[k])
    assert len(ms.detection_mask) == n_seg[rank]
    torch.distributed.barrier()
    if rank == 0:
        print("SHARDED_OK")
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
