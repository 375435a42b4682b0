"""GPU: MeanAveragePrecision with instance masks (iou_type "segm" and ("bbox", "segm")) — K12 (csrc/maskiou.cu) and the mask
mode of the matching kernel, through the metric class and the C-ABI, against the fp64 oracle (oracle/coco_map.py::mask_iou +
COCOeval restated) and the reference's docstring known answer (detection/mean_ap.py:285-340).

pycocotools is not installed in this project's containers, so — as for boxes — the oracle is pinned only by the reference's
own known answers; what these tests prove is that the kernels compute what the oracle states.
"""
import warnings

import numpy as np
import pytest
import torch

from oracle.coco_map import coco_evaluate, mask_iou

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
STATS = ["map", "map_50", "map_75", "map_small", "map_medium", "map_large", "mar_1", "mar_10", "mar_100", "mar_small",
         "mar_medium", "mar_large"]


def synth_masks(seed, n_img, n_gt, n_det, n_cls, sizes=((37, 53), (64, 64), (20, 131)), crowd_frac=0.0, empty_every=0,
                dup_scores=False, with_boxes=False):
    """Per image: ground-truth blobs (filled rectangles with a random hole) and detections that are jittered copies of them
    or noise; image sizes vary (and are not multiples of 32 pixels)."""
    g = torch.Generator().manual_seed(seed)
    preds, target = [], []

    def blob(h, w):
        y0, x0 = int(torch.randint(0, h - 4, (1,), generator=g)), int(torch.randint(0, w - 4, (1,), generator=g))
        y1 = int(torch.randint(y0 + 2, h + 1, (1,), generator=g))
        x1 = int(torch.randint(x0 + 2, w + 1, (1,), generator=g))
        m = torch.zeros(h, w, dtype=torch.bool)
        m[y0:y1, x0:x1] = True
        m &= torch.rand(h, w, generator=g) > 0.1
        return m, (x0, y0, x1, y1)

    for i in range(n_img):
        h, w = sizes[i % len(sizes)]
        ng = 0 if (empty_every and i % empty_every == 1) else int(torch.randint(1, n_gt + 1, (1,), generator=g))
        nd = 0 if (empty_every and i % empty_every == 2) else int(torch.randint(1, n_det + 1, (1,), generator=g))
        gm, gb = zip(*[blob(h, w) for _ in range(ng)]) if ng else ((), ())
        glab = torch.randint(0, n_cls, (ng,), generator=g)
        dm, db, dlab = [], [], []
        for k in range(nd):
            if ng and torch.rand(1, generator=g) < 0.7:
                j = int(torch.randint(0, ng, (1,), generator=g))
                shift = int(torch.randint(-3, 4, (1,), generator=g))
                dm.append(torch.roll(gm[j], shifts=shift, dims=1) & (torch.rand(h, w, generator=g) > 0.15))
                x0, y0, x1, y1 = gb[j]
                db.append((x0 + shift, y0, x1 + shift, y1))
                dlab.append(int(glab[j]) if torch.rand(1, generator=g) < 0.8 else int(torch.randint(0, n_cls, (1,), generator=g)))
            else:
                m, b = blob(h, w)
                dm.append(m), db.append(b), dlab.append(int(torch.randint(0, n_cls, (1,), generator=g)))
        scores = torch.rand(nd, generator=g)
        if dup_scores:
            scores = (scores * 8).floor() / 8
        p = dict(masks=torch.stack(dm) if nd else torch.zeros((0, h, w), dtype=torch.bool), scores=scores,
                 labels=torch.tensor(dlab, dtype=torch.long))
        t = dict(masks=torch.stack(gm) if ng else torch.zeros((0, h, w), dtype=torch.bool), labels=glab)
        if crowd_frac:
            t["iscrowd"] = (torch.rand(ng, generator=g) < crowd_frac).long()
        if with_boxes:
            p["boxes"] = torch.tensor(db, dtype=torch.float32).reshape(-1, 4)
            t["boxes"] = torch.tensor(gb, dtype=torch.float32).reshape(-1, 4)
        preds.append(p), target.append(t)
    return preds, target


def _to_dev(items):
    return [{k: v.to(DEV) for k, v in d.items()} for d in items]


def _np(preds, target, boxes):
    kw = dict(det_boxes=[p["boxes"].numpy() for p in preds] if boxes else None,
              gt_boxes=[t["boxes"].numpy() for t in target] if boxes else None,
              det_scores=[p["scores"].numpy() for p in preds], det_labels=[p["labels"].numpy() for p in preds],
              gt_labels=[t["labels"].numpy() for t in target], det_masks=[p["masks"].numpy() for p in preds],
              gt_masks=[t["masks"].numpy() for t in target])
    if any("iscrowd" in t for t in target):
        kw["gt_crowds"] = [t.get("iscrowd", torch.zeros_like(t["labels"])).numpy() for t in target]
    if any("area" in t for t in target):
        kw["gt_areas"] = [t.get("area", torch.zeros_like(t["labels"])).numpy() for t in target]
    return kw


def _run(preds, target, batch=5, **kw):
    from metrics_b200.detection import MeanAveragePrecision

    m = MeanAveragePrecision(extended_summary=True, class_metrics=True, **kw).to(DEV)
    m.warn_on_many_detections = False
    for i in range(0, len(preds), batch):
        m.update(_to_dev(preds[i:i + batch]), _to_dev(target[i:i + batch]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return m.compute()


def _check(res, ref, prefix="", micro=False):
    for name in ("precision", "recall", "scores"):
        np.testing.assert_allclose(res[prefix + name].cpu().numpy(), ref[name], rtol=0, atol=1e-12, err_msg=prefix + name)
    for k in STATS:
        np.testing.assert_allclose(float(res[prefix + k]), ref[k], rtol=1e-6, atol=1e-7, err_msg=prefix + k)
    assert res["classes"].cpu().reshape(-1).tolist() == ref["classes"].tolist()
    if not micro:
        np.testing.assert_allclose(res[prefix + "map_per_class"].cpu().numpy().reshape(-1), ref["map_per_class_values"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(res[prefix + "mar_100_per_class"].cpu().numpy().reshape(-1), ref["mar_100_per_class_values"],
                                   rtol=1e-6, atol=1e-7)


def test_reference_docstring_example():
    """detection/mean_ap.py:285-340: IoU 3/5 matches the thresholds 0.5 and 0.55 only -> map 0.2."""
    from metrics_b200.detection import MeanAveragePrecision

    mask_pred = [[0, 0, 0, 0, 0], [0, 0, 1, 1, 0], [0, 0, 1, 1, 0], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]]
    mask_tgt = [[0, 0, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 1, 1, 0], [0, 0, 1, 0, 0], [0, 0, 0, 0, 0]]
    preds = [dict(masks=torch.tensor([mask_pred], dtype=torch.bool), scores=torch.tensor([0.536]), labels=torch.tensor([0]))]
    target = [dict(masks=torch.tensor([mask_tgt], dtype=torch.bool), labels=torch.tensor([0]))]
    m = MeanAveragePrecision(iou_type="segm").to(DEV)
    m.update(_to_dev(preds), _to_dev(target))
    r = m.compute()
    want = dict(map=0.2, map_50=1.0, map_75=0.0, map_large=-1.0, map_medium=-1.0, map_per_class=-1.0, map_small=0.2, mar_1=0.2,
                mar_10=0.2, mar_100=0.2, mar_100_per_class=-1.0, mar_large=-1.0, mar_medium=-1.0, mar_small=0.2)
    assert set(r) == set(want) | {"classes"}
    for k, v in want.items():
        assert float(r[k]) == pytest.approx(v, abs=1e-6), k
    assert r["classes"].reshape(-1).tolist() == [0] and r["classes"].dtype == torch.int32


def test_pack_and_pair_kernels_vs_numpy():
    """`mb200_mask_pack_bits` (bit order, tail pixels, areas) and `mb200_mask_pair_intersections` against numpy."""
    from metrics_b200 import _native

    g = torch.Generator().manual_seed(0)
    for h, w in ((1, 1), (5, 7), (32, 32), (33, 31), (64, 100)):
        m = torch.rand(6, h, w, generator=g) > 0.4
        words, area = _native.mask_pack_bits(m.to(DEV))
        assert area.cpu().tolist() == m.reshape(6, -1).sum(1).tolist()
        flat = m.reshape(6, -1).numpy()
        nw = (h * w + 31) // 32
        pad = np.zeros((6, nw * 32), bool)
        pad[:, : h * w] = flat
        want = (pad.reshape(6, nw, 32) * (1 << np.arange(32, dtype=np.uint64))).sum(2).astype(np.uint32)
        np.testing.assert_array_equal(words.cpu().numpy().view(np.uint32), want)
        entry = _native.mask_pack_entry(m.to(DEV)).cpu()  # the metric's state entry: [n, H, W, areas, bit rows] in one call
        assert entry.dtype == torch.int32 and entry[:3].tolist() == [6, h, w] and entry[3:9].tolist() == area.cpu().tolist()
        np.testing.assert_array_equal(entry[9:].numpy().view(np.uint32).reshape(6, nw), want)
        assert _native.mask_pack_entry(m[:0].to(DEV)).cpu().tolist() == [0, h, w]
        as_bytes = (m.to(torch.uint8) * 7).to(DEV)  # any non-zero byte counts as set
        assert torch.equal(_native.mask_pack_entry(as_bytes).cpu(), entry)
        # two "images": masks 0-2 vs 3-5 (all pairs) and 3-5 vs 0-2, labels filter half of the pairs
        off = (torch.arange(6, dtype=torch.int64) * nw).to(DEV)
        lab = torch.tensor([0, 1, 0, 0, 0, 1], device=DEV)
        det_off = torch.tensor([0, 3, 6], dtype=torch.int32, device=DEV)
        perm = torch.tensor([3, 4, 5, 0, 1, 2], device=DEV)
        inter = _native.mask_pair_intersections(words.reshape(-1), off, words.reshape(-1), off[perm], det_off, det_off,
                                                torch.tensor([nw, nw], dtype=torch.int32, device=DEV), lab, lab[perm], False,
                                                torch.tensor([0, 9], device=DEV), 18, 9).cpu().numpy()
        for img in range(2):
            for d in range(3):
                for k in range(3):
                    di, gi = img * 3 + d, int(perm[img * 3 + k])
                    want_i = float((flat[di] & flat[gi]).sum()) if int(lab[di]) == int(lab[gi]) else 0.0
                    assert inter[img * 9 + d * 3 + k] == want_i


@pytest.mark.parametrize("case", [
    dict(seed=1, n_img=9, n_gt=4, n_det=8, n_cls=3),
    dict(seed=2, n_img=14, n_gt=6, n_det=12, n_cls=4, crowd_frac=0.3, dup_scores=True),
    dict(seed=3, n_img=10, n_gt=5, n_det=9, n_cls=2, empty_every=3),
    dict(seed=4, n_img=6, n_gt=8, n_det=20, n_cls=1, sizes=((150, 210), (97, 33))),
])
def test_segm_vs_oracle(case):
    preds, target = synth_masks(**case)
    _check(_run(preds, target, iou_type="segm"), coco_evaluate(**_np(preds, target, False), iou_type="segm"))


def test_segm_micro_and_given_areas():
    preds, target = synth_masks(seed=5, n_img=12, n_gt=5, n_det=10, n_cls=4, crowd_frac=0.2)
    g = torch.Generator().manual_seed(9)
    for t in target:  # explicit areas for some ground truths (some of them push the annotation into another size range)
        n = t["labels"].numel()
        t["area"] = torch.where(torch.rand(n, generator=g) < 0.4, torch.rand(n, generator=g) * 3000, torch.zeros(n))
    ref = coco_evaluate(**_np(preds, target, False), iou_type="segm", average="micro")
    res = _run(preds, target, iou_type="segm", average="micro")
    _check(res, ref, micro=True)
    # class metrics under micro averaging come from a macro re-evaluation (reference :562-585)
    macro = coco_evaluate(**_np(preds, target, False), iou_type="segm")
    np.testing.assert_allclose(res["map_per_class"].cpu().numpy().reshape(-1), macro["map_per_class_values"], rtol=1e-6, atol=1e-7)


def test_bbox_and_segm_together():
    """("bbox", "segm"): prefixed keys; the box evaluation uses the MASK area of ground truths without `area`
    (detection/mean_ap.py:920-925) and the boxes' own areas for the detections."""
    preds, target = synth_masks(seed=6, n_img=12, n_gt=5, n_det=10, n_cls=3, crowd_frac=0.2, with_boxes=True)
    res = _run(preds, target, iou_type=("bbox", "segm"))
    kw = _np(preds, target, True)
    _check(res, coco_evaluate(**kw, iou_type="segm"), prefix="segm_")
    _check(res, coco_evaluate(**kw, iou_type="bbox"), prefix="bbox_")
    assert "map" not in res and "bbox_map" in res and "segm_ious" in res


def test_extended_summary_ious_are_mask_ious():
    preds, target = synth_masks(seed=7, n_img=5, n_gt=4, n_det=6, n_cls=2, crowd_frac=0.3)
    res = _run(preds, target, iou_type="segm")
    classes = res["classes"].reshape(-1).tolist()
    for i, (p, t) in enumerate(zip(preds, target)):
        for c in classes:
            dm, gm = p["labels"] == c, t["labels"] == c
            got = res["ious"][(i, c)]
            if int(dm.sum()) == 0 or int(gm.sum()) == 0:
                assert len(got) == 0
                continue
            order = np.argsort(-p["scores"][dm].numpy(), kind="mergesort")
            want = mask_iou(p["masks"][dm].numpy()[order], t["masks"][gm].numpy(), t["iscrowd"][gm].numpy())
            np.testing.assert_allclose(got.cpu().numpy(), want.astype(np.float32), rtol=1e-6, atol=0)


def test_input_checks():
    from metrics_b200.detection import MeanAveragePrecision

    m = MeanAveragePrecision(iou_type="segm").to(DEV)
    ok = dict(masks=torch.zeros((1, 4, 4), dtype=torch.bool, device=DEV), labels=torch.zeros(1, dtype=torch.long, device=DEV))
    with pytest.raises(ValueError, match="Expected all dicts in `preds` to contain the `masks` key"):
        m.update([dict(scores=torch.zeros(1), labels=torch.zeros(1, dtype=torch.long))], [ok])
    with pytest.raises(ValueError, match="Input 'masks' and labels of sample 0 in targets"):
        m.update([dict(ok, scores=torch.zeros(1, device=DEV))], [dict(ok, labels=torch.zeros(2, dtype=torch.long, device=DEV))])
    m.update([dict(ok, scores=torch.zeros(1, device=DEV))], [dict(masks=torch.zeros((1, 5, 4), dtype=torch.bool, device=DEV), labels=ok["labels"])])
    with pytest.raises(ValueError, match="same height and width"):
        m.compute()


@pytest.mark.parametrize("iou_type", ["segm", ("bbox", "segm")])
def test_coco_json_round_trip(tmp_path, iou_type):
    """`tm_to_coco` writes the cached masks as compressed run-length codes, `coco_to_tm` reads them back: a second metric fed
    from the files computes the very same result; the files follow the reference's layout (:867-958)."""
    import json

    from metrics_b200.detection import MeanAveragePrecision
    from metrics_b200.detection.rle import segmentation_to_mask

    preds, target = synth_masks(seed=8, n_img=6, n_gt=3, n_det=5, n_cls=3, crowd_frac=0.3, with_boxes=True)
    h2, w2 = preds[2]["masks"].shape[1:]  # one image without any detection (images without ground truth are not in the files)
    preds[2] = dict(masks=torch.zeros((0, h2, w2), dtype=torch.bool), scores=torch.zeros(0), labels=torch.zeros(0, dtype=torch.long),
                    boxes=torch.zeros((0, 4)))
    m = MeanAveragePrecision(iou_type=iou_type, box_format="xywh").to(DEV)
    m.update(_to_dev(preds), _to_dev(target))
    want = m.compute()
    name = str(tmp_path / "segm_io")
    m.tm_to_coco(name)
    gt_file, dt_file = json.load(open(name + "_target.json")), json.load(open(name + "_preds.json"))
    ann = gt_file["annotations"][0]
    assert set(ann) >= {"id", "image_id", "area", "category_id", "iscrowd", "segmentation"} and ann["id"] == 1
    assert ("bbox" in ann) == (iou_type != "segm") and ("area_segm" in ann) == (iou_type != "segm")
    assert isinstance(ann["segmentation"]["counts"], str) and ann["segmentation"]["size"] == list(target[0]["masks"].shape[1:])
    first = 0
    np.testing.assert_array_equal(segmentation_to_mask(ann["segmentation"]).astype(bool), target[first]["masks"][0].numpy())
    assert ann["area"] == int(target[first]["masks"][0].sum())  # no `area` given: the mask area (reference :923-925)
    assert gt_file["images"][first]["height"] == target[first]["masks"].shape[1] and "score" in dt_file[0]
    p2, t2 = MeanAveragePrecision.coco_to_tm(name + "_preds.json", name + "_target.json", iou_type=iou_type)
    assert p2[0]["masks"].dtype == torch.uint8 and len(p2) == 6 and p2[2]["masks"].numel() == 0
    m2 = MeanAveragePrecision(iou_type=iou_type, box_format="xywh").to(DEV)
    m2.update(_to_dev(p2), _to_dev(t2))
    got = m2.compute()
    for k in want:
        np.testing.assert_allclose(got[k].cpu().numpy(), want[k].cpu().numpy(), rtol=1e-6, atol=1e-7, err_msg=k)
