"""GPU: detection.MeanAveragePrecision kernels (through the C-ABI) vs the fp64 oracle (oracle/coco_map.py).

The raw `precision [T,R,K,A,M]` / `recall` / `scores` tensors must agree with the oracle to 1e-12 (both are fp64 and
follow the same operation order); the 12 summary statistics (float32) to 1e-6 relative.
"""
import warnings

import numpy as np
import pytest
import torch

from oracle.coco_map import coco_evaluate
from tests.helpers import LEGACY_MAP_CASES, det_to_numpy, synth_detection

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
STATS = ["map", "map_50", "map_75", "map_small", "map_medium", "map_large", "mar_small", "mar_medium", "mar_large"]


def _to_dev(items):
    return [{k: v.to(DEV) for k, v in d.items()} for d in items]


def _run(preds, target, batch=None, **kw):
    from metrics_b200.detection import MeanAveragePrecision

    m = MeanAveragePrecision(extended_summary=True, class_metrics=True, **kw).to(DEV)
    m.warn_on_many_detections = False
    batch = batch or len(preds)
    for i in range(0, len(preds), batch):
        m.update(_to_dev(preds[i:i + batch]), _to_dev(target[i:i + batch]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return m.compute()


def _check(res, ref, max_dets=(1, 10, 100), micro=False):
    np.testing.assert_allclose(res["precision"].cpu().numpy(), ref["precision"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(res["recall"].cpu().numpy(), ref["recall"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(res["scores"].cpu().numpy(), ref["scores"], rtol=0, atol=1e-12)
    for k in STATS + [f"mar_{d}" for d in max_dets]:
        assert res[k].dtype == torch.float32 and res[k].numel() == 1  # 0-d after `_squeeze_if_scalar`, like the reference
        np.testing.assert_allclose(float(res[k]), ref[k], rtol=1e-6, atol=1e-7, err_msg=k)
    assert res["classes"].cpu().reshape(-1).tolist() == ref["classes"].tolist() and res["classes"].dtype == torch.int32
    if not micro:
        np.testing.assert_allclose(res["map_per_class"].cpu().numpy().reshape(-1), ref["map_per_class_values"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(res[f"mar_{max_dets[-1]}_per_class"].cpu().numpy().reshape(-1),
                                   ref[f"mar_{max_dets[-1]}_per_class_values"], rtol=1e-6, atol=1e-7)


def test_docstring_example():
    from metrics_b200.detection import MeanAveragePrecision

    preds = [dict(boxes=torch.tensor([[258.0, 41.0, 606.0, 285.0]]), scores=torch.tensor([0.536]), labels=torch.tensor([0]))]
    target = [dict(boxes=torch.tensor([[214.0, 41.0, 562.0, 285.0]]), labels=torch.tensor([0]))]
    m = MeanAveragePrecision(iou_type="bbox").to(DEV)
    m.update(_to_dev(preds), _to_dev(target))
    r = m.compute()
    want = dict(map=0.6, map_50=1.0, map_75=1.0, map_small=-1.0, map_medium=-1.0, map_large=0.6, mar_1=0.6, mar_10=0.6,
                mar_100=0.6, mar_small=-1.0, mar_medium=-1.0, mar_large=0.6, map_per_class=-1.0, mar_100_per_class=-1.0)
    for k, v in want.items():
        assert float(r[k]) == pytest.approx(v, abs=1e-6), k
    assert r["classes"].reshape(-1).tolist() == [0]


@pytest.mark.parametrize("name", list(LEGACY_MAP_CASES))
def test_synthetic_vs_oracle(name):
    preds, target = synth_detection(**LEGACY_MAP_CASES[name])
    _check(_run(preds, target, batch=7), coco_evaluate(**det_to_numpy(preds, target)))


def test_crowds_given_areas_and_small_objects():
    preds, target = synth_detection(seed=11, n_img=40, n_gt=9, n_det=35, n_cls=6, crowd_frac=0.2, dup_scores=True)
    g = torch.Generator().manual_seed(3)
    for t in target:  # shrink some boxes so that small / medium ranges are populated, give explicit areas to some
        scale = torch.where(torch.rand(9, generator=g) < 0.5, 0.15, 1.0)
        wh = (t["boxes"][:, 2:] - t["boxes"][:, :2]) * scale[:, None]
        t["boxes"] = torch.cat([t["boxes"][:, :2], t["boxes"][:, :2] + wh], 1)
        t["area"] = torch.where(torch.rand(9, generator=g) < 0.3, wh[:, 0] * wh[:, 1] * 0.5, torch.zeros(9))
    for p, t in zip(preds, target):
        p["boxes"][:9] = t["boxes"] + torch.randn(9, 4, generator=g) * 2.0
        p["boxes"][:, 2:] = torch.maximum(p["boxes"][:, 2:], p["boxes"][:, :2] + 0.5)
    ref = coco_evaluate(**det_to_numpy(preds, target))
    assert float(ref["map_small"]) > -1 and float(ref["map_medium"]) > -1
    _check(_run(preds, target, batch=16), ref)


@pytest.mark.parametrize("fmt", ["xywh", "cxcywh"])
def test_box_formats_vs_oracle(fmt):
    preds, target = synth_detection(seed=12, n_img=10, n_gt=5, n_det=12, n_cls=3)
    for items in (preds, target):
        for d in items:
            b = d["boxes"]
            wh = b[:, 2:] - b[:, :2]
            d["boxes"] = torch.cat([b[:, :2], wh], 1) if fmt == "xywh" else torch.cat([b[:, :2] + wh / 2, wh], 1)
    _check(_run(preds, target, box_format=fmt), coco_evaluate(**det_to_numpy(preds, target), box_format=fmt))


def test_micro_average_custom_thresholds_and_max_dets():
    preds, target = synth_detection(seed=13, n_img=25, n_gt=7, n_det=150, n_cls=5)
    kw = dict(iou_thresholds=[0.3, 0.5, 0.75, 0.9], rec_thresholds=[0.0, 0.25, 0.5, 0.75, 1.0], max_detection_thresholds=[2, 20, 120])
    ref = coco_evaluate(**det_to_numpy(preds, target), **kw)
    _check(_run(preds, target, **kw), ref, max_dets=(2, 20, 120))
    ref = coco_evaluate(**det_to_numpy(preds, target), average="micro")
    res = _run(preds, target, average="micro")
    _check(res, ref, micro=True)
    # class_metrics under micro averaging re-evaluates with the true labels (detection/mean_ap.py:566-569)
    macro = coco_evaluate(**det_to_numpy(preds, target))
    np.testing.assert_allclose(res["map_per_class"].cpu().numpy().reshape(-1), macro["map_per_class_values"], rtol=1e-6, atol=1e-7)


def test_empty_sides_and_missing_thresholds():
    from metrics_b200.detection import MeanAveragePrecision

    box = torch.tensor([[214.15, 41.29, 562.41, 285.07]])
    empty = dict(boxes=torch.zeros(0, 4), scores=torch.zeros(0), labels=torch.zeros(0, dtype=torch.long))
    for preds, target, want in (
        ([empty], [dict(boxes=box, labels=torch.tensor([4]))], 0.0),
        ([dict(boxes=box, scores=torch.tensor([0.5]), labels=torch.tensor([4]))], [dict(boxes=torch.zeros(0, 4), labels=torch.zeros(0, dtype=torch.long))], -1.0),
        ([dict(boxes=torch.tensor([]), scores=torch.tensor([]), labels=torch.tensor([], dtype=torch.long))], [dict(boxes=box, labels=torch.tensor([4]))], 0.0),
    ):
        m = MeanAveragePrecision().to(DEV)
        m.update(_to_dev(preds), _to_dev(target))
        assert float(m.compute()["map"]) == want
    m = MeanAveragePrecision(iou_thresholds=[0.1, 0.2]).to(DEV)
    m.update(_to_dev([dict(boxes=box, scores=torch.tensor([0.5]), labels=torch.tensor([4]))]), _to_dev([dict(boxes=box, labels=torch.tensor([4]))]))
    r = m.compute()
    assert float(r["map_50"]) == -1.0 and float(r["map_75"]) == -1.0 and float(r["map"]) == 1.0
    fresh = MeanAveragePrecision().to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert float(fresh.compute()["map"]) == -1.0
    with pytest.raises(ValueError, match="Expected argument `preds` and `target` to have the same length"):
        fresh.update([], [dict(boxes=box, labels=torch.tensor([4]))])
    with pytest.raises(ValueError, match="Expected argument `class_metrics` to be a boolean"):
        MeanAveragePrecision(class_metrics=0)


def test_cfg4_shape_invariances_at_full_size():
    """BASELINE cfg4 shape (5000 images x 100 detections x 80 classes): too large for the Python oracle, so check
    size-independent properties: image-order permutation invariance (scores are distinct fp32 values) and agreement of
    a 150-image prefix with the oracle."""
    preds, target = synth_detection(seed=0, n_img=5000, n_gt=20, n_det=100, n_cls=80, crowd_frac=0.02)
    res = _run(preds, target, batch=100)
    perm = torch.randperm(5000, generator=torch.Generator().manual_seed(1)).tolist()
    res_p = _run([preds[i] for i in perm], [target[i] for i in perm], batch=100)
    for k in STATS + ["mar_1", "mar_10", "mar_100"]:
        np.testing.assert_allclose(res[k].cpu().numpy(), res_p[k].cpu().numpy(), rtol=1e-6, atol=1e-7, err_msg=k)
    assert 0.0 < float(res["map"]) < 1.0 and res["precision"].shape == (10, 101, 80, 4, 3)
    sub_p, sub_t = preds[:150], target[:150]
    _check(_run(sub_p, sub_t, batch=50), coco_evaluate(**det_to_numpy(sub_p, sub_t)))


def test_more_than_256_ground_truths_of_one_class_in_an_image():
    """Past 256 ground truths per image the matcher's "already matched" masks move from registers to shared memory
    (csrc/cocomap.cu kSmemMask): a crowded image of ONE class (300 and 700 boxes) must still follow the oracle."""
    preds, target = synth_detection(seed=5, n_img=3, n_gt=300, n_det=100, n_cls=1, crowd_frac=0.05, dup_scores=True)
    big_p, big_t = synth_detection(seed=6, n_img=1, n_gt=700, n_det=100, n_cls=2, crowd_frac=0.0)
    big_t[0]["labels"][:] = 0  # 700 ground truths of one class
    preds, target = preds + big_p, target + big_t
    _check(_run(preds, target, batch=2), coco_evaluate(**det_to_numpy(preds, target)))
