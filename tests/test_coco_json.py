"""CPU: the COCO json formats on either side of `MeanAveragePrecision` (reference detection/mean_ap.py:651-825, :867-958).
`tm_to_coco` is pinned by the reference's own formatter (tests/golden/coco_format.npz, made by make_golden.py coco_format);
`coco_to_tm` by round trip and by hand-written files in the format https://cocodataset.org/#format-data defines."""
import json

import pytest
import torch

from metrics_b200.detection import MeanAveragePrecision


def _images(g):
    n = int(g["n_images"])
    keys = ("d_box", "d_score", "d_label", "g_box", "g_label", "g_crowd", "g_area")
    return [{k: torch.from_numpy(g[f"img{i}/{k}"]) for k in keys} for i in range(n)]


def _updated_metric(images):
    metric = MeanAveragePrecision(box_format="xywh")
    metric.update(
        [{"boxes": im["d_box"], "scores": im["d_score"], "labels": im["d_label"]} for im in images],
        [{"boxes": im["g_box"], "labels": im["g_label"], "iscrowd": im["g_crowd"], "area": im["g_area"]} for im in images],
    )
    return metric


def test_tm_to_coco_matches_reference_formatter(golden_coco_format, tmp_path):
    images = _images(golden_coco_format)
    metric = _updated_metric(images)
    metric.tm_to_coco(str(tmp_path / "out"))
    target = json.loads((tmp_path / "out_target.json").read_text())
    preds = json.loads((tmp_path / "out_preds.json").read_text())
    assert target == json.loads(str(golden_coco_format["target_json"]))
    assert preds == json.loads(str(golden_coco_format["preds_json"]))
    assert [a["id"] for a in target["annotations"]] == list(range(1, len(target["annotations"]) + 1))
    assert (tmp_path / "out_preds.json").read_text().startswith("[\n    {\n        \"id\": 1,")  # indent=4 like the reference


def test_default_crowd_and_area_and_box_format_conversion(tmp_path):
    metric = MeanAveragePrecision(box_format="xyxy")
    metric.update(
        [{"boxes": torch.tensor([[258.0, 41.0, 606.0, 285.0]]), "scores": torch.tensor([0.536]), "labels": torch.tensor([0])}],
        [{"boxes": torch.tensor([[214.0, 41.0, 562.0, 285.0]]), "labels": torch.tensor([0])}],
    )
    metric.tm_to_coco(str(tmp_path / "x"))
    target = json.loads((tmp_path / "x_target.json").read_text())
    assert target["annotations"] == [{"id": 1, "image_id": 0, "area": 348.0 * 244.0, "category_id": 0, "iscrowd": 0,
                                      "bbox": [214.0, 41.0, 348.0, 244.0]}]
    assert target["images"] == [{"id": 0}] and target["categories"] == [{"id": 0, "name": "0"}]
    preds = json.loads((tmp_path / "x_preds.json").read_text())
    assert preds[0]["bbox"] == [258.0, 41.0, 348.0, 244.0] and preds[0]["score"] == pytest.approx(0.536)


def test_round_trip_through_files(golden_coco_format, tmp_path):
    images = _images(golden_coco_format)
    _updated_metric(images).tm_to_coco(str(tmp_path / "rt"))
    preds, target = MeanAveragePrecision.coco_to_tm(str(tmp_path / "rt_preds.json"), str(tmp_path / "rt_target.json"))
    with_gt = [im for im in images if im["g_label"].numel() > 0]  # images without ground truth leave no trace in the files
    assert len(preds) == len(target) == len(with_gt)
    for p, t, im in zip(preds, target, with_gt):
        assert t["labels"].dtype == torch.int32 and t["iscrowd"].dtype == torch.int32 and p["labels"].dtype == torch.int32
        assert torch.equal(t["boxes"], im["g_box"]) and torch.equal(t["labels"].long(), im["g_label"])
        assert torch.equal(t["iscrowd"].long(), im["g_crowd"].long())
        fallback = im["g_box"][:, 2].double() * im["g_box"][:, 3].double()
        expected = torch.where(im["g_area"].double() > 0, im["g_area"].double(), fallback).float()
        assert torch.equal(t["area"], expected)
        assert torch.equal(p["scores"], im["d_score"]) and torch.equal(p["labels"].long(), im["d_label"])
        if im["d_label"].numel():
            assert torch.equal(p["boxes"], im["d_box"])
        else:
            assert p["boxes"].numel() == 0
    again = MeanAveragePrecision(box_format="xywh")
    again.update(preds, target)  # the converted lists are valid `update` input
    assert len(again.detection_labels) == len(with_gt)


def test_coco_to_tm_reads_plain_coco_files(tmp_path):
    gt = {"images": [{"id": 7}, {"id": 9}, {"id": 11}], "categories": [{"id": 1}],
          "annotations": [{"id": 1, "image_id": 9, "bbox": [1, 2, 3, 4], "category_id": 1, "iscrowd": 0, "area": 12},
                          {"id": 2, "image_id": 7, "bbox": [0, 0, 5, 5], "category_id": 1, "iscrowd": 1, "area": 25},
                          {"id": 3, "image_id": 9, "bbox": [2, 2, 2, 2], "category_id": 1, "iscrowd": 0, "area": 4}]}
    dt = [{"image_id": 7, "category_id": 1, "bbox": [0, 0, 4, 5], "score": 0.9},
          {"image_id": 11, "category_id": 1, "bbox": [0, 0, 1, 1], "score": 0.5}]  # image 11 has no ground truth: dropped
    (tmp_path / "gt.json").write_text(json.dumps(gt))
    (tmp_path / "dt.json").write_text(json.dumps(dt))
    preds, target = MeanAveragePrecision.coco_to_tm(str(tmp_path / "dt.json"), str(tmp_path / "gt.json"), iou_type="bbox")
    assert [t["boxes"].tolist() for t in target] == [[[1, 2, 3, 4], [2, 2, 2, 2]], [[0, 0, 5, 5]]]  # first-appearance order
    assert [t["iscrowd"].tolist() for t in target] == [[0, 0], [1]] and target[0]["area"].tolist() == [12.0, 4.0]
    assert preds[0]["scores"].numel() == 0 and preds[0]["boxes"].numel() == 0
    assert preds[1]["boxes"].tolist() == [[0, 0, 4, 5]] and preds[1]["scores"].tolist() == pytest.approx([0.9])
    (tmp_path / "bad.json").write_text(json.dumps([{"image_id": 99, "category_id": 1, "bbox": [0, 0, 1, 1], "score": 0.1}]))
    with pytest.raises(ValueError, match="do not correspond"):
        MeanAveragePrecision.coco_to_tm(str(tmp_path / "bad.json"), str(tmp_path / "gt.json"))
    with pytest.raises(KeyError, match="segmentation"):  # these annotations carry boxes only (segm json: tests/test_rle.py)
        MeanAveragePrecision.coco_to_tm(str(tmp_path / "dt.json"), str(tmp_path / "gt.json"), iou_type="segm")
    with pytest.raises(ValueError, match="iou_type"):
        MeanAveragePrecision.coco_to_tm(str(tmp_path / "dt.json"), str(tmp_path / "gt.json"), iou_type="boxes")


def test_tm_to_coco_rejects_malformed_states(tmp_path):
    metric = MeanAveragePrecision(box_format="xywh")
    metric.update([{"boxes": torch.rand(1, 4), "scores": torch.rand(1), "labels": torch.tensor([1])}],
                  [{"boxes": torch.rand(1, 4), "labels": torch.tensor([1])}])
    metric.detection_labels[0] = torch.tensor([1.5])
    with pytest.raises(ValueError, match="Invalid input class of sample 0, element 0"):
        metric.tm_to_coco(str(tmp_path / "bad"))
    metric.detection_labels[0] = torch.tensor([1])
    metric.detection_scores[0] = torch.tensor([1])
    with pytest.raises(ValueError, match="Invalid input score of sample 0, element 0"):
        metric.tm_to_coco(str(tmp_path / "bad"))
