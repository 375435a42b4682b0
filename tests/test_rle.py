"""CPU: the COCO run-length codec of metrics_b200/detection/rle.py (json formats of segm mAP): hand-computed cases of the
published format, round trips, the bit-row decoder of the metric's mask state."""
import numpy as np
import pytest

from metrics_b200.detection import rle


def test_counts_are_column_major_and_start_with_zeros():
    m = np.array([[0, 1, 1],
                  [0, 1, 0]], bool)  # columns: [0,0] [1,1] [1,0]
    assert rle.mask_to_counts(m) == [2, 3, 1]
    assert rle.mask_to_counts(~m) == [0, 2, 3, 1]
    assert rle.mask_to_counts(np.zeros((2, 2), bool)) == [4]
    assert rle.mask_to_counts(np.ones((2, 2), bool)) == [0, 4]
    np.testing.assert_array_equal(rle.counts_to_mask([2, 3, 1], 2, 3), m)
    with pytest.raises(ValueError):
        rle.counts_to_mask([2, 3], 2, 3)


def test_string_code_by_hand():
    """maskApi.c:rleToString: 5-bit groups, bit 5 = continuation, sign in bit 4 of the last group, +48; from the fourth count
    on the DIFFERENCE to the count two places back is coded."""
    assert rle.counts_to_string([0]) == "0"     # group 0, nothing follows -> chr(0 + 48)
    assert rle.counts_to_string([6]) == "6"     # chr(6 + 48)
    assert rle.counts_to_string([15]) == "?"    # chr(15 + 48); bit 4 clear and the rest is 0: done
    assert rle.counts_to_string([16]) == "`0"   # group 16 has bit 4 set and the rest (0) is not -1: continue -> chr(16 + 32 + 48), then 0
    assert rle.counts_to_string([31]) == "o0"   # chr(31 + 32 + 48), then 0
    assert rle.counts_to_string([32]) == "P1"   # group 0 with continuation -> chr(32 + 48), then 1
    # 4th count 1 is coded as 1 - 10 = -9: group (-9 & 31) = 23, rest -1 and bit 4 set: done -> chr(23 + 48) = "G";
    # 5th count 30 as 30 - 2 = 28: group 28 (bit 4 set, rest 0): continue -> chr(28 + 32 + 48) = "l", then "0"
    assert rle.counts_to_string([3, 10, 2, 1, 30]) == "3:2Gl0"
    assert rle.string_to_counts("3:2Gl0") == [3, 10, 2, 1, 30]


def test_string_code_round_trips():
    g = np.random.default_rng(0)
    for _ in range(200):
        n = int(g.integers(1, 40))
        counts = [int(x) for x in g.integers(0, 3000, n)]
        counts[0] = int(g.integers(0, 2)) * counts[0]
        assert rle.string_to_counts(rle.counts_to_string(counts)) == counts
        assert rle.string_to_counts(rle.counts_to_string(counts).encode()) == counts
    big = [0, 307200 - 5, 5]
    assert rle.string_to_counts(rle.counts_to_string(big)) == big
    code = rle.counts_to_string([3, 10, 2, 1, 30])     # 4th and 5th counts are coded as 1 - 10 = -9 and 30 - 2 = 28
    assert all(48 <= ord(c) < 48 + 64 for c in code)


def test_masks_round_trip_through_both_codes():
    g = np.random.default_rng(1)
    for h, w in ((1, 1), (5, 7), (33, 31), (64, 100)):
        m = g.random((h, w)) > 0.6
        counts = rle.mask_to_counts(m)
        assert sum(counts) == h * w
        np.testing.assert_array_equal(rle.counts_to_mask(counts, h, w), m)
        for seg in ({"size": [h, w], "counts": counts}, {"size": [h, w], "counts": rle.counts_to_string(counts)}):
            got = rle.segmentation_to_mask(seg)
            assert got.dtype == np.uint8
            np.testing.assert_array_equal(got.astype(bool), m)
    with pytest.raises(NotImplementedError):
        rle.segmentation_to_mask([[1.0, 1.0, 5.0, 1.0, 5.0, 5.0]], 10, 10)


def test_entry_decoder_matches_the_pack_layout():
    import torch

    from tests.reference_runtime import cpu_kernels

    g = torch.Generator().manual_seed(2)
    for h, w in ((3, 5), (32, 32), (17, 40)):
        m = torch.rand(4, h, w, generator=g) > 0.5
        entry = cpu_kernels.mask_pack_entry(m).numpy()
        np.testing.assert_array_equal(rle.entry_to_masks(entry), m.numpy())
    assert rle.entry_to_masks(np.array([0, 8, 9], np.int32)).shape == (0, 8, 9)


def test_tm_to_coco_and_back_on_the_host(tmp_path, monkeypatch):
    """The json layer around the mask states with the pack kernel's stand-in: `update` -> `tm_to_coco` -> files ->
    `coco_to_tm` gives the masks, labels, scores, crowds back (the evaluation of the result is a GPU test)."""
    import json

    import torch

    from metrics_b200 import _native
    from metrics_b200.detection import MeanAveragePrecision
    from tests.reference_runtime import cpu_kernels

    monkeypatch.setattr(_native, "mask_pack_entry", cpu_kernels.mask_pack_entry)
    g = torch.Generator().manual_seed(5)
    preds = [dict(masks=torch.rand(3, 9, 14, generator=g) > 0.5, scores=torch.rand(3, generator=g), labels=torch.tensor([1, 0, 1])),
             dict(masks=torch.zeros((0, 6, 6), dtype=torch.bool), scores=torch.zeros(0), labels=torch.zeros(0, dtype=torch.long))]
    target = [dict(masks=torch.rand(2, 9, 14, generator=g) > 0.5, labels=torch.tensor([1, 1]), iscrowd=torch.tensor([0, 1])),
              dict(masks=torch.rand(1, 6, 6, generator=g) > 0.5, labels=torch.tensor([0]), area=torch.tensor([12.5]))]
    m = MeanAveragePrecision(iou_type="segm")
    m.update(preds, target)
    name = str(tmp_path / "io")
    m.tm_to_coco(name)
    gt_file = json.load(open(name + "_target.json"))
    assert [a["area"] for a in gt_file["annotations"]] == [int(target[0]["masks"][0].sum()), int(target[0]["masks"][1].sum()), 12.5]
    assert [a["iscrowd"] for a in gt_file["annotations"]] == [0, 1, 0] and "bbox" not in gt_file["annotations"][0]
    assert gt_file["images"] == [{"id": 0, "height": 9, "width": 14}, {"id": 1, "height": 6, "width": 6}]
    p2, t2 = MeanAveragePrecision.coco_to_tm(name + "_preds.json", name + "_target.json", iou_type="segm")
    assert len(p2) == 2 and p2[1]["masks"].numel() == 0 and "boxes" not in p2[0]
    assert torch.equal(p2[0]["masks"].bool(), preds[0]["masks"]) and torch.equal(t2[0]["masks"].bool(), target[0]["masks"])
    assert torch.equal(t2[1]["masks"].bool(), target[1]["masks"]) and t2[0]["iscrowd"].tolist() == [0, 1]
    assert torch.allclose(p2[0]["scores"], preds[0]["scores"]) and p2[0]["labels"].tolist() == [1, 0, 1]
    m2 = MeanAveragePrecision(iou_type="segm")
    m2.update(p2, t2)  # the empty 1-D `masks` of the image without detections is accepted
    assert all(torch.equal(a, b) for a, b in zip(m.detection_mask[:1] + m.groundtruth_mask, m2.detection_mask[:1] + m2.groundtruth_mask))
    assert m2.detection_mask[1].tolist() == [0, 0, 0]
