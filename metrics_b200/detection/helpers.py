"""Input checks for detection metrics (reference: detection/helpers.py:20-88)."""
from __future__ import annotations

from collections.abc import Sequence
from typing import Dict, Tuple, Union

from torch import Tensor


_ITEM_KEY = {"bbox": "boxes", "segm": "masks"}


def _input_validator(preds: Sequence[Dict[str, Tensor]], targets: Sequence[Dict[str, Tensor]],
                     iou_type: Union[str, Tuple[str, ...]] = "bbox", ignore_score: bool = False) -> None:
    """Type / key / length checks of the list-of-dict inputs (reference :20-81): per IoU type the geometry key is ``boxes``
    (bbox) or ``masks`` (segm); messages as in the reference."""
    kinds = (iou_type,) if isinstance(iou_type, str) else tuple(iou_type)
    if any(k not in _ITEM_KEY for k in kinds):
        raise Exception(f"IOU type {kinds} is not supported")
    geometry = [_ITEM_KEY[k] for k in kinds]
    if not isinstance(preds, Sequence):
        raise ValueError(f"Expected argument `preds` to be of type Sequence, but got {preds}")
    if not isinstance(targets, Sequence):
        raise ValueError(f"Expected argument `target` to be of type Sequence, but got {targets}")
    if len(preds) != len(targets):
        raise ValueError(
            f"Expected argument `preds` and `target` to have the same length, but got {len(preds)} and {len(targets)}"
        )
    pred_keys = geometry + ["labels"] + ([] if ignore_score else ["scores"])
    target_keys = geometry + ["labels"]
    for key in pred_keys:
        if any(key not in p for p in preds):
            raise ValueError(f"Expected all dicts in `preds` to contain the `{key}` key")
    for key in target_keys:
        if any(key not in t for t in targets):
            raise ValueError(f"Expected all dicts in `target` to contain the `{key}` key")
    for key in geometry + ([] if ignore_score else ["scores"]) + ["labels"]:
        if not all(isinstance(p[key], Tensor) for p in preds):
            raise ValueError(f"Expected all {key} in `preds` to be of type Tensor")
    for key in target_keys:
        if not all(isinstance(t[key], Tensor) for t in targets):
            raise ValueError(f"Expected all {key} in `target` to be of type Tensor")
    for i, item in enumerate(targets):
        for key in geometry:
            if item[key].size(0) != item["labels"].size(0):
                raise ValueError(
                    f"Input '{key}' and labels of sample {i} in targets have a"
                    f" different length (expected {item[key].size(0)} labels, got {item['labels'].size(0)})"
                )
    if ignore_score:
        return
    for i, item in enumerate(preds):
        for key in geometry:
            if not (item[key].size(0) == item["labels"].size(0) == item["scores"].size(0)):
                raise ValueError(
                    f"Input '{key}', labels and scores of sample {i} in predictions have a"
                    f" different length (expected {item[key].size(0)} labels and scores,"
                    f" got {item['labels'].size(0)} labels and {item['scores'].size(0)})"
                )


def _fix_empty_tensors(boxes: Tensor) -> Tensor:
    """A 1-D empty box tensor becomes ``[1, 0]`` (reference :84-88; keeps DDP gathers well-formed)."""
    if boxes.numel() == 0 and boxes.ndim == 1:
        return boxes.unsqueeze(0)
    return boxes
