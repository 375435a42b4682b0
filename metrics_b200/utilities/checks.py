"""Shape checks shared by the functional front-ends (reference: utilities/checks.py:38-43)."""
from torch import Tensor


def _check_same_shape(preds: Tensor, target: Tensor) -> None:
    if preds.shape != target.shape:
        raise RuntimeError(
            f"Predictions and targets are expected to have the same shape, but got {preds.shape} and {target.shape}."
        )
