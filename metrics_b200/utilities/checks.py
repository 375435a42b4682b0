"""Shape checks shared by the functional front-ends (reference: utilities/checks.py:38-43)."""
from torch import Tensor


def _check_same_shape(preds: Tensor, target: Tensor) -> None:
    if preds.shape != target.shape:
        raise RuntimeError(
            f"Predictions and targets are expected to have the same shape, but got {preds.shape} and {target.shape}."
        )


def _allclose_recursive(res1, res2, atol: float = 1e-6) -> bool:
    """Structural comparison of two metric results (tensors within ``atol``, containers element-wise) — reference
    utilities/checks.py:620-631."""
    from collections.abc import Mapping, Sequence

    import torch

    if isinstance(res1, Tensor):
        return bool(torch.allclose(res1, res2, atol=atol))
    if isinstance(res1, str):
        return res1 == res2
    if isinstance(res1, Mapping):
        return all(_allclose_recursive(res1[key], res2[key], atol) for key in res1)
    if isinstance(res1, Sequence):
        return all(_allclose_recursive(a, b, atol) for a, b in zip(res1, res2))
    return res1 == res2
