"""State-reduction helpers of the runtime (reference: utilities/data.py:29-61, 150-176).

The arithmetic helpers that the reference keeps here (`_bincount`, `_cumsum`, `select_topk`, `to_onehot`) are NOT used by
this package's metrics — their work is fused into the CUDA kernels (csrc/confmat.cu, csrc/curve.cu); they are kept at the
end of the file as plain tensor helpers for user code written against the reference.
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict
from collections.abc import Mapping, Sequence
from typing import Any, Callable, Optional, Union

import torch
from torch import Tensor


def apply_to_collection(data: Any, dtype: Union[type, tuple], function: Callable, *args: Any, **kwargs: Any) -> Any:
    """Apply ``function`` to every leaf of type ``dtype`` in a nested dict / list / tuple / namedtuple.

    Extra ``*args`` / ``**kwargs`` are forwarded to ``function`` — this is how ``group=`` reaches the
    ``dist_sync_fn`` hook (reference call site: metric.py:518-523).  Containers are rebuilt with their own type.
    """
    if isinstance(data, dtype):
        return function(data, *args, **kwargs)
    if isinstance(data, Mapping):
        items = [(k, apply_to_collection(v, dtype, function, *args, **kwargs)) for k, v in data.items()]
        if isinstance(data, defaultdict):
            return type(data)(data.default_factory, OrderedDict(items))
        return type(data)(OrderedDict(items))
    if isinstance(data, tuple) and hasattr(data, "_fields"):  # namedtuple
        return type(data)(*(apply_to_collection(v, dtype, function, *args, **kwargs) for v in data))
    if isinstance(data, Sequence) and not isinstance(data, str):
        return type(data)([apply_to_collection(v, dtype, function, *args, **kwargs) for v in data])
    return data


def dim_zero_cat(x: Union[Tensor, list[Tensor]]) -> Tensor:
    """Concatenate along dim 0 (scalars become 1-element vectors first); a tensor passes through."""
    if isinstance(x, Tensor):
        return x
    if isinstance(x, (list, tuple)):
        x = [t.unsqueeze(0) if t.numel() == 1 and t.ndim == 0 else t for t in x]
    if not x:
        raise ValueError("No samples to concatenate")
    return torch.cat(x, dim=0)


def dim_zero_sum(x: Tensor) -> Tensor:
    return torch.sum(x, dim=0)


def dim_zero_mean(x: Tensor) -> Tensor:
    return torch.mean(x, dim=0)


def dim_zero_max(x: Tensor) -> Tensor:
    return torch.max(x, dim=0).values


def dim_zero_min(x: Tensor) -> Tensor:
    return torch.min(x, dim=0).values


def _flatten(x: Sequence) -> list:
    """One level of list flattening."""
    return [item for sub in x for item in sub]


def _squeeze_scalar_element_tensor(x: Tensor) -> Tensor:
    return x.squeeze() if x.numel() == 1 else x


def _squeeze_if_scalar(data: Any) -> Any:
    return apply_to_collection(data, Tensor, _squeeze_scalar_element_tensor)


def interp(x: Tensor, xp: Tensor, fp: Tensor) -> Tensor:
    """numpy.interp-like evaluation used by the log-AUC window (reference utilities/data.py:249-272; NOT the variant of
    utilities/compute.py that the macro-averaged curves use): sample points sorted by ``xp``, left-sided interval lookup,
    plain slopes (a repeated ``xp`` gives an infinite slope there, like the reference).  The sort is stable so that curves
    with repeated abscissae keep their original (monotone) ordinate order on every device."""
    order = torch.argsort(xp, stable=True)
    xp, fp = xp[order], fp[order]
    slopes = (fp[1:] - fp[:-1]) / (xp[1:] - xp[:-1])
    idx = torch.clamp(torch.searchsorted(xp, x) - 1, 0, slopes.numel() - 1)
    return fp[idx] + slopes[idx] * (x - xp[idx])


# =========================================================================================================
# Operator-level helpers of the reference that user code imports (utilities/data.py).  None of the metrics in this package
# calls them — the kernels fuse what they do (one-hot, top-k, bincount, cumsum) — they exist so that code written against
# the reference keeps importing and running, on whatever device its tensors live.
# =========================================================================================================
def _flatten_dict(x: dict) -> tuple[dict, bool]:
    """One level of dict flattening; the flag tells whether a key was seen twice (reference :64-78)."""
    flat: dict = {}
    clash = False
    for key, value in x.items():
        for k, v in (value.items() if isinstance(value, dict) else ((key, value),)):
            clash = clash or k in flat
            flat[k] = v
    return flat, clash


def to_onehot(label_tensor: Tensor, num_classes: Optional[int] = None) -> Tensor:
    """``[N, d1, ...]`` labels -> ``[N, C, d1, ...]`` one-hot of the same dtype (reference :81-113)."""
    if num_classes is None:
        num_classes = int(label_tensor.max().item()) + 1
    shape = (label_tensor.shape[0], num_classes, *label_tensor.shape[1:])
    out = torch.zeros(shape, dtype=label_tensor.dtype, device=label_tensor.device)
    return out.scatter_(1, label_tensor.long().unsqueeze(1), 1)


def select_topk(prob_tensor: Tensor, topk: int = 1, dim: int = 1) -> Tensor:
    """int32 mask of the ``topk`` largest entries along ``dim`` (reference :124-148)."""
    if topk == 1:
        best = prob_tensor.argmax(dim=dim, keepdim=True)
    elif prob_tensor.dtype == torch.half and not prob_tensor.is_cuda:  # no half top-k on the CPU
        best = torch.argsort(prob_tensor, dim=dim, stable=True).flip(dim).narrow(dim, 0, topk)
    else:
        best = prob_tensor.topk(k=topk, dim=dim).indices
    return torch.zeros_like(prob_tensor, dtype=torch.int32).scatter_(dim, best, 1)


def to_categorical(x: Tensor, argmax_dim: int = 1) -> Tensor:
    """Scores -> labels along ``argmax_dim`` (reference :151-167)."""
    return torch.argmax(x, dim=argmax_dim)


def _bincount(x: Tensor, minlength: Optional[int] = None) -> Tensor:
    """Occurrences of every value in ``[0, minlength)`` (reference :178-206).  Integer adds commute, so the one
    ``index_add_`` below is exact and run-to-run identical on every device — no separate deterministic-mode branch."""
    if minlength is None:
        minlength = len(torch.unique(x))
    flat = x.reshape(-1).long()
    return torch.zeros(int(minlength), dtype=torch.long, device=x.device).index_add_(0, flat, torch.ones_like(flat))


def _flexible_bincount(x: Tensor) -> Tensor:
    """`_bincount` for values that are not ``0..K-1``: counts of the distinct values, ascending (reference :223-239)."""
    return torch.unique(x, return_counts=True)[1]


def _cumsum(x: Tensor, dim: Optional[int] = 0, dtype: Optional[torch.dtype] = None) -> Tensor:
    """`torch.cumsum` (reference :209-220; its CPU round trip for old torch in deterministic mode is not needed with the
    torch versions this package supports)."""
    return torch.cumsum(x, dim=dim, dtype=dtype)


def allclose(tensor1: Tensor, tensor2: Tensor) -> bool:
    """`torch.allclose` after casting the second tensor to the first one's dtype (reference :242-246)."""
    return torch.allclose(tensor1, tensor2.to(tensor1.dtype))
