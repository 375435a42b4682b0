"""State-reduction helpers of the runtime (reference: utilities/data.py:29-61, 150-176).

The arithmetic helpers that the reference keeps here (`_bincount`, `_cumsum`) have no equivalent in this
package: they are fused into the CUDA kernels (see csrc/confmat.cu, csrc/curve.cu).
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict
from collections.abc import Mapping, Sequence
from typing import Any, Callable, Union

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
