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


def _unwrapped(fn):
    """Peel `functools.wraps` layers, `Mock(wraps=...)` and `functools.partial` off a callable."""
    from functools import partial
    from unittest.mock import Mock

    for _ in range(16):
        if fn is None:
            break
        if hasattr(fn, "__wrapped__"):
            fn = fn.__wrapped__
        elif isinstance(fn, Mock):
            fn = fn._mock_wraps
        elif isinstance(fn, partial):
            fn = fn.func
        else:
            break
    return fn


def is_overridden(method_name: str, instance: object, parent: object) -> bool:
    """Whether ``instance`` provides its own ``method_name`` rather than ``parent``'s (reference checks.py:740-762);
    decorated, mocked and partial-bound methods are compared by the function underneath."""
    own = _unwrapped(getattr(instance, method_name, None))
    if own is None:
        return False
    inherited = getattr(parent, method_name, None)
    if inherited is None:
        raise ValueError("The parent should define the method")
    return own.__code__ is not inherited.__code__


def check_forward_full_state_property(metric_class, init_args=None, input_args=None, num_update_to_compare=[10, 100, 1000],  # noqa: B006 (read only; the reference's default)
                                      reps: int = 5) -> None:
    """Tell whether ``full_state_update = False`` is safe for ``metric_class`` and whether it is faster
    (reference checks.py:635-737): run ``forward`` both ways on the same inputs, compare every batch value and the final
    ``compute()``; when they agree, time both and print the recommendation in the reference's wording."""
    from time import perf_counter

    import torch

    init_args, input_args = init_args or {}, input_args or {}
    variants = [type(name, (metric_class,), {"full_state_update": flag})(**init_args)
                for name, flag in (("FullState", True), ("PartState", False))]
    full, part = variants
    same = True
    try:  # an exception in the reduced-state path means `update` needs the accumulated state
        for _ in range(num_update_to_compare[0]):
            same = same and _allclose_recursive(full(**input_args), part(**input_args))
        same = same and _allclose_recursive(full.compute(), part.compute())
    except RuntimeError:
        same = False
    if not same:
        print("Recommended setting `full_state_update=True`")
        return
    seconds = torch.zeros(2, len(num_update_to_compare), reps)
    for i, metric in enumerate(variants):
        for j, steps in enumerate(num_update_to_compare):
            for r in range(reps):
                start = perf_counter()
                for _ in range(steps):
                    metric(**input_args)
                seconds[i, j, r] = perf_counter() - start
                metric.reset()
    mean, std = seconds.mean(-1), seconds.std(-1)
    for j, steps in enumerate(num_update_to_compare):
        print(f"Full state for {steps} steps took: {mean[0, j]}+-{std[0, j]:0.3f}")
        print(f"Partial state for {steps} steps took: {mean[1, j]:0.3f}+-{std[1, j]:0.3f}")
    print(f"Recommended setting `full_state_update={not bool(mean[1, -1] < mean[0, -1])}`")
