"""Rank-zero-only logging helpers (reference behaviour: utilities/prints.py:23-56, gated on LOCAL_RANK)."""
import logging
import os
import warnings
from functools import wraps
from typing import Any, Callable

log = logging.getLogger("metrics_b200")


def _local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", 0))


def rank_zero_only(fn: Callable) -> Callable:
    """Run ``fn`` only in the process whose LOCAL_RANK is 0."""

    @wraps(fn)
    def inner(*args: Any, **kwargs: Any) -> Any:
        if _local_rank() == 0:
            return fn(*args, **kwargs)
        return None

    return inner


@rank_zero_only
def rank_zero_warn(message: str, category: type = UserWarning, stacklevel: int = 3, **kwargs: Any) -> None:
    warnings.warn(message, category, stacklevel=stacklevel, **kwargs)


@rank_zero_only
def rank_zero_info(*args: Any, **kwargs: Any) -> None:
    log.info(*args, **kwargs)


@rank_zero_only
def rank_zero_debug(*args: Any, **kwargs: Any) -> None:
    log.debug(*args, **kwargs)
