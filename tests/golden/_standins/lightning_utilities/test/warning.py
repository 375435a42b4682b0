"""Stand-in for `lightning_utilities.test.warning` (test-only dependency of the reference's unit tests)."""
import re
import warnings
from contextlib import contextmanager


@contextmanager
def no_warning_call(expected_warning=Warning, match=None):
    """Fail if the body raises a warning of the given category (and message pattern)."""
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        yield
    for w in caught:
        if issubclass(w.category, expected_warning) and (match is None or re.search(match, str(w.message))):
            raise AssertionError(f"`{w.category.__name__}` was raised: {w.message}")
