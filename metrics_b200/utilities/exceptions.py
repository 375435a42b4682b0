"""Exception types kept name-compatible with the reference (utilities/exceptions.py:14-21)."""


class TorchMetricsUserError(Exception):
    """Raised on misuse of the metric API (double sync, unsync without sync, forward while synced)."""


class TorchMetricsUserWarning(Warning):
    """Warning category for recoverable misuse."""
