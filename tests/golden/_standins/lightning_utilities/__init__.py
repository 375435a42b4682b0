"""STAND-IN for the third-party `lightning_utilities` package (NOT the real package).

The reference (TorchMetrics) hard-depends on `lightning-utilities`, which is not installed in the build
container and cannot be fetched (no network).  The reference uses exactly four symbols from it; this
directory re-creates their documented behaviour so that the *unmodified* reference
under /root/reference can be imported by `tests/golden/make_golden.py` to generate golden vectors.

It is test tooling only: nothing in `metrics_b200/` imports it.
"""
from lightning_utilities.core.apply_func import apply_to_collection  # noqa: F401
