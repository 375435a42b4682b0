"""GPU: randomised differential replay.  tests/golden/fuzz.npz holds ~90 random (functional, kwargs, inputs) cases with
the unmodified reference's outputs (make_golden.py fuzz); every case is re-run through the product functionals (C-ABI
kernels + reducers).  Integer outputs must be bit-exact, floating outputs within 1e-6 relative (+1e-7 absolute; kappa /
MCC and half-precision inputs get the tolerances stated below)."""
import pytest

from tests.fuzz_cases import n_cases, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", range(n_cases()))
def test_case(golden_fuzz, k):
    run_case(golden_fuzz, k, "cuda:0")
