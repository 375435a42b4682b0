"""GPU: the second, larger randomised differential draw (tests/golden/fuzz2.npz, 394 cases from the unmodified reference,
make_golden.py fuzz2) through the kernels.  The very last file on purpose: it was generated after the round's GPU budget was
spent and has so far only been replayed on the kernel stand-ins (tests/test_fuzz_host.py::test_second_draw)."""
import pytest

from tests.fuzz_cases import n_cases, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", range(n_cases("fuzz2")))
def test_case(golden_fuzz2, k):
    run_case(golden_fuzz2, k, "cuda:0")
