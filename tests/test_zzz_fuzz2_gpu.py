"""GPU: the second, larger randomised differential draw (tests/golden/fuzz2.npz, 394 cases from the unmodified reference,
make_golden.py fuzz2) through the kernels.  The very last file on purpose: it was generated after the round's GPU budget was
spent and has so far only been replayed on the kernel stand-ins (tests/test_fuzz_host.py::test_second_draw).

float32 / float64 / integer inputs must meet the same bars as the first draw.  Half-precision inputs (whose tolerances in
tests/fuzz_cases.py were tuned on the first draw only) are reported as `xfail` instead of failing when they miss them, so that
an untriaged rounding corner shows up in the report without masking the rest of the run."""
import json

import pytest

from tests.fuzz_cases import n_cases, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", range(n_cases("fuzz2")))
def test_case(golden_fuzz2, k):
    spec = json.loads(str(golden_fuzz2[f"{k}/spec"]))
    half = spec["preds_dtype"] in ("float16", "bfloat16")
    # exact curves list one point per DISTINCT score: two logits whose float32 sigmoids differ by one ulp on one device and
    # coincide on the other change the number of points (the reference's own CPU and CUDA results differ the same way)
    curve = any(token in spec["fn"] for token in ("roc", "curve"))
    try:
        run_case(golden_fuzz2, k, "cuda:0")
    except AssertionError as err:
        if half:
            pytest.xfail(f"half-precision case outside the first draw's tolerances, to triage: {str(err)[:300]}")
        if curve and "shape" in str(err).lower():
            pytest.xfail(f"different number of distinct thresholds (1-ulp sigmoid / softmax ties), to triage: {str(err)[:300]}")
        raise
