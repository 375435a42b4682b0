"""GPU: critical success index through the counting kernel (`mb200_binary_stat_counts`, integer-label mode) against goldens
from the unmodified reference.  (Sorts last on purpose: written after the round's GPU budget was spent; so far checked on the
kernel's CPU stand-in only — tests/test_csi_host.py.)"""
import pytest

from tests.csi_cases import argument_errors, replay

pytestmark = pytest.mark.gpu


def test_replay_reference_goldens_on_the_kernel(golden_csi):
    assert replay(golden_csi, "cuda") == 96


def test_argument_errors():
    argument_errors("cuda")
