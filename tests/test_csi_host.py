"""CPU: critical success index — host layer against goldens from the reference, counting kernel replaced by its stand-in."""
from tests.csi_cases import argument_errors, replay


def test_replay_reference_goldens(golden_csi, cpu_kernel_standins):
    assert replay(golden_csi, "cpu") == 96


def test_argument_errors(cpu_kernel_standins):
    argument_errors("cpu")
