"""CPU: the randomised differential replay of tests/test_fuzz_gpu.py with the kernel wrappers replaced by their torch stand-ins
(fixture `cpu_kernel_standins`): checks the host layer above the C-ABI — validation, format seams, reducers, output layouts —
against the unmodified reference's outputs without a GPU.  The kernels themselves are checked by the GPU twin."""
import pytest

from tests.fuzz_cases import n_cases, run_case


@pytest.mark.parametrize("k", range(n_cases()))
def test_case(golden_fuzz, cpu_kernel_standins, k):
    run_case(golden_fuzz, k, "cpu")


@pytest.mark.parametrize("k", range(n_cases("fuzz2")))
def test_second_draw(golden_fuzz2, cpu_kernel_standins, k):
    """A second, three times larger draw with another seed (make_golden.py fuzz2)."""
    run_case(golden_fuzz2, k, "cpu")
