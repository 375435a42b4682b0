"""GPU: K6 (`normalize_logits_if_needed`, utilities/compute.py:190-229) is BIT-IDENTICAL to the ATen ops the reference calls
on the same device — `tensor.sigmoid()` and `torch.softmax(tensor, dim=1)` on CUDA — so the set of distinct thresholds of an
exact curve (one per distinct score) is the one the reference gets on a B200.  This is a floating-point kernel: its
reference is the plain torch op, the bar is equality (tolerance 0), written here."""
import pytest
import torch

from metrics_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n", [1, 33, 4097, 1 << 20])
def test_sigmoid_matches_aten_cuda_bitwise(dtype, n):
    g = torch.Generator().manual_seed(n)
    x = (torch.randn(n, generator=g) * 6).to(dtype).to(DEV)
    x[0] = -0.5  # make sure the batch counts as logits
    if n > 8:
        x[1:8] = torch.tensor([0.0, -0.0, 88.0, -88.0, 104.0, -104.0, 1e-30], dtype=dtype, device=DEV)
    got = _native.sigmoid_if_logits(x)
    assert got.dtype == dtype and torch.equal(got, torch.sigmoid(x))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(7, 2), (300, 5), (257, 11), (64, 33), (1025, 100), (512, 1000), (96, 1024)])
def test_softmax_matches_aten_cuda_bitwise(dtype, shape):
    """C <= 1024: ATen's CUDA softmax runs its warp kernel (lane-strided sums, butterfly reduction) — the same summation
    tree as K6's, so even the last bit agrees."""
    g = torch.Generator().manual_seed(shape[0] * 31 + shape[1])
    x = (torch.randn(*shape, generator=g) * 3).to(dtype).to(DEV)
    got = _native.softmax_if_logits(x)
    assert got.dtype == dtype and torch.equal(got, torch.softmax(x, dim=1))


@pytest.mark.parametrize("shape", [(64, 1500), (33, 4096)])
def test_softmax_wide_rows_within_one_ulp(shape):
    """C > 1024: ATen switches to a block-wide kernel with another summation order; results agree to float32 rounding of the
    row sum (<= 2 ulp), far inside the 1e-6 relative bar of the curve metrics."""
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(*shape, generator=g) * 3).to(DEV)
    got, want = _native.softmax_if_logits(x), torch.softmax(x, dim=1)
    assert torch.allclose(got, want, rtol=5e-7, atol=0.0)


def test_probabilities_pass_through_untouched():
    p = torch.rand(1000, device=DEV)
    assert torch.equal(_native.sigmoid_if_logits(p), p)
    q = torch.softmax(torch.randn(50, 9, device=DEV), 1)
    assert torch.equal(_native.softmax_if_logits(q), q)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n", [32769, 1 << 20, (1 << 22) + 3, 4096 * 257 + 1])
def test_speculative_single_pass_sigmoid(dtype, n):
    """Large batches take the speculative single pass (tile-local vote, pending tiles, fix-up launch): logits, probabilities,
    and the adversarial mix — in-range stretches of tiles with ONE out-of-range score at the very end (or the very start)."""
    g = torch.Generator().manual_seed(n)
    logits = (torch.randn(n, generator=g) * 5).to(dtype).to(DEV)
    assert torch.equal(_native.sigmoid_if_logits(logits), torch.sigmoid(logits))
    probs = torch.rand(n, generator=g).to(dtype).to(DEV)
    assert torch.equal(_native.sigmoid_if_logits(probs), probs)
    for where in (n - 1, 0, n // 2):
        mixed = probs.clone()
        mixed[where] = -0.25  # a single logit: every other tile is pending and must be revisited
        assert torch.equal(_native.sigmoid_if_logits(mixed), torch.sigmoid(mixed))
    nan = probs.clone()
    nan[17] = float("nan")  # NaN compares false: still probabilities
    got = _native.sigmoid_if_logits(nan)
    assert torch.equal(got.nan_to_num(7.0), nan.nan_to_num(7.0))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_speculative_single_pass_softmax_mixed_batches(dtype):
    """Rows entirely inside [0, 1] are written through speculatively and must be revisited when ANOTHER row makes the batch
    logits (the reference applies softmax to the whole tensor, utilities/compute.py:223-229); a batch of probabilities stays as
    it is; wide rows (C > 1024) keep the original kernels."""
    g = torch.Generator().manual_seed(12)
    probs = torch.softmax(torch.randn(3000, 37, generator=g), 1).to(dtype).to(DEV)
    assert torch.equal(_native.softmax_if_logits(probs), probs)
    for where in (0, 1499, 2999):
        mixed = probs.clone()
        mixed[where, 5] = -2.0  # one logit row: every other row was pending
        assert torch.equal(_native.softmax_if_logits(mixed), torch.softmax(mixed, dim=1))
    logits = (torch.randn(2048, 1000, generator=g) * 3).to(dtype).to(DEV)
    assert torch.equal(_native.softmax_if_logits(logits), torch.softmax(logits, dim=1))
