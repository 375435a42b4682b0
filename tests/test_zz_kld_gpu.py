"""GPU: KL divergence through the row kernel (`mb200_kl_divergence_rows`, csrc/kldiv.cu) against goldens from the unmodified
reference (float32 within 1e-6, float64 within 1e-12), the fp64 oracle on larger and awkward shapes, half precision."""
import numpy as np
import pytest
import torch

from oracle import regression as oreg
from tests.kld_cases import argument_errors, replay

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_replay_reference_goldens_on_the_kernel(golden_kld):
    assert replay(golden_kld, DEV) == 24


def test_argument_errors():
    argument_errors(DEV)


@pytest.mark.parametrize("shape", [(1, 1), (3, 31), (9, 32), (17, 33), (4099, 7), (1000, 1000), (2, 70001)])
@pytest.mark.parametrize("log_prob", [False, True])
def test_rows_vs_oracle(shape, log_prob):
    from metrics_b200 import _native

    g = torch.Generator().manual_seed(shape[0] * 31 + shape[1])
    p = torch.rand(shape, generator=g) + 1e-4
    q = torch.rand(shape, generator=g) + 1e-4
    if log_prob:
        p, q = torch.log_softmax(p * 4, 1), torch.log_softmax(q * 4, 1)
    else:
        p[:, ::5] = 0.0
        if shape[1] > 1:
            p[:, 1] = 0.5
    want = oreg.kl_divergence_rows(p.numpy(), q.numpy(), log_prob)
    got = _native.kl_divergence_rows(p.to(DEV), q.to(DEV), log_prob)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-6, atol=2e-7)
    got64 = _native.kl_divergence_rows(p.double().to(DEV), q.double().to(DEV), log_prob)
    np.testing.assert_allclose(got64.cpu().numpy(), oreg.kl_divergence_rows(p.double().numpy(), q.double().numpy(), log_prob),
                               rtol=1e-12, atol=1e-14)
    for half in (torch.float16, torch.bfloat16):
        ph, qh = p.to(half), q.to(half)
        got_h = _native.kl_divergence_rows(ph.to(DEV), qh.to(DEV), log_prob)
        assert got_h.dtype == half
        want_h = oreg.kl_divergence_rows(ph.float().numpy(), qh.float().numpy(), log_prob)
        np.testing.assert_allclose(got_h.float().cpu().numpy(), want_h, rtol=1e-2, atol=2e-3)


def test_special_values():
    """q = 0 where p > 0 -> inf; p = 0 -> the term is 0 whatever q is; NaN propagates (reference `res[x == 0] = 0`)."""
    from metrics_b200 import _native

    p = torch.tensor([[0.5, 0.5, 0.0], [0.0, 1.0, 0.0], [0.2, float("nan"), 0.8]], device=DEV)
    q = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.3, 0.3, 0.4]], device=DEV)
    got = _native.kl_divergence_rows(p, q, False).cpu()
    assert torch.isinf(got[0]) and got[0] > 0 and float(got[1]) == 0.0 and torch.isnan(got[2])
