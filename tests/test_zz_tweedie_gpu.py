"""GPU: Tweedie deviance through kernel K9 op 10 (`mb200_regression_sums`) against goldens from the unmodified reference.
(Sorts last on purpose: written after the round's GPU budget was spent; until the round-end run it has been checked on the
kernel's CPU stand-in and, formula by formula, through the host build of csrc/regression_terms.cuh — tests/test_tweedie_host.py.)"""
import pytest
import torch

from tests.tweedie_cases import domain_errors, replay

pytestmark = pytest.mark.gpu


def test_replay_reference_goldens_on_the_kernel(golden_tweedie):
    assert replay(golden_tweedie, "cuda") == 24


def test_domain_errors_and_corners():
    domain_errors("cuda")


def test_large_input_matches_float64_torch():
    g = torch.Generator(device="cuda").manual_seed(9)
    preds = torch.rand(1 << 22, device="cuda", generator=g) * 5 + 0.1
    targets = torch.rand(1 << 22, device="cuda", generator=g) * 5 + 0.1
    from metrics_b200.functional import tweedie_deviance_score

    p, t = preds.double(), targets.double()
    for power, want in ((1.0, 2 * (t * torch.log(t / p) + p - t)), (2.0, 2 * (torch.log(p / t) + t / p - 1)),
                        (3.0, 2 * (t.pow(-1) / 2 + t * p.pow(-2) / 2 - p.pow(-1)))):
        got = float(tweedie_deviance_score(preds, targets, power))
        assert got == pytest.approx(float(want.mean()), rel=2e-5), power
