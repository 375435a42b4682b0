"""Shared by tests/test_tweedie_host.py (CPU, kernel stand-in) and tests/test_zz_tweedie_gpu.py (kernel K9 op 10): replay
tests/golden/tweedie.npz, produced by the unmodified reference (make_golden.py tweedie)."""
import numpy as np
import pytest
import torch

from metrics_b200 import TweedieDevianceScore
from metrics_b200.functional import tweedie_deviance_score
from metrics_b200.functional.regression.tweedie_deviance import _tweedie_deviance_score_update


def replay(g, device: str) -> int:
    n = int(g["n_cases"])
    for c in range(n):
        key = f"case{c}"
        preds, targets = torch.from_numpy(g[f"{key}/preds"]).to(device), torch.from_numpy(g[f"{key}/targets"]).to(device)
        power = float(g[f"{key}/power"])
        # float32: the reference evaluates the terms with ATen's float32 pow / log and sums them with a float32 `torch.sum`;
        # the kernel uses CUDA's float32 math and a float64 accumulator -> same tolerance as the other K9 ops (test_fuzz)
        tol = dict(rtol=1e-5, atol=0) if preds.dtype == torch.float32 else dict(rtol=1e-12, atol=0)
        value = tweedie_deviance_score(preds, targets, power)
        assert value.dtype == preds.dtype and value.ndim == 0
        np.testing.assert_allclose(value.cpu().numpy(), g[f"{key}/value"], err_msg=key, **tol)
        total, count = _tweedie_deviance_score_update(preds, targets, power)
        np.testing.assert_allclose(total.cpu().numpy(), g[f"{key}/sum"], err_msg=key, **tol)
        assert int(count) == int(g[f"{key}/count"]) == preds.numel()
        metric = TweedieDevianceScore(power=power).to(device)
        half = preds.shape[0] // 2
        metric.update(preds[:half], targets[:half])
        metric.update(preds[half:], targets[half:])
        np.testing.assert_allclose(metric.compute().cpu().numpy(), g[f"{key}/class_value"], err_msg=key, rtol=1e-5)
    return n


def domain_errors(device: str) -> None:
    pos = torch.tensor([1.0, 2.0, 3.0], device=device)
    with_zero, with_neg = torch.tensor([1.0, 0.0, 3.0], device=device), torch.tensor([1.0, -2.0, 3.0], device=device)
    with pytest.raises(ValueError, match="Deviance Score is not defined for power=0.5"):
        tweedie_deviance_score(pos, pos, power=0.5)
    with pytest.raises(ValueError, match="Deviance Score is not defined for power=0.5"):
        TweedieDevianceScore(power=0.5)
    with pytest.raises(RuntimeError, match="same shape"):
        tweedie_deviance_score(pos, pos[:2], power=1)
    for power, preds, targets, message in (
        (1, with_zero, pos, "'preds' has to be strictly positive and 'targets' cannot be negative"),
        (1, pos, with_neg, "'preds' has to be strictly positive and 'targets' cannot be negative"),
        (2, pos, with_zero, "both 'preds' and 'targets' have to be strictly positive"),
        (2, with_neg, pos, "both 'preds' and 'targets' have to be strictly positive"),
        (-1, with_zero, pos, "'preds' has to be strictly positive."),
        (1.5, pos, with_neg, "'targets' has to be strictly positive and 'preds' cannot be negative"),
        (3, pos, with_zero, "both 'preds' and 'targets' have to be strictly positive"),
    ):
        with pytest.raises(ValueError, match=f"For power={power}, " + message):
            tweedie_deviance_score(preds, targets, power=power)
    # legal corners: zero targets for Poisson / compound Poisson, anything for power 0, negative targets for power < 0
    assert float(tweedie_deviance_score(pos, with_zero, power=1)) > 0
    assert float(tweedie_deviance_score(pos, with_zero, power=1.5)) > 0
    assert float(tweedie_deviance_score(with_neg, with_zero, power=0)) == pytest.approx((0 + 4 + 0) / 3)
    assert torch.isfinite(tweedie_deviance_score(pos, with_neg, power=-1))
    ints = tweedie_deviance_score(torch.tensor([4, 3, 2, 1], device=device), torch.tensor([1, 2, 3, 4], device=device), power=2)
    assert float(ints) == pytest.approx(1.2083, abs=1e-4)  # the reference's docstring value
