"""Shared by tests/test_csi_host.py (CPU, kernel stand-in) and tests/test_zz_csi_gpu.py (counting kernel K2): replay
tests/golden/csi.npz, produced by the unmodified reference (make_golden.py csi)."""
import numpy as np
import pytest
import torch

from metrics_b200 import CriticalSuccessIndex
from metrics_b200.functional import critical_success_index
from metrics_b200.functional.regression.csi import _critical_success_index_update

_DTYPES = (torch.float32, torch.float64, torch.float16)


def replay(g, device: str) -> int:
    n = int(g["n_cases"])
    for c in range(n):
        key = f"case{c}"
        threshold, keep, dt = g[f"{key}/meta"].tolist()
        keep = None if keep < 0 else int(keep)
        preds = torch.from_numpy(g[f"{key}/preds"]).to(_DTYPES[int(dt)]).to(device)
        target = torch.from_numpy(g[f"{key}/target"]).to(_DTYPES[int(dt)]).to(device)
        hits, misses, false_alarms = _critical_success_index_update(preds, target, threshold, keep)
        assert hits.dtype == torch.int32
        np.testing.assert_array_equal(torch.stack([hits, misses, false_alarms]).cpu().numpy(), g[f"{key}/counts"], err_msg=key)
        value = critical_success_index(preds, target, threshold, keep)
        np.testing.assert_allclose(value.cpu().numpy(), g[f"{key}/value"], rtol=1e-6, err_msg=key)
        metric = CriticalSuccessIndex(threshold, keep_sequence_dim=keep).to(device)
        metric.update(preds, target)
        metric.update(target, preds)
        np.testing.assert_allclose(metric.compute().cpu().numpy(), g[f"{key}/class_value"], rtol=1e-6, err_msg=key)
    return n


def argument_errors(device: str) -> None:
    x = torch.rand(4, 3, device=device)
    with pytest.raises(ValueError, match="Expected keep_sequence dim to be in range"):
        critical_success_index(x, x, 0.5, keep_sequence_dim=2)
    with pytest.raises(ValueError, match="Expected keep_sequence_dim to be a non-negative integer"):
        CriticalSuccessIndex(0.5, keep_sequence_dim=-1)
    with pytest.raises(RuntimeError, match="same shape"):
        critical_success_index(x, x[:2], 0.5)
    none_above = critical_success_index(x, x, 2.0)
    assert float(none_above) == 0.0  # 0 / 0 -> 0 (`_safe_divide`)
