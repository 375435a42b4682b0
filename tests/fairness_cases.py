"""Shared by the CPU host-layer test and the GPU parity test of the group-fairness metrics: replay every case of
tests/golden/fairness.npz (generated from the unmodified reference by make_golden.py fairness) on a device."""
import warnings

import numpy as np
import torch

from metrics_b200.classification import BinaryFairness, BinaryGroupStatRates
from metrics_b200.functional.classification import (
    binary_fairness,
    binary_groups_stat_rates,
    demographic_parity,
    equal_opportunity,
)
from metrics_b200.functional.classification.group_fairness import _binary_groups_stat_scores, _group_counts


def _as_dict(keys, values):
    return dict(zip(str(keys).split(","), np.asarray(values, dtype=np.float32).tolist()))


def _close(got: dict, want: dict, what: str):
    assert list(got) == list(want), f"{what}: keys {list(got)} != {list(want)}"
    for k in want:
        g, w = float(got[k]), want[k]
        assert (np.isnan(g) and np.isnan(w)) or abs(g - w) <= 1e-6 * max(1.0, abs(w)), f"{what}[{k}]: {g} != {w}"


def replay(golden, device: str) -> int:
    n_cases = int(golden["n_cases"])
    for c in range(n_cases):
        key = f"case{c}"
        num_groups, ign = (int(v) for v in golden[f"{key}/meta"])
        ign = None if ign == -999 else ign
        preds = torch.from_numpy(golden[f"{key}/preds"]).to(device)
        target = torch.from_numpy(golden[f"{key}/target"]).to(device)
        groups = torch.from_numpy(golden[f"{key}/groups"]).to(device)
        want_counts = torch.from_numpy(golden[f"{key}/counts"])
        # integer counters: bit-exact, through the kernel-backed core and through the reference's list-of-tuples seam
        assert torch.equal(_group_counts(preds, target, groups, num_groups, 0.5, ign, True).cpu(), want_counts), key
        stats = _binary_groups_stat_scores(preds, target, groups, num_groups, 0.5, ign, True)
        assert torch.equal(torch.stack([torch.stack(s) for s in stats]).cpu(), want_counts), key
        rates = binary_groups_stat_rates(preds, target, groups, num_groups, 0.5, ign)
        want_rates = golden[f"{key}/rates"]
        assert list(rates) == [f"group_{i}" for i in range(num_groups)]
        np.testing.assert_allclose(torch.stack(list(rates.values())).cpu().numpy(), want_rates, rtol=1e-6, atol=0, equal_nan=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = _as_dict(golden[f"{key}/fair_keys"], golden[f"{key}/fair_values"])
            _close(binary_fairness(preds, target, groups, "all", 0.5, ign), want, f"{key} functional")
            dp_key, eo_key = list(want)
            _close(equal_opportunity(preds, target, groups, 0.5, ign), {eo_key: want[eo_key]}, f"{key} EO")
            _close(binary_fairness(preds, target, groups, "equal_opportunity", 0.5, ign), {eo_key: want[eo_key]}, f"{key} EO task")
            dp = demographic_parity(preds, groups, 0.5, ign)
            assert list(dp)[0].startswith("DP_") and len(dp) == 1
            if ign is None:  # with an ignore_index the all-zero stand-in target changes which samples count (reference too)
                _close(dp, {dp_key: want[dp_key]}, f"{key} DP")
            n = preds.shape[0]
            metric = BinaryFairness(num_groups, ignore_index=ign).to(device)
            metric.update(preds[: n // 2], target[: n // 2], groups[: n // 2])
            metric.update(preds[n // 2:], target[n // 2:], groups[n // 2:])
            _close(metric.compute(), _as_dict(golden[f"{key}/fair2_keys"], golden[f"{key}/fair2_values"]), f"{key} class")
            table = BinaryGroupStatRates(num_groups, ignore_index=ign).to(device)
            table(preds, target, groups)
            np.testing.assert_allclose(torch.stack(list(table.compute().values())).cpu().numpy(), want_rates, rtol=1e-6,
                                       equal_nan=True)
            assert torch.equal(torch.stack((table.tp, table.fp, table.tn, table.fn), 1).cpu(), want_counts)
    return n_cases
