"""CPU: operating-point selection (masked lexicographic arg-max, functional/classification/at_fixed.py) on curves computed
by the numpy oracle, against goldens from the unmodified reference (make_golden.py atfixed)."""
import numpy as np
import pytest
import torch

from metrics_b200.functional.classification.at_fixed import _FAMILIES
from oracle import curves as oc

FLOORS = [0.0, 0.35, 0.6, 0.9, 1.0]


@pytest.mark.parametrize("fam", sorted(_FAMILIES))
@pytest.mark.parametrize("floor", FLOORS)
def test_binary_exact_selection(golden_atfixed, fam, floor):
    g = golden_atfixed
    p, t = g["b/preds"], g["b/target"]
    if _FAMILIES[fam].curve == "prc":
        a, b, thr = oc.binary_prc_ref32(p, t)
    else:
        a, b, thr = oc.binary_roc_ref32(p, t)
    val, th = _FAMILIES[fam].pick(torch.from_numpy(a.copy()), torch.from_numpy(b.copy()), torch.from_numpy(thr.copy()), floor)
    np.testing.assert_allclose(val.numpy(), g[f"b/{fam}/{floor}/exact/value"], rtol=1e-6)
    np.testing.assert_allclose(th.numpy(), g[f"b/{fam}/{floor}/exact/thr"], rtol=1e-6)


@pytest.mark.parametrize("fam", sorted(_FAMILIES))
def test_multilabel_exact_selection_with_ignore(golden_atfixed, fam):
    g = golden_atfixed
    p, t = g["ml/preds"], g["ml/target_ign"]
    curves = oc.multilabel_prc_ref32(p, t, -1) if _FAMILIES[fam].curve == "prc" else oc.multilabel_roc_ref32(p, t, -1)
    for floor in FLOORS:
        vals, thrs = [], []
        for a, b, thr in curves:
            v, th = _FAMILIES[fam].pick(torch.from_numpy(a.copy()), torch.from_numpy(b.copy()), torch.from_numpy(thr.copy()), floor)
            vals.append(float(v)), thrs.append(float(th))
        np.testing.assert_allclose(vals, g[f"mli/{fam}/{floor}/exact/value"], rtol=1e-6)
        np.testing.assert_allclose(thrs, g[f"mli/{fam}/{floor}/exact/thr"], rtol=1e-6)
