"""CPU: the numpy curve oracle reproduces the goldens produced by the unmodified reference."""
import numpy as np
import pytest

from oracle import curves as oc
from tests.helpers import MC_CASES, cfg3_inputs, mc_inputs, sha

BIN_CASES = ["doc", "rand", "ties", "logits", "allpos", "allneg", "alltied", "one", "odd", "skew"]
RTOL = 1e-6  # north_star tolerance for floating-point AUROC / AP


@pytest.mark.parametrize("name", BIN_CASES)
def test_binary_cases(golden_curves, name):
    g = golden_curves
    p = oc.sigmoid_if_logits(g[f"bin/{name}/preds"])
    t = g[f"bin/{name}/target"]
    fps, tps, thr = oc.binary_clf_curve(p, t)
    np.testing.assert_array_equal(fps, g[f"bin/{name}/clf_fps"].astype(np.int64))
    np.testing.assert_array_equal(tps, g[f"bin/{name}/clf_tps"].astype(np.int64))
    np.testing.assert_allclose(thr, g[f"bin/{name}/clf_thr"], rtol=1e-6)  # sigmoid may differ by an ulp between libm and ATen
    fpr, tpr, th = oc.binary_roc_ref32(p, t)
    np.testing.assert_allclose(fpr, g[f"bin/{name}/roc_fpr"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(tpr, g[f"bin/{name}/roc_tpr"], rtol=RTOL, atol=0)
    pr, rc, th2 = oc.binary_prc_ref32(p, t)
    np.testing.assert_allclose(pr, g[f"bin/{name}/prc_p"], rtol=RTOL, atol=0, equal_nan=True)
    np.testing.assert_allclose(rc, g[f"bin/{name}/prc_r"], rtol=RTOL, atol=0, equal_nan=True)
    # scalars: reference-dtype restatement and the exact value both sit within the tolerance
    np.testing.assert_allclose(oc.binary_auroc_ref32(p, t), g[f"bin/{name}/auroc"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(oc.binary_auroc_exact(p, t), g[f"bin/{name}/auroc"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(oc.binary_average_precision_ref32(p, t), g[f"bin/{name}/ap"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(oc.binary_average_precision_exact(p, t), g[f"bin/{name}/ap"], rtol=RTOL, atol=1e-7)
    for mf in (0.5, 0.8):
        np.testing.assert_allclose(oc.binary_auroc_ref32(p, t, max_fpr=mf), g[f"bin/{name}/auroc_maxfpr{mf}"], rtol=2e-6, atol=1e-7)


def test_cfg3_full_size(golden_curves):
    preds, target = cfg3_inputs()
    assert sha(preds) == str(golden_curves["cfg3/preds_sha256"])
    p, t = preds.reshape(-1).numpy(), target.reshape(-1).numpy()
    np.testing.assert_allclose(oc.binary_auroc_exact(p, t), golden_curves["cfg3/auroc"], rtol=RTOL)
    np.testing.assert_allclose(oc.binary_average_precision_exact(p, t), golden_curves["cfg3/ap"], rtol=RTOL)
    assert float(golden_curves["cfg3/auroc"]) == pytest.approx(0.49975747, abs=1e-7)


@pytest.mark.parametrize("C,N,kind", MC_CASES)
def test_multiclass_cases(golden_curves, C, N, kind):
    p, t = mc_inputs(C, N, kind)
    key = f"mc/C{C}_{kind}"
    if C > 37:
        assert sha(p) == str(golden_curves[f"{key}/preds_sha256"])
    pn = oc.softmax_if_logits(p.numpy())
    tn = t.numpy()
    auc = oc.multiclass_auroc_exact(pn, tn, C)
    ap = oc.multiclass_average_precision_exact(pn, tn, C)
    w = np.bincount(tn, minlength=C).astype(np.float64)
    for avg in ("macro", "weighted", "none"):
        np.testing.assert_allclose(oc.reduce_per_class(auc, avg, w), golden_curves[f"{key}/auroc_{avg}"], rtol=2e-6, atol=1e-7, equal_nan=True)
        np.testing.assert_allclose(oc.reduce_per_class(ap, avg, w), golden_curves[f"{key}/ap_{avg}"], rtol=2e-6, atol=1e-7, equal_nan=True)
