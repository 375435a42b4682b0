"""CPU: the oracle's multilabel curve restatement (oracle/curves.py) against goldens produced by the unmodified reference
(tests/golden/make_golden.py multilabel)."""
import numpy as np
import pytest

from oracle import curves as oc

CASES = ["L4_probs", "L6_logits", "L3_ties", "L5_extra"]


@pytest.fixture(scope="module")
def g(golden_multilabel):
    return golden_multilabel


def _inputs(g, name, ign):
    p, t = g[f"{name}/preds"], g[f"{name}/target_ignore" if ign else f"{name}/target"]
    p, t = oc.multilabel_flatten(p, t)
    return oc.sigmoid_if_logits(p), t


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ign", [False, True])
def test_scalars(g, name, ign):
    p, t = _inputs(g, name, ign)
    ig, tag = (-1, "ign_") if ign else (None, "")
    auroc = oc.multilabel_auroc_exact(p, t, ig)
    ap = oc.multilabel_average_precision_exact(p, t, ig)
    w = oc.multilabel_positive_counts(t)
    np.testing.assert_allclose(auroc, g[f"{name}/{tag}auroc_none"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(ap, g[f"{name}/{tag}ap_none"], rtol=2e-6, atol=1e-7)
    for avg in ("macro", "weighted"):
        np.testing.assert_allclose(oc.reduce_per_class(auroc, avg, w), g[f"{name}/{tag}auroc_{avg}"], rtol=2e-6)
        np.testing.assert_allclose(oc.reduce_per_class(ap, avg, w), g[f"{name}/{tag}ap_{avg}"], rtol=2e-6)
    pm, tm = oc.multilabel_micro(p, t, ig)
    np.testing.assert_allclose(oc.binary_auroc_exact(pm, tm), g[f"{name}/{tag}auroc_micro"], rtol=2e-6)
    np.testing.assert_allclose(oc.binary_average_precision_exact(pm, tm), g[f"{name}/{tag}ap_micro"], rtol=2e-6)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ign", [False, True])
def test_curves(g, name, ign):
    p, t = _inputs(g, name, ign)
    ig, tag = (-1, "ign_") if ign else (None, "")
    logits = name == "L6_logits"
    for l, ((fpr, tpr, thr), (pr, rc, th2)) in enumerate(zip(oc.multilabel_roc_ref32(p, t, ig), oc.multilabel_prc_ref32(p, t, ig))):
        np.testing.assert_allclose(fpr, g[f"{name}/{tag}roc_fpr{l}"], rtol=1e-6)
        np.testing.assert_allclose(tpr, g[f"{name}/{tag}roc_tpr{l}"], rtol=1e-6)
        np.testing.assert_allclose(thr, g[f"{name}/{tag}roc_thr{l}"], rtol=1e-6 if logits else 0)
        np.testing.assert_allclose(pr, g[f"{name}/{tag}prc_p{l}"], rtol=1e-6, equal_nan=True)
        np.testing.assert_allclose(rc, g[f"{name}/{tag}prc_r{l}"], rtol=1e-6, equal_nan=True)
        np.testing.assert_allclose(th2, g[f"{name}/{tag}prc_thr{l}"], rtol=1e-6 if logits else 0)


@pytest.mark.parametrize("name", ["L4_probs", "L3_ties", "L5_extra"])  # logits: a float sigmoid on a threshold may flip a bin
@pytest.mark.parametrize("ign", [False, True])
def test_binned_confmat(g, name, ign):
    p, t = _inputs(g, name, ign)
    tag = "ign_" if ign else ""
    for tname, thr in (("int9", np.linspace(0, 1, 9, dtype=np.float32)), ("list", np.array([0.8, 0.15, 0.5], np.float32))):
        np.testing.assert_array_equal(oc.multilabel_binned_confmat(p, t, thr), g[f"{name}/{tag}{tname}/confmat"])
