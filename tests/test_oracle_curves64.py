"""CPU: the numpy oracle on float64 scores and on `sample_weights` against goldens produced by the unmodified reference
(tests/golden/curves64.npz).  The float64 cases contain score pairs that differ only below float32 resolution: an
implementation that compares them as float32 produces fewer thresholds and fails on the lengths."""
import numpy as np
import pytest

from oracle import curves as oc
from tests import curves64_cases as cc


@pytest.mark.parametrize("k", range(cc.n_cases()))
def test_float64_binary_cases(k):
    g = cc.load()
    fn = str(g[f"case{k}/fn"])
    if fn not in ("binary_roc", "binary_precision_recall_curve", "binary_auroc", "binary_average_precision"):
        pytest.skip("multiclass / multilabel float64 cases are checked through the kernels (same scan, more segments)")
    preds, target = g[f"case{k}/preds"], g[f"case{k}/target"]
    assert preds.dtype == np.float64
    p = oc.sigmoid_if_logits(preds) if ((preds < 0).any() or (preds > 1).any()) else preds
    if p is not preds:  # float64 logits: the oracle's sigmoid helper is float32 — redo it in float64 like the reference
        p = 1.0 / (1.0 + np.exp(-preds))
    if fn == "binary_roc":
        got = oc.binary_roc_ref32(p, target)
    elif fn == "binary_precision_recall_curve":
        got = oc.binary_prc_ref32(p, target)
    elif fn == "binary_auroc":
        got = [np.array(oc.binary_auroc_exact(p, target))]
    else:
        got = [np.array(oc.binary_average_precision_exact(p, target))]
    assert len(got) == int(g[f"case{k}/n_out"])
    for i, arr in enumerate(got):
        exp = g[f"case{k}/out{i}"]
        assert np.shape(arr) == exp.shape, (fn, i, np.shape(arr), exp.shape)
        np.testing.assert_allclose(np.asarray(arr, dtype=np.float64), exp, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("k", range(cc.n_weighted()))
def test_weighted_clf_curve(k):
    g = cc.load()
    preds, target, w, pos = g[f"w{k}/preds"], g[f"w{k}/target"], g[f"w{k}/weights"], int(g[f"w{k}/pos"])
    fps, tps, thr = oc.binary_clf_curve(preds, target, pos, sample_weights=w)
    assert fps.shape == g[f"w{k}/fps"].shape
    np.testing.assert_allclose(fps, g[f"w{k}/fps"], rtol=1e-5, atol=1e-5)  # the reference accumulates in float32 for f32 weights
    np.testing.assert_allclose(tps, g[f"w{k}/tps"], rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(thr.astype(np.float64), g[f"w{k}/thr"].astype(np.float64))
