"""GPU: float64 scores (64-bit sort keys, 8 radix passes) and `sample_weights` through the exact curve kernels against
goldens produced by the unmodified reference (tests/golden/curves64.npz).  The float64 cases hold score pairs that differ only
below float32 resolution, so the number of thresholds is itself the test that nothing was down-cast."""
import warnings

import numpy as np
import pytest
import torch

from tests import curves64_cases as cc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _flat(res):
    parts = []
    for part in (res if isinstance(res, (tuple, list)) else [res]):
        parts.extend(part if isinstance(part, (tuple, list)) else [part])
    return parts


@pytest.mark.parametrize("k", range(cc.n_cases()))
def test_float64_scores(k):
    import metrics_b200.functional.classification as F

    g = cc.load()
    fn, c = str(g[f"case{k}/fn"]), int(g[f"case{k}/num_classes"])
    preds = torch.from_numpy(g[f"case{k}/preds"]).to(DEV)
    target = torch.from_numpy(g[f"case{k}/target"]).to(DEV)
    assert preds.dtype == torch.float64
    kw = dict(thresholds=None)
    if fn.startswith("multiclass"):
        kw["num_classes"] = c
    if fn.startswith("multilabel"):
        kw["num_labels"] = c
    if not fn.startswith("binary") and ("auroc" in fn or "average_precision" in fn):
        kw["average"] = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = _flat(getattr(F, fn)(preds, target, **kw))
    assert len(got) == int(g[f"case{k}/n_out"])
    for i, t in enumerate(got):
        exp = g[f"case{k}/out{i}"]
        assert tuple(t.shape) == exp.shape, f"{fn} out{i}: {tuple(t.shape)} vs {exp.shape}"
        if exp.dtype == np.float64:  # thresholds: the scores themselves, bit for bit (sigmoid / softmax in float64: 1e-15)
            assert t.dtype == torch.float64
            np.testing.assert_allclose(t.cpu().numpy(), exp, rtol=1e-14, atol=1e-15, equal_nan=True, err_msg=f"{fn} out{i}")
        else:
            np.testing.assert_allclose(t.double().cpu().numpy(), exp, rtol=1e-6, atol=1e-7, equal_nan=True, err_msg=f"{fn} out{i}")


@pytest.mark.parametrize("k", range(cc.n_weighted()))
def test_sample_weights(k):
    from metrics_b200.functional.classification.precision_recall_curve import _binary_clf_curve

    g = cc.load()
    half = bool(g[f"w{k}/half"])
    preds = torch.from_numpy(g[f"w{k}/preds"]).to(DEV)
    preds = preds.half() if half else preds
    target = torch.from_numpy(g[f"w{k}/target"]).to(DEV)
    weights = torch.from_numpy(g[f"w{k}/weights"]).to(DEV)
    fps, tps, thr = _binary_clf_curve(preds, target, sample_weights=weights, pos_label=int(g[f"w{k}/pos"]))
    assert fps.dtype == weights.dtype and tps.dtype == weights.dtype and thr.dtype == preds.dtype
    assert tuple(fps.shape) == g[f"w{k}/fps"].shape
    # the reference accumulates float32 weights in float32 (sequentially on CPU): ~1e-6 relative per prefix sum
    np.testing.assert_allclose(fps.double().cpu().numpy(), g[f"w{k}/fps"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(tps.double().cpu().numpy(), g[f"w{k}/tps"], rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(thr.double().cpu().numpy(), g[f"w{k}/thr"].astype(np.float64))


def test_weights_as_python_list():
    from metrics_b200.functional.classification.precision_recall_curve import _binary_clf_curve

    g = cc.load()
    preds = torch.tensor([0.1, 0.4, 0.35, 0.8, 0.4], device=DEV)
    target = torch.tensor([0, 0, 1, 1, 1], device=DEV)
    fps, tps, thr = _binary_clf_curve(preds, target, sample_weights=[1.0, 2.0, 0.5, 1.5, 1.0])
    assert fps.dtype == torch.float32
    np.testing.assert_allclose(fps.cpu().numpy(), g["wlist/fps"], rtol=1e-6)
    np.testing.assert_allclose(tps.cpu().numpy(), g["wlist/tps"], rtol=1e-6)
    np.testing.assert_array_equal(thr.cpu().numpy(), g["wlist/thr"])
