"""CPU: host-side reducers of the stat-score / confusion-matrix consumer metrics (no kernel involved: counters come from
the numpy oracle) against goldens produced by the unmodified reference (tests/golden/make_golden.py consumers)."""
import numpy as np
import pytest
import torch

from metrics_b200.functional.classification.confmat_metrics import (
    _cohen_kappa_reduce,
    _jaccard_index_reduce,
    _matthews_corrcoef_reduce,
)
from metrics_b200.functional.classification.ratio_metrics import _ratio_reduce
from oracle import classification as ocl

KINDS = ["precision", "recall", "specificity", "negative_predictive_value", "hamming_distance"]
AVGS = ["micro", "macro", "weighted", "none"]
C, L = 7, 5


def _t(*arrs):
    return [torch.as_tensor(np.asarray(a)) for a in arrs]


@pytest.mark.parametrize("kind", KINDS)
def test_binary(golden_consumers, kind):
    g = golden_consumers
    for tag, tkey, ig, thr in (("", "b/target_good", None, 0.5), ("_ign", "b/target_ign", -1, 0.5), ("_thr0.8", "b/target_good", None, 0.8)):
        st = _t(*ocl.binary_stat_scores(g["b/preds"], g[tkey], thr, ig))
        got = _ratio_reduce(kind, *st, "binary")
        np.testing.assert_allclose(got.numpy(), g[f"b/{kind}{tag}"], rtol=1e-6)
    st = _t(*ocl.binary_stat_scores(g["b/preds_multi"], g["b/target_multi"], samplewise=True))
    got = _ratio_reduce(kind, *st, "binary", "samplewise")
    np.testing.assert_allclose(got.numpy(), g[f"b/{kind}_samplewise"], rtol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("avg", AVGS)
def test_multiclass_and_multilabel(golden_consumers, kind, avg):
    g = golden_consumers
    for tag, tkey, ig in (("", "mc/target", None), ("_ign", "mc/target_ign", -1)):
        st = _t(*ocl.multiclass_stat_scores(g["mc/logits"], g[tkey], C, avg, ig))
        got = _ratio_reduce(kind, *st, avg)
        np.testing.assert_allclose(got.numpy(), g[f"mc/{kind}_{avg}{tag}"], rtol=1e-6)
    if kind in ("precision", "recall", "negative_predictive_value"):
        st = _t(*ocl.multiclass_stat_scores(g["mc/logits"], g["mc/target"], C, avg))
        got = _ratio_reduce(kind, *st, avg, zero_division=1)
        np.testing.assert_allclose(got.numpy(), g[f"mc/{kind}_{avg}_zd1"], rtol=1e-6)
    for tag, tkey, ig in (("", "ml/target", None), ("_ign", "ml/target_ign", -1)):
        st = _t(*ocl.multilabel_stat_scores(g["ml/preds"], g[tkey], L, 0.5, ig))
        got = _ratio_reduce(kind, *st, avg, multilabel=True)
        np.testing.assert_allclose(got.numpy(), g[f"ml/{kind}_{avg}{tag}"], rtol=1e-6)
    st = _t(*ocl.multilabel_stat_scores(g["ml/preds_multi"], g["ml/target_multi"], L, samplewise=True))
    got = _ratio_reduce(kind, *st, avg, "samplewise", multilabel=True)
    np.testing.assert_allclose(got.numpy(), g[f"ml/{kind}_{avg}_samplewise"], rtol=1e-6)


def _mc_confmat(g, tkey, ig=None):
    return torch.from_numpy(ocl.multiclass_confusion_matrix(g["mc/logits"], g[tkey], C, ig))


def _ml_confmat(g, tkey, ig=None):
    return torch.from_numpy(ocl.confmat_from_counts(*ocl.multilabel_stat_scores(g["ml/preds"], g[tkey], L, 0.5, ig)))


def _b_confmat(g, preds, target, ig=None):
    return torch.from_numpy(ocl.confmat_from_counts(*ocl.binary_stat_scores(preds, target, 0.5, ig)))


@pytest.mark.parametrize("avg", AVGS)
def test_jaccard(golden_consumers, avg):
    g = golden_consumers
    np.testing.assert_allclose(_jaccard_index_reduce(_mc_confmat(g, "mc/target"), avg).numpy(), g[f"mc/jaccard_{avg}"], rtol=1e-6)
    np.testing.assert_allclose(_jaccard_index_reduce(_mc_confmat(g, "mc/target_ign", -1), avg, -1).numpy(), g[f"mc/jaccard_{avg}_ign"], rtol=1e-6)
    np.testing.assert_allclose(_jaccard_index_reduce(_mc_confmat(g, "mc/target", 2), avg, 2).numpy(), g[f"mc/jaccard_{avg}_ign2"], rtol=1e-6)
    np.testing.assert_allclose(_jaccard_index_reduce(_mc_confmat(g, "mc/target"), avg, zero_division=1.0).numpy(), g[f"mc/jaccard_{avg}_zd1"], rtol=1e-6)
    np.testing.assert_allclose(_jaccard_index_reduce(_ml_confmat(g, "ml/target"), avg).numpy(), g[f"ml/jaccard_{avg}"], rtol=1e-6)
    np.testing.assert_allclose(_jaccard_index_reduce(_ml_confmat(g, "ml/target_ign", -1), avg, -1).numpy(), g[f"ml/jaccard_{avg}_ign"], rtol=1e-6)


def test_binary_jaccard_kappa_mcc(golden_consumers):
    g = golden_consumers
    cm = _b_confmat(g, g["b/preds"], g["b/target_good"])
    cmi = _b_confmat(g, g["b/preds"], g["b/target_ign"], -1)
    np.testing.assert_allclose(_jaccard_index_reduce(cm, "binary").numpy(), g["b/jaccard"], rtol=1e-6)
    np.testing.assert_allclose(_jaccard_index_reduce(cmi, "binary").numpy(), g["b/jaccard_ign"], rtol=1e-6)
    for w in ("none", "linear", "quadratic"):
        np.testing.assert_allclose(_cohen_kappa_reduce(cm, w).numpy(), g[f"b/kappa_{w}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_cohen_kappa_reduce(_mc_confmat(g, "mc/target"), w).numpy(), g[f"mc/kappa_{w}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_cohen_kappa_reduce(_mc_confmat(g, "mc/target_ign", -1), w).numpy(), g[f"mc/kappa_{w}_ign"], rtol=1e-5, atol=1e-6)
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(cm).numpy(), g["b/mcc"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_b_confmat(g, g["b/preds"], g["b/target"])).numpy(), g["b/mcc_rand"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(cmi).numpy(), g["b/mcc_ign"], **tol)
    t = g["b/target"]
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_b_confmat(g, (t > 0).astype(np.float32), t)).numpy(), g["b/mcc_perfect"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_b_confmat(g, (t == 0).astype(np.float32), t)).numpy(), g["b/mcc_inverse"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_b_confmat(g, np.ones(700, np.float32), t)).numpy(), g["b/mcc_allpos_pred"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_b_confmat(g, g["b/preds"], np.zeros(700, np.int64))).numpy(), g["b/mcc_allneg_target"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_mc_confmat(g, "mc/target")).numpy(), g["mc/mcc"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_mc_confmat(g, "mc/target_ign", -1)).numpy(), g["mc/mcc_ign"], **tol)
    const = torch.zeros(C, C, dtype=torch.long)
    const[0, 0] = 50
    np.testing.assert_allclose(_matthews_corrcoef_reduce(const).numpy(), g["mc/mcc_const"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_ml_confmat(g, "ml/target")).numpy(), g["ml/mcc"], **tol)
    np.testing.assert_allclose(_matthews_corrcoef_reduce(_ml_confmat(g, "ml/target_ign", -1)).numpy(), g["ml/mcc_ign"], **tol)
