"""GPU: stat-score / confusion-matrix consumer metrics end to end (kernels K1b / K2 / K1 through the C-ABI + the host
reducers) vs goldens from the unmodified reference.  Tolerance 1e-6 relative (kappa / MCC: + 1e-6 absolute, their
formulas subtract nearly equal fp32 numbers)."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KINDS = ["precision", "recall", "specificity", "negative_predictive_value", "hamming_distance"]
AVGS = ["micro", "macro", "weighted", "none"]
C, L = 7, 5


def _fc():
    import metrics_b200.functional.classification as fc

    return fc


def _d(g, key):
    return torch.from_numpy(g[key]).to(DEV)


@pytest.mark.parametrize("kind", KINDS)
def test_binary_functionals(golden_consumers, kind):
    g, fc = golden_consumers, _fc()
    fn = getattr(fc, f"binary_{kind}")
    p = _d(g, "b/preds")
    np.testing.assert_allclose(fn(p, _d(g, "b/target_good")).cpu().numpy(), g[f"b/{kind}"], rtol=1e-6)
    np.testing.assert_allclose(fn(p, _d(g, "b/target_ign"), ignore_index=-1).cpu().numpy(), g[f"b/{kind}_ign"], rtol=1e-6)
    np.testing.assert_allclose(fn(p, _d(g, "b/target_good"), threshold=0.8).cpu().numpy(), g[f"b/{kind}_thr0.8"], rtol=1e-6)
    got = fn(_d(g, "b/preds_multi"), _d(g, "b/target_multi"), multidim_average="samplewise")
    np.testing.assert_allclose(got.cpu().numpy(), g[f"b/{kind}_samplewise"], rtol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("avg", AVGS)
def test_multiclass_functionals(golden_consumers, kind, avg):
    g, fc = golden_consumers, _fc()
    fn = getattr(fc, f"multiclass_{kind}")
    lg, t = _d(g, "mc/logits"), _d(g, "mc/target")
    np.testing.assert_allclose(fn(lg, t, C, average=avg).cpu().numpy(), g[f"mc/{kind}_{avg}"], rtol=1e-6)
    np.testing.assert_allclose(fn(lg, _d(g, "mc/target_ign"), C, average=avg, ignore_index=-1).cpu().numpy(), g[f"mc/{kind}_{avg}_ign"], rtol=1e-6)
    np.testing.assert_allclose(fn(lg, t, C, average=avg, top_k=2).cpu().numpy(), g[f"mc/{kind}_{avg}_top2"], rtol=1e-6)
    got = fn(_d(g, "mc/logits_multi"), _d(g, "mc/target_multi"), C, average=avg, multidim_average="samplewise")
    np.testing.assert_allclose(got.cpu().numpy(), g[f"mc/{kind}_{avg}_samplewise"], rtol=1e-6)
    if kind in ("precision", "recall", "negative_predictive_value"):
        np.testing.assert_allclose(fn(lg, t, C, average=avg, zero_division=1).cpu().numpy(), g[f"mc/{kind}_{avg}_zd1"], rtol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("avg", AVGS)
def test_multilabel_functionals(golden_consumers, kind, avg):
    g, fc = golden_consumers, _fc()
    fn = getattr(fc, f"multilabel_{kind}")
    p, t = _d(g, "ml/preds"), _d(g, "ml/target")
    np.testing.assert_allclose(fn(p, t, L, average=avg).cpu().numpy(), g[f"ml/{kind}_{avg}"], rtol=1e-6)
    np.testing.assert_allclose(fn(p, _d(g, "ml/target_ign"), L, average=avg, ignore_index=-1).cpu().numpy(), g[f"ml/{kind}_{avg}_ign"], rtol=1e-6)
    got = fn(_d(g, "ml/preds_multi"), _d(g, "ml/target_multi"), L, average=avg, multidim_average="samplewise")
    np.testing.assert_allclose(got.cpu().numpy(), g[f"ml/{kind}_{avg}_samplewise"], rtol=1e-6)
    if kind in ("precision", "recall", "negative_predictive_value"):
        np.testing.assert_allclose(fn(p, t, L, average=avg, zero_division=1).cpu().numpy(), g[f"ml/{kind}_{avg}_zd1"], rtol=1e-6)


def test_task_wrappers_functional(golden_consumers):
    g, fc = golden_consumers, _fc()
    got = fc.precision(_d(g, "mc/logits"), _d(g, "mc/target"), task="multiclass", num_classes=C, average="macro")
    np.testing.assert_allclose(got.cpu().numpy(), g["mc/precision_macro"], rtol=1e-6)
    got = fc.recall(_d(g, "ml/preds"), _d(g, "ml/target"), task="multilabel", num_labels=L, average="weighted")
    np.testing.assert_allclose(got.cpu().numpy(), g["ml/recall_weighted"], rtol=1e-6)
    got = fc.specificity(_d(g, "b/preds"), _d(g, "b/target_good"), task="binary")
    np.testing.assert_allclose(got.cpu().numpy(), g["b/specificity"], rtol=1e-6)


@pytest.mark.parametrize("avg", AVGS)
def test_jaccard(golden_consumers, avg):
    g, fc = golden_consumers, _fc()
    lg, t = _d(g, "mc/logits"), _d(g, "mc/target")
    np.testing.assert_allclose(fc.multiclass_jaccard_index(lg, t, C, average=avg).cpu().numpy(), g[f"mc/jaccard_{avg}"], rtol=1e-6)
    np.testing.assert_allclose(fc.multiclass_jaccard_index(lg, _d(g, "mc/target_ign"), C, average=avg, ignore_index=-1).cpu().numpy(), g[f"mc/jaccard_{avg}_ign"], rtol=1e-6)
    np.testing.assert_allclose(fc.multiclass_jaccard_index(lg, t, C, average=avg, ignore_index=2).cpu().numpy(), g[f"mc/jaccard_{avg}_ign2"], rtol=1e-6)
    np.testing.assert_allclose(fc.multiclass_jaccard_index(lg, t, C, average=avg, zero_division=1.0).cpu().numpy(), g[f"mc/jaccard_{avg}_zd1"], rtol=1e-6)
    p, tl = _d(g, "ml/preds"), _d(g, "ml/target")
    np.testing.assert_allclose(fc.multilabel_jaccard_index(p, tl, L, average=avg).cpu().numpy(), g[f"ml/jaccard_{avg}"], rtol=1e-6)
    np.testing.assert_allclose(fc.multilabel_jaccard_index(p, _d(g, "ml/target_ign"), L, average=avg, ignore_index=-1).cpu().numpy(), g[f"ml/jaccard_{avg}_ign"], rtol=1e-6)


def test_kappa_mcc_binary_jaccard(golden_consumers):
    g, fc = golden_consumers, _fc()
    tol = dict(rtol=1e-5, atol=1e-6)
    bp, bt, bg, bi = _d(g, "b/preds"), _d(g, "b/target"), _d(g, "b/target_good"), _d(g, "b/target_ign")
    lg, t, ti = _d(g, "mc/logits"), _d(g, "mc/target"), _d(g, "mc/target_ign")
    np.testing.assert_allclose(fc.binary_jaccard_index(bp, bg).cpu().numpy(), g["b/jaccard"], rtol=1e-6)
    np.testing.assert_allclose(fc.binary_jaccard_index(bp, bi, ignore_index=-1).cpu().numpy(), g["b/jaccard_ign"], rtol=1e-6)
    for w in ("none", "linear", "quadratic"):
        np.testing.assert_allclose(fc.binary_cohen_kappa(bp, bg, weights=w).cpu().numpy(), g[f"b/kappa_{w}"], **tol)
        np.testing.assert_allclose(fc.multiclass_cohen_kappa(lg, t, C, weights=w).cpu().numpy(), g[f"mc/kappa_{w}"], **tol)
        np.testing.assert_allclose(fc.multiclass_cohen_kappa(lg, ti, C, weights=w, ignore_index=-1).cpu().numpy(), g[f"mc/kappa_{w}_ign"], **tol)
    np.testing.assert_allclose(fc.binary_matthews_corrcoef(bp, bg).cpu().numpy(), g["b/mcc"], **tol)
    np.testing.assert_allclose(fc.binary_matthews_corrcoef(bp, bt).cpu().numpy(), g["b/mcc_rand"], **tol)
    np.testing.assert_allclose(fc.binary_matthews_corrcoef(bp, bi, ignore_index=-1).cpu().numpy(), g["b/mcc_ign"], **tol)
    np.testing.assert_allclose(fc.binary_matthews_corrcoef((bt > 0).float(), bt).cpu().numpy(), g["b/mcc_perfect"], **tol)
    np.testing.assert_allclose(fc.binary_matthews_corrcoef((bt == 0).float(), bt).cpu().numpy(), g["b/mcc_inverse"], **tol)
    np.testing.assert_allclose(fc.binary_matthews_corrcoef(torch.ones(700, device=DEV), bt).cpu().numpy(), g["b/mcc_allpos_pred"], **tol)
    np.testing.assert_allclose(fc.binary_matthews_corrcoef(bp, torch.zeros(700, dtype=torch.long, device=DEV)).cpu().numpy(), g["b/mcc_allneg_target"], **tol)
    np.testing.assert_allclose(fc.multiclass_matthews_corrcoef(lg, t, C).cpu().numpy(), g["mc/mcc"], **tol)
    np.testing.assert_allclose(fc.multiclass_matthews_corrcoef(lg, ti, C, ignore_index=-1).cpu().numpy(), g["mc/mcc_ign"], **tol)
    z = torch.zeros(50, dtype=torch.long, device=DEV)
    np.testing.assert_allclose(fc.multiclass_matthews_corrcoef(z, z, C).cpu().numpy(), g["mc/mcc_const"], **tol)
    p, tl = _d(g, "ml/preds"), _d(g, "ml/target")
    np.testing.assert_allclose(fc.multilabel_matthews_corrcoef(p, tl, L).cpu().numpy(), g["ml/mcc"], **tol)
    np.testing.assert_allclose(fc.multilabel_matthews_corrcoef(p, _d(g, "ml/target_ign"), L, ignore_index=-1).cpu().numpy(), g["ml/mcc_ign"], **tol)


def test_modular_classes(golden_consumers):
    import metrics_b200.classification as TC
    from metrics_b200 import MetricCollection

    g = golden_consumers
    tol = dict(rtol=1e-5, atol=1e-6)
    lg, t = _d(g, "mc/logits"), _d(g, "mc/target")
    mods = {
        "MulticlassPrecision": TC.MulticlassPrecision(num_classes=C, average="macro"),
        "MulticlassRecall_top2": TC.MulticlassRecall(num_classes=C, average="weighted", top_k=2),
        "MulticlassSpecificity": TC.MulticlassSpecificity(num_classes=C, average="none"),
        "MulticlassHammingDistance": TC.MulticlassHammingDistance(num_classes=C, average="micro"),
        "MulticlassJaccardIndex": TC.MulticlassJaccardIndex(num_classes=C),
        "MulticlassCohenKappa": TC.MulticlassCohenKappa(num_classes=C, weights="linear"),
        "MulticlassMatthewsCorrCoef": TC.MulticlassMatthewsCorrCoef(num_classes=C),
    }
    for name, m in mods.items():
        m = m.to(DEV)
        for a, b in zip(lg.chunk(3), t.chunk(3)):
            m.update(a, b)
        np.testing.assert_allclose(m.compute().cpu().numpy(), g[f"class/{name}"], err_msg=name, **tol)
    p, tl = _d(g, "ml/preds"), _d(g, "ml/target")
    mods = {
        "MultilabelPrecision": TC.MultilabelPrecision(num_labels=L, average="macro"),
        "MultilabelNegativePredictiveValue": TC.MultilabelNegativePredictiveValue(num_labels=L, average="none"),
        "MultilabelJaccardIndex": TC.MultilabelJaccardIndex(num_labels=L, average="weighted"),
        "MultilabelMatthewsCorrCoef": TC.MultilabelMatthewsCorrCoef(num_labels=L),
    }
    for name, m in mods.items():
        m = m.to(DEV)
        for a, b in zip(p.chunk(3), tl.chunk(3)):
            m.update(a, b)
        np.testing.assert_allclose(m.compute().cpu().numpy(), g[f"class/{name}"], err_msg=name, **tol)
    bp, bg = _d(g, "b/preds"), _d(g, "b/target_good")
    mods = {"BinaryRecall": TC.BinaryRecall(), "BinaryCohenKappa": TC.BinaryCohenKappa(), "BinaryJaccardIndex": TC.BinaryJaccardIndex(),
            "BinaryMatthewsCorrCoef": TC.BinaryMatthewsCorrCoef(), "BinaryHammingDistance": TC.BinaryHammingDistance()}
    for name, m in mods.items():
        m = m.to(DEV)
        for a, b in zip(bp.chunk(3), bg.chunk(3)):
            m.update(a, b)
        np.testing.assert_allclose(m.compute().cpu().numpy(), g[f"class/{name}"], err_msg=name, **tol)
    # stat-score consumers with identical constructor args fall into ONE compute group: one kernel per update for all
    mc = MetricCollection([TC.MulticlassPrecision(num_classes=C), TC.MulticlassRecall(num_classes=C),
                           TC.MulticlassSpecificity(num_classes=C), TC.MulticlassF1Score(num_classes=C)]).to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mc.update(lg, t)
        mc.update(lg, t)
    assert len(mc.compute_groups) == 1
    res = mc.compute()
    np.testing.assert_allclose(res["MulticlassPrecision"].cpu().numpy(), g["mc/precision_macro"], rtol=1e-6)
    np.testing.assert_allclose(res["MulticlassRecall"].cpu().numpy(), g["mc/recall_macro"], rtol=1e-6)
    # task wrappers
    assert isinstance(TC.Precision(task="multiclass", num_classes=C), TC.MulticlassPrecision)
    assert isinstance(TC.JaccardIndex(task="multilabel", num_labels=L), TC.MultilabelJaccardIndex)
    assert isinstance(TC.CohenKappa(task="binary"), TC.BinaryCohenKappa)
    assert isinstance(TC.MatthewsCorrCoef(task="multiclass", num_classes=C), TC.MulticlassMatthewsCorrCoef)
