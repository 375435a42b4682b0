"""GPU: multilabel curve family (one batched sort + scan for all labels, K4 binned kernel in multilabel mode) through the
C-ABI vs goldens from the unmodified reference.  Integer states (binned confmat) bit-exact; floats within 1e-6 relative."""
import warnings

import numpy as np
import pytest
import torch

from oracle import curves as oc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["L4_probs", "L6_logits", "L3_ties", "L5_extra"]
RTOL = 1e-6


def _fc():
    import metrics_b200.functional.classification as fc

    return fc


def _load(g, name, ign):
    p = torch.from_numpy(g[f"{name}/preds"]).to(DEV)
    t = torch.from_numpy(g[f"{name}/target_ignore" if ign else f"{name}/target"]).to(DEV)
    return p, t, p.shape[1]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ign", [False, True])
def test_scalars_vs_golden(golden_multilabel, name, ign):
    fc, g = _fc(), golden_multilabel
    p, t, L = _load(g, name, ign)
    ig, tag = (-1, "ign_") if ign else (None, "")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for avg in ("micro", "macro", "weighted", "none"):
            a = fc.multilabel_auroc(p, t, L, average=avg, ignore_index=ig)
            b = fc.multilabel_average_precision(p, t, L, average=avg, ignore_index=ig)
            assert a.dtype == torch.float32 and b.dtype == torch.float32
            np.testing.assert_allclose(a.cpu().numpy(), g[f"{name}/{tag}auroc_{avg}"], rtol=RTOL, atol=1e-7)
            np.testing.assert_allclose(b.cpu().numpy(), g[f"{name}/{tag}ap_{avg}"], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ign", [False, True])
def test_curves_vs_golden(golden_multilabel, name, ign):
    fc, g = _fc(), golden_multilabel
    p, t, L = _load(g, name, ign)
    ig, tag = (-1, "ign_") if ign else (None, "")
    thr_tol = RTOL if name == "L6_logits" else 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fpr, tpr, thr = fc.multilabel_roc(p, t, L, ignore_index=ig)
        pr, rc, th2 = fc.multilabel_precision_recall_curve(p, t, L, ignore_index=ig)
    assert len(fpr) == L and len(pr) == L
    for l in range(L):
        np.testing.assert_allclose(fpr[l].cpu().numpy(), g[f"{name}/{tag}roc_fpr{l}"], rtol=RTOL)
        np.testing.assert_allclose(tpr[l].cpu().numpy(), g[f"{name}/{tag}roc_tpr{l}"], rtol=RTOL)
        np.testing.assert_allclose(thr[l].cpu().numpy(), g[f"{name}/{tag}roc_thr{l}"], rtol=thr_tol)
        np.testing.assert_allclose(pr[l].cpu().numpy(), g[f"{name}/{tag}prc_p{l}"], rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(rc[l].cpu().numpy(), g[f"{name}/{tag}prc_r{l}"], rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(th2[l].cpu().numpy(), g[f"{name}/{tag}prc_thr{l}"], rtol=thr_tol)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ign", [False, True])
def test_counts_vs_oracle_bit_exact(golden_multilabel, name, ign):
    """fps / tps of every label against the oracle's integer curve (kept entries only)."""
    from metrics_b200 import _native

    g = golden_multilabel
    p, t, L = _load(g, name, ign)
    pn, tn = oc.multilabel_flatten(p.cpu().numpy(), t.cpu().numpy())
    pd = _native.sigmoid_if_logits(torch.from_numpy(np.ascontiguousarray(pn)).to(DEV))
    td = torch.from_numpy(np.ascontiguousarray(tn)).to(DEV)
    _, _, counts, (fps, tps, thr) = _native.curve_evaluate_multilabel(pd, td, L, -1 if ign else None, want_curve=True)
    pn = pd.cpu().numpy()
    for l in range(L):
        pl, tl = oc._label_column(pn, tn, l, -1 if ign else None)
        efps, etps, ethr = oc.binary_clf_curve(pl, tl)
        u = int(counts[l, 2])
        assert u == efps.size
        assert int(counts[l, 0]) == int((tl == 1).sum()) and int(counts[l, 1]) == int((tl != 1).sum())
        np.testing.assert_array_equal(fps[l, :u].cpu().numpy().astype(np.int64), efps)
        np.testing.assert_array_equal(tps[l, :u].cpu().numpy().astype(np.int64), etps)
        np.testing.assert_array_equal(thr[l, :u].cpu().numpy(), ethr)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ign", [False, True])
@pytest.mark.parametrize("tname", ["int9", "list"])
def test_binned_vs_golden(golden_multilabel, name, ign, tname):
    from metrics_b200.functional.classification.precision_recall_curve import (
        _multilabel_precision_recall_curve_format,
        _multilabel_precision_recall_curve_update,
    )

    fc, g = _fc(), golden_multilabel
    p, t, L = _load(g, name, ign)
    ig, tag = (-1, "ign_") if ign else (None, "")
    thrs = 9 if tname == "int9" else [0.8, 0.15, 0.5]
    pf, tf, th = _multilabel_precision_recall_curve_format(p, t, L, thrs, ig)
    cm = _multilabel_precision_recall_curve_update(pf, tf, L, th)
    assert cm.dtype == torch.int64 and tuple(cm.shape) == (len(th), L, 2, 2)
    if name != "L6_logits":  # a sigmoid output within 1 ulp of a threshold may land in the other bin
        np.testing.assert_array_equal(cm.cpu().numpy(), g[f"{name}/{tag}{tname}/confmat"])
    else:
        assert np.abs(cm.cpu().numpy() - g[f"{name}/{tag}{tname}/confmat"]).max() <= 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for avg in ("micro", "macro", "weighted", "none"):
            a = fc.multilabel_auroc(p, t, L, average=avg, thresholds=thrs, ignore_index=ig)
            b = fc.multilabel_average_precision(p, t, L, average=avg, thresholds=thrs, ignore_index=ig)
            np.testing.assert_allclose(a.cpu().numpy(), g[f"{name}/{tag}{tname}/auroc_{avg}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(b.cpu().numpy(), g[f"{name}/{tag}{tname}/ap_{avg}"], rtol=2e-6, atol=1e-7, equal_nan=True)
        f, tp_, h = fc.multilabel_roc(p, t, L, thresholds=thrs, ignore_index=ig)
        np.testing.assert_allclose(f.cpu().numpy(), g[f"{name}/{tag}{tname}/roc_fpr"], rtol=RTOL)
        np.testing.assert_allclose(tp_.cpu().numpy(), g[f"{name}/{tag}{tname}/roc_tpr"], rtol=RTOL)
        np.testing.assert_allclose(h.cpu().numpy(), g[f"{name}/{tag}{tname}/roc_thr"], rtol=0)
        pr, rc, _ = fc.multilabel_precision_recall_curve(p, t, L, thresholds=thrs, ignore_index=ig)
        np.testing.assert_allclose(pr.cpu().numpy(), g[f"{name}/{tag}{tname}/prc_p"], rtol=RTOL)
        np.testing.assert_allclose(rc.cpu().numpy(), g[f"{name}/{tag}{tname}/prc_r"], rtol=RTOL)


def test_modular_classes_and_shared_evaluation(golden_multilabel):
    from metrics_b200 import MetricCollection, _native
    from metrics_b200.classification import AUROC, AveragePrecision, MultilabelAUROC, MultilabelAveragePrecision

    g = golden_multilabel
    p = torch.from_numpy(g["L6_logits/preds"]).to(DEV)
    t = torch.from_numpy(g["L6_logits/target"]).to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for avg in ("macro", "micro"):
            mc = MetricCollection([MultilabelAUROC(num_labels=6, average=avg), MultilabelAveragePrecision(num_labels=6, average=avg)]).to(DEV)
            m3 = MultilabelAUROC(num_labels=6, average=avg, thresholds=25).to(DEV)
            for a, b in zip(p.chunk(3), t.chunk(3)):
                mc.update(a, b)
                m3.update(a, b)
            before = _native.launch_count()
            res = mc.compute()
            launches = _native.launch_count() - before
            np.testing.assert_allclose(res["MultilabelAUROC"].cpu().numpy(), g[f"class/auroc_{avg}"], rtol=RTOL)
            np.testing.assert_allclose(res["MultilabelAveragePrecision"].cpu().numpy(), g[f"class/ap_{avg}"], rtol=RTOL)
            np.testing.assert_allclose(m3.compute().cpu().numpy(), g[f"class/auroc_binned25_{avg}"], rtol=2e-6)
            if avg == "macro":
                assert len(mc.compute_groups) == 1
                assert launches <= 16, launches  # ONE pack + sort + scan for both metrics of the group
        assert isinstance(AUROC(task="multilabel", num_labels=6), MultilabelAUROC)
        assert isinstance(AveragePrecision(task="multilabel", num_labels=6), MultilabelAveragePrecision)
