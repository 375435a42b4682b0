"""GPU: K4 binned curve update + binned compute paths vs reference goldens.  Integer confusion matrices bit-exact for
untransformed scores (<= 2 boundary flips where a sigmoid/softmax was applied: ATen CPU vs fp32 expf); float outputs 1e-6."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
THR = {"int11": 11, "int200": 200, "list": [0.9, 0.1, 0.5, 0.3], "tensor": torch.tensor([0.0, 0.2, 0.7, 1.0])}
TOL = dict(rtol=1e-6, atol=1e-7)


def _fc():
    import metrics_b200.functional.classification as fc

    return fc


@pytest.mark.parametrize("name", list(THR))
@pytest.mark.parametrize("kind", ["probs", "logits"])
def test_binary_binned_vs_golden(golden_binned, name, kind):
    from metrics_b200.functional.classification.precision_recall_curve import (
        _binary_precision_recall_curve_format,
        _binary_precision_recall_curve_update,
    )

    fc, g = _fc(), golden_binned
    p = torch.from_numpy(g["b/preds"] if kind == "probs" else g["b/logits"]).to(DEV)
    t = torch.from_numpy(g["b/target"]).to(DEV)[: p.numel()]
    thr = THR[name]
    tag = f"b/{name}/{kind}"
    pf, tf, th = _binary_precision_recall_curve_format(p, t, thr)
    cm = _binary_precision_recall_curve_update(pf, tf, th)
    assert cm.dtype == torch.int64
    if kind == "probs":
        np.testing.assert_array_equal(cm.cpu().numpy(), g[f"{tag}/confmat"])
        tol = TOL
    else:
        assert np.abs(cm.cpu().numpy() - g[f"{tag}/confmat"]).max() <= 2
        tol = dict(rtol=2e-3, atol=1e-3)  # a flipped boundary sample moves a rate by 1/N
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_allclose(fc.binary_auroc(p, t, thresholds=thr).cpu().numpy(), g[f"{tag}/auroc"], **tol)
        np.testing.assert_allclose(fc.binary_auroc(p, t, thresholds=thr, max_fpr=0.6).cpu().numpy(), g[f"{tag}/auroc_maxfpr"], **tol)
        np.testing.assert_allclose(fc.binary_average_precision(p, t, thresholds=thr).cpu().numpy(), g[f"{tag}/ap"], **tol)
        f, tp_, h = fc.binary_roc(p, t, thresholds=thr)
        np.testing.assert_allclose(f.cpu().numpy(), g[f"{tag}/roc_fpr"], **tol)
        np.testing.assert_allclose(tp_.cpu().numpy(), g[f"{tag}/roc_tpr"], **tol)
        np.testing.assert_array_equal(h.cpu().numpy(), g[f"{tag}/roc_thr"])
        pr, rc, h = fc.binary_precision_recall_curve(p, t, thresholds=thr)
        np.testing.assert_allclose(pr.cpu().numpy(), g[f"{tag}/prc_p"], **tol)
        np.testing.assert_allclose(rc.cpu().numpy(), g[f"{tag}/prc_r"], **tol)
        np.testing.assert_array_equal(h.cpu().numpy(), g[f"{tag}/prc_thr"])


@pytest.mark.parametrize("name,thr", [("int7", 7), ("list", [0.05, 0.2, 0.6])])
def test_multiclass_binned_vs_golden(golden_binned, name, thr):
    from metrics_b200.functional.classification.precision_recall_curve import (
        _multiclass_precision_recall_curve_format,
        _multiclass_precision_recall_curve_update,
    )

    fc, g = _fc(), golden_binned
    p = torch.from_numpy(g["m/logits"]).to(DEV)
    t = torch.from_numpy(g["m/target"]).to(DEV)
    pf, tf, th = _multiclass_precision_recall_curve_format(p, t, 6, thr)
    cm = _multiclass_precision_recall_curve_update(pf, tf, 6, th)
    assert np.abs(cm.cpu().numpy() - g[f"m/{name}/confmat"]).max() <= 2
    tol = dict(rtol=3e-3, atol=2e-3, equal_nan=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for avg in ("macro", "weighted", "none"):
            np.testing.assert_allclose(fc.multiclass_auroc(p, t, 6, average=avg, thresholds=thr).cpu().numpy(), g[f"m/{name}/auroc_{avg}"], **tol)
            np.testing.assert_allclose(fc.multiclass_average_precision(p, t, 6, average=avg, thresholds=thr).cpu().numpy(), g[f"m/{name}/ap_{avg}"], **tol)
        f, tp_, h = fc.multiclass_roc(p, t, 6, thresholds=thr)
        np.testing.assert_allclose(f.cpu().numpy(), g[f"m/{name}/roc_fpr"], **tol)
        np.testing.assert_allclose(tp_.cpu().numpy(), g[f"m/{name}/roc_tpr"], **tol)
        pr, rc, h = fc.multiclass_precision_recall_curve(p, t, 6, thresholds=thr)
        np.testing.assert_allclose(pr.cpu().numpy(), g[f"m/{name}/prc_p"], **tol)
        np.testing.assert_allclose(rc.cpu().numpy(), g[f"m/{name}/prc_r"], **tol)
        for avg in ("micro", "macro"):
            pr, rc, h = fc.multiclass_precision_recall_curve(p, t, 6, thresholds=thr, average=avg)
            np.testing.assert_allclose(pr.cpu().numpy(), g[f"m/{name}/prc_{avg}_p"], **tol)
            got, ref = rc.cpu().numpy(), g[f"m/{name}/prc_{avg}_r"]
            if avg == "macro":
                # the reference's `interp` divides by zero on repeated precision values (inf/NaN garbage that depends
                # on the last ulp): compare only where both sides are finite and of sane magnitude
                ok = np.isfinite(got) & np.isfinite(ref) & (np.abs(ref) <= 1.0)
                got, ref = got[ok], ref[ok]
            np.testing.assert_allclose(got, ref, **tol)


def test_binned_classes_and_large_threshold_counts(golden_binned):
    from metrics_b200.classification import BinaryAUROC, MulticlassAveragePrecision
    from oracle import curves as oc

    g = golden_binned
    bp, bt = torch.from_numpy(g["b/preds"]).to(DEV), torch.from_numpy(g["b/target"]).to(DEV)
    m = BinaryAUROC(thresholds=50).to(DEV)
    for a, b in zip(bp.chunk(5), bt.chunk(5)):
        m.update(a, b)
    np.testing.assert_array_equal(m.confmat.cpu().numpy(), g["class/binary_auroc_50_confmat"])
    np.testing.assert_allclose(m.compute().cpu().numpy(), g["class/binary_auroc_50"], **TOL)
    assert "confmat" in m.metric_state and m.thresholds.device.type == torch.device(DEV).type
    ml, mt = torch.from_numpy(g["m/logits"]).to(DEV), torch.from_numpy(g["m/target"]).to(DEV)
    m2 = MulticlassAveragePrecision(num_classes=6, thresholds=20).to(DEV)
    for a, b in zip(ml.chunk(3), mt.chunk(3)):
        m2.update(a, b)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_allclose(m2.compute().cpu().numpy(), g["class/mc_ap_20"], rtol=3e-3, atol=2e-3)
    # many thresholds (global-atomic path) and many classes vs the oracle, untransformed probabilities
    from metrics_b200 import _native

    gen = torch.Generator().manual_seed(2)
    p = torch.rand(20000, generator=gen)
    t = torch.randint(0, 2, (20000,), generator=gen)
    thr = torch.linspace(0, 1, 9000)
    np.testing.assert_array_equal(_native.binned_curve_update(p.to(DEV), t.to(DEV), thr.to(DEV), 1).cpu().numpy(),
                                  oc.binned_confmat(p.numpy(), t.numpy(), thr.numpy()))
    p = torch.rand(3000, 40, generator=gen)
    t = torch.randint(0, 40, (3000,), generator=gen)
    thr = torch.linspace(0, 1, 33)
    np.testing.assert_array_equal(_native.binned_curve_update(p.to(DEV), t.to(DEV), thr.to(DEV), 40).cpu().numpy(),
                                  oc.binned_confmat(p.numpy(), t.numpy(), thr.numpy(), 40))


@pytest.mark.parametrize("label_dtype", [torch.int64, torch.int32, torch.uint8, torch.bool, torch.int8])
@pytest.mark.parametrize("thr_kind", ["linspace200", "linspace3", "irregular", "single", "dense", "equal"])
def test_binary_fast_path_equals_generic_kernel_and_oracle(label_dtype, thr_kind):
    """The binary fast path of K4 (float32 scores, 16-byte aligned, n >= 4096: vector loads, branch-free bucket search) against
    the generic kernel (same data, forced by a 4-byte misaligned view) and the numpy oracle: integer confusion matrices,
    bit-exact — including scores sitting exactly on thresholds, NaN, +-inf, and targets outside {0, 1}."""
    from metrics_b200 import _native
    from oracle import curves as oc

    n = 100_003
    g = torch.Generator().manual_seed(17)
    thr = {"linspace200": torch.linspace(0, 1, 200), "linspace3": torch.linspace(0, 1, 3),
           "irregular": torch.tensor([0.01, 0.011, 0.2, 0.5, 0.50001, 0.9, 0.97, 0.99]), "single": torch.tensor([0.5]),
           "dense": torch.linspace(0.4, 0.6, 2000), "equal": torch.full((5,), 0.25)}[thr_kind]
    p = torch.rand(n + 1, generator=g)
    p[1:4001] = thr[torch.randint(0, thr.numel(), (4000,), generator=g)]  # exactly on a threshold
    p[5000:5004] = torch.tensor([float("nan"), float("inf"), float("-inf"), -0.0])
    t = torch.randint(0, 2, (n + 1,), generator=g)
    if label_dtype in (torch.int64, torch.int32, torch.int8):
        t[6000:6010] = -1  # ignored entries
    if label_dtype != torch.bool:
        t[6010:6020] = 2 if label_dtype != torch.int8 else 2
    pd, td, thr_d = p.to(DEV), t.to(label_dtype).to(DEV), thr.to(DEV)
    fast = _native.binned_curve_update(pd[1:].clone(), td[1:].clone(), thr_d, 1)   # fresh allocations: aligned
    slow = _native.binned_curve_update(pd[1:], td[1:].clone() if label_dtype != torch.int64 else td[1:], thr_d, 1)  # misaligned view
    assert torch.equal(fast, slow)
    tn = t[1:].to(label_dtype).to(torch.int64).numpy()
    keep = (tn == 0) | (tn == 1)
    want = oc.binned_confmat(p[1:].numpy()[keep], tn[keep], thr.numpy(), 1)
    np.testing.assert_array_equal(fast.cpu().numpy(), want.reshape(fast.shape))
