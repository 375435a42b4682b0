"""GPU: K2 counting kernel behind the binary / multilabel stat-score, confusion-matrix, accuracy and F-beta metrics.
Integer outputs bit-exact vs the reference goldens; float reductions within 1e-6 relative."""
import numpy as np
import pytest
import torch

from oracle import classification as oc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KINDS = ["probs", "logits", "labels"]


def _fc():
    import metrics_b200.functional.classification as fc

    return fc


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ign", [None, -1])
def test_binary_functionals_vs_golden(golden_cls, kind, ign):
    fc, g = _fc(), golden_cls
    p = torch.from_numpy(g[f"bin2/{kind}/preds"]).to(DEV)
    t = torch.from_numpy(g["bin2/target"] if ign is None else g["bin2/target_ign"]).to(DEV)
    it = "none" if ign is None else str(ign)
    for mda in ("global", "samplewise"):
        tag = f"bin2/{kind}/ign{it}/{mda}"
        got = fc.binary_stat_scores(p, t, multidim_average=mda, ignore_index=ign)
        assert got.dtype == torch.int64
        np.testing.assert_array_equal(got.cpu().numpy(), g[f"{tag}/stat_scores"])
        np.testing.assert_allclose(fc.binary_accuracy(p, t, multidim_average=mda, ignore_index=ign).cpu().numpy(), g[f"{tag}/accuracy"], rtol=1e-6)
        np.testing.assert_allclose(fc.binary_f1_score(p, t, multidim_average=mda, ignore_index=ign).cpu().numpy(), g[f"{tag}/f1"], rtol=1e-6)
    np.testing.assert_array_equal(fc.binary_confusion_matrix(p, t, ignore_index=ign).cpu().numpy(), g[f"bin2/{kind}/ign{it}/confmat"])
    if kind == "probs" and ign is None:
        np.testing.assert_array_equal(fc.binary_stat_scores(p, t, threshold=0.3).cpu().numpy(), g["bin2/probs/thr0.3/stat_scores"])
        for dt in (torch.float64, torch.float16, torch.bfloat16):
            pp = p.to(dt)
            tp, fp, tn, fn = oc.binary_stat_scores(pp.float().cpu().numpy() if dt != torch.float64 else pp.cpu().numpy(), t.cpu().numpy())
            np.testing.assert_array_equal(fc.binary_stat_scores(pp, t).cpu().numpy(), np.array([tp, fp, tn, fn, tp + fn]))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ign", [None, -1])
def test_multilabel_functionals_vs_golden(golden_cls, kind, ign):
    fc, g = _fc(), golden_cls
    p = torch.from_numpy(g[f"ml/{kind}/preds"]).to(DEV)
    t = torch.from_numpy(g["ml/target"] if ign is None else g["ml/target_ign"]).to(DEV)
    it = "none" if ign is None else str(ign)
    for mda in ("global", "samplewise"):
        for avg in ("micro", "macro", "weighted", "none"):
            tag = f"ml/{kind}/ign{it}/{mda}/{avg}"
            got = fc.multilabel_stat_scores(p, t, 6, average=avg, multidim_average=mda, ignore_index=ign)
            if avg in ("micro", "none"):
                np.testing.assert_array_equal(got.cpu().numpy(), g[f"{tag}/stat_scores"])
            else:
                np.testing.assert_allclose(got.cpu().numpy(), g[f"{tag}/stat_scores"], rtol=1e-6)
            np.testing.assert_allclose(
                fc.multilabel_accuracy(p, t, 6, average=avg, multidim_average=mda, ignore_index=ign).cpu().numpy(),
                g[f"{tag}/accuracy"], rtol=1e-6, equal_nan=True)
            np.testing.assert_allclose(
                fc.multilabel_f1_score(p, t, 6, average=avg, multidim_average=mda, ignore_index=ign).cpu().numpy(),
                g[f"{tag}/f1"], rtol=1e-6, equal_nan=True)
    np.testing.assert_array_equal(fc.multilabel_confusion_matrix(p, t, 6, ignore_index=ign).cpu().numpy(), g[f"ml/{kind}/ign{it}/confmat"])


def test_modular_classes_task_wrappers_and_validation(golden_cls):
    from metrics_b200.classification import Accuracy, BinaryStatScores, ConfusionMatrix, F1Score, MultilabelStatScores, StatScores

    g = golden_cls
    p = torch.from_numpy(g["bin2/logits/preds"]).to(DEV)
    t = torch.from_numpy(g["bin2/target"]).to(DEV)
    m = BinaryStatScores().to(DEV)
    acc = Accuracy(task="binary").to(DEV)
    f1 = F1Score(task="binary").to(DEV)
    cm = ConfusionMatrix(task="binary").to(DEV)
    # NOTE: the logits decision is per update() call; feed the whole tensor at once like the golden did
    for metric in (m, acc, f1, cm):
        metric.update(p, t)
    np.testing.assert_array_equal(m.compute().cpu().numpy(), g["bin2/logits/ignnone/global/stat_scores"])
    np.testing.assert_allclose(acc.compute().cpu().numpy(), g["bin2/logits/ignnone/global/accuracy"], rtol=1e-6)
    np.testing.assert_allclose(f1.compute().cpu().numpy(), g["bin2/logits/ignnone/global/f1"], rtol=1e-6)
    np.testing.assert_array_equal(cm.compute().cpu().numpy(), g["bin2/logits/ignnone/confmat"])
    sw = BinaryStatScores(multidim_average="samplewise").to(DEV)
    sw.update(p[:32], t[:32])
    sw.update(p[32:], t[32:])
    # samplewise results of two half batches concatenate (logit decision happens to agree for both halves here)
    assert sw.compute().shape == (64, 5)
    mp = torch.from_numpy(g["ml/probs/preds"]).to(DEV)
    mt = torch.from_numpy(g["ml/target"]).to(DEV)
    ml = MultilabelStatScores(num_labels=6, average="none").to(DEV)
    ml.update(mp[:20], mt[:20])
    ml.update(mp[20:], mt[20:])
    np.testing.assert_array_equal(ml.compute().cpu().numpy(), g["ml/probs/ignnone/global/none/stat_scores"])
    assert isinstance(StatScores(task="multilabel", num_labels=6), MultilabelStatScores)
    fc = _fc()
    with pytest.raises(RuntimeError, match="Detected the following values in `target`"):
        fc.binary_stat_scores(torch.rand(8, device=DEV), torch.full((8,), 2, device=DEV))
    with pytest.raises(RuntimeError, match="Detected the following values in `preds`"):
        fc.binary_stat_scores(torch.full((8,), 3, device=DEV), torch.ones(8, dtype=torch.long, device=DEV))
    with pytest.raises(ValueError, match="Expected argument `threshold` to be a float"):
        fc.binary_stat_scores(torch.rand(8, device=DEV), torch.ones(8, dtype=torch.long, device=DEV), threshold=2)


def test_large_random_vs_oracle_paths():
    """big single-group (warp-reduced), many-label (shared-memory) and > 2048-group (global atomics) paths"""
    fc = _fc()
    g = torch.Generator().manual_seed(4)
    p = torch.randn(300000, generator=g)
    t = torch.randint(0, 2, (300000,), generator=g)
    tp, fp, tn, fn = oc.binary_stat_scores(p.numpy(), t.numpy())
    np.testing.assert_array_equal(fc.binary_stat_scores(p.to(DEV), t.to(DEV), validate_args=False).cpu().numpy(), np.array([tp, fp, tn, fn, tp + fn]))
    for L, N in ((300, 500), (3000, 40)):
        p = torch.rand(N, L, 3, generator=g)
        t = torch.randint(0, 2, (N, L, 3), generator=g)
        tp, fp, tn, fn = oc.multilabel_stat_scores(p.numpy(), t.numpy(), L)
        got = fc.multilabel_stat_scores(p.to(DEV), t.to(DEV), L, average="none", validate_args=False)
        np.testing.assert_array_equal(got.cpu().numpy(), np.stack([tp, fp, tn, fn, tp + fn], -1))
        tp, fp, tn, fn = oc.multilabel_stat_scores(p.numpy(), t.numpy(), L, samplewise=True)
        got = fc.multilabel_stat_scores(p.to(DEV), t.to(DEV), L, average="none", multidim_average="samplewise", validate_args=False)
        np.testing.assert_array_equal(got.cpu().numpy(), np.stack([tp, fp, tn, fn, tp + fn], -1))
