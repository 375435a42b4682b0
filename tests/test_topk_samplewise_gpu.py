"""GPU: multiclass top-k and samplewise stat scores (K1b variants) vs the reference goldens: integer outputs bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fc():
    import metrics_b200.functional.classification as fc

    return fc


@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("avg", ["micro", "macro", "none"])
@pytest.mark.parametrize("ign", [None, -1])
def test_topk_vs_golden(golden_cls, k, avg, ign):
    fc, g = _fc(), golden_cls
    p = torch.from_numpy(g["topk/logits"]).to(DEV)
    t = torch.from_numpy(g["topk/target"] if ign is None else g["topk/target_ign"]).to(DEV)
    tag = f"topk/k{k}/{avg}/ign{'none' if ign is None else ign}"
    got = fc.multiclass_stat_scores(p, t, 7, average=avg, top_k=k, ignore_index=ign)
    if avg == "macro":
        np.testing.assert_allclose(got.cpu().numpy(), g[f"{tag}/stat_scores"], rtol=1e-6)
    else:
        np.testing.assert_array_equal(got.cpu().numpy(), g[f"{tag}/stat_scores"])
    np.testing.assert_allclose(fc.multiclass_accuracy(p, t, 7, average=avg, top_k=k, ignore_index=ign).cpu().numpy(), g[f"{tag}/accuracy"], rtol=1e-6)
    np.testing.assert_allclose(fc.multiclass_f1_score(p, t, 7, average=avg, top_k=k, ignore_index=ign).cpu().numpy(), g[f"{tag}/f1"], rtol=1e-6)
    for dt in (torch.bfloat16, torch.float64):  # other dtypes: compare against the fp32 result on exactly representable data
        pp = p.to(torch.bfloat16).to(dt)
        ref = fc.multiclass_stat_scores(p.to(torch.bfloat16).float(), t, 7, average="none", top_k=k, ignore_index=ign)
        assert torch.equal(fc.multiclass_stat_scores(pp, t, 7, average="none", top_k=k, ignore_index=ign), ref)


@pytest.mark.parametrize("kind", ["logits", "labels"])
@pytest.mark.parametrize("avg", ["micro", "macro", "none"])
@pytest.mark.parametrize("ign", [None, -1, 1])
def test_samplewise_vs_golden(golden_cls, kind, avg, ign):
    fc, g = _fc(), golden_cls
    p = torch.from_numpy(g[f"sw/{kind}"]).to(DEV)
    t = torch.from_numpy(g["sw/target"]).to(DEV).clone()
    if ign == -1:
        t[:, ::4] = -1
    tag = f"sw/{kind}/{avg}/ign{'none' if ign is None else ign}"
    got = fc.multiclass_stat_scores(p, t, 5, average=avg, multidim_average="samplewise", ignore_index=ign)
    if avg == "macro":
        np.testing.assert_allclose(got.cpu().numpy(), g[f"{tag}/stat_scores"], rtol=1e-6)
    else:
        np.testing.assert_array_equal(got.cpu().numpy(), g[f"{tag}/stat_scores"])
    np.testing.assert_allclose(
        fc.multiclass_accuracy(p, t, 5, average=avg, multidim_average="samplewise", ignore_index=ign).cpu().numpy(),
        g[f"{tag}/accuracy"], rtol=1e-6, equal_nan=True)


def test_modular_topk_and_samplewise(golden_cls):
    from metrics_b200.classification import MulticlassAccuracy, MulticlassStatScores

    g = golden_cls
    p, t = torch.from_numpy(g["topk/logits"]).to(DEV), torch.from_numpy(g["topk/target"]).to(DEV)
    m = MulticlassAccuracy(num_classes=7, top_k=2, average="micro").to(DEV)
    for a, b in zip(p.chunk(3), t.chunk(3)):
        m.update(a, b)
    np.testing.assert_allclose(m.compute().cpu().numpy(), g["topk/class_acc_k2_micro"], rtol=1e-6)
    sp, st = torch.from_numpy(g["sw/logits"]).to(DEV), torch.from_numpy(g["sw/target"]).to(DEV)
    msw = MulticlassStatScores(num_classes=5, average="none", multidim_average="samplewise").to(DEV)
    msw.update(sp[:8], st[:8])
    msw.update(sp[8:], st[8:])
    np.testing.assert_array_equal(msw.compute().cpu().numpy(), g["sw/class_none"])
    with pytest.raises(NotImplementedError, match="top_k > 1"):
        MulticlassStatScores(num_classes=5, top_k=2, multidim_average="samplewise")
