"""LogAUC.  CPU: the window integration against the reference's on identical, strictly increasing ROC curves, and on the
oracle's ROC of the end-to-end case.  GPU: the functional / class end to end.  (With repeated fpr values the reference
interpolates through an unstable argsort — implementation-defined — so those cases are not parity cases; ours is stable.)"""
import numpy as np
import pytest
import torch

from metrics_b200.functional.classification.logauc import _binary_logauc_compute
from oracle import curves as oc

DEV = "cuda:0"

RANGES = ((0.001, 0.1), (0.01, 0.5), (0.0005, 1.0))


def test_window_integration_vs_reference(golden_logauc):
    g = golden_logauc
    for k in range(12):
        fpr, tpr = torch.from_numpy(g[f"curve/{k}/fpr"]), torch.from_numpy(g[f"curve/{k}/tpr"])
        for j, rng in enumerate(RANGES):
            np.testing.assert_allclose(_binary_logauc_compute(fpr, tpr, rng).numpy(), g[f"curve/{k}/logauc{j}"], rtol=2e-6, atol=1e-7)


def test_end_to_end_case_on_oracle_roc(golden_logauc):
    g = golden_logauc
    fpr, tpr, _ = oc.binary_roc_ref32(g["b/preds"], g["b/target"])
    for j, rng in enumerate(RANGES[:2]):
        got = _binary_logauc_compute(torch.from_numpy(fpr.copy()), torch.from_numpy(tpr.copy()), rng)
        np.testing.assert_allclose(got.numpy(), g[f"b/logauc{j}"], rtol=2e-6)


@pytest.mark.gpu
def test_functional_and_class_gpu(golden_logauc):
    import metrics_b200.classification as TC
    import metrics_b200.functional.classification as F

    g = golden_logauc
    p, t = torch.from_numpy(g["b/preds"]).to(DEV), torch.from_numpy(g["b/target"]).to(DEV)
    for j, rng in enumerate(RANGES[:2]):
        np.testing.assert_allclose(F.binary_logauc(p, t, fpr_range=rng).cpu().numpy(), g[f"b/logauc{j}"], rtol=2e-6)
    m = TC.BinaryLogAUC(fpr_range=RANGES[0]).to(DEV)
    for a, b in zip(p.chunk(3), t.chunk(3)):
        m.update(a, b)
    np.testing.assert_allclose(m.compute().cpu().numpy(), g["b/logauc0"], rtol=2e-6)
    # multiclass / multilabel: consistent with the per-class binary evaluation of the one-vs-rest problems
    gen = torch.Generator().manual_seed(9)
    lg = torch.randn(2000, 4, generator=gen).to(DEV)
    tg = torch.randint(0, 4, (2000,), generator=gen).to(DEV)
    per_class = F.multiclass_logauc(lg, tg, 4, average="none")
    probs = torch.softmax(lg, 1)
    for c in range(4):
        one = F.binary_logauc(probs[:, c].contiguous(), (tg == c).long())
        np.testing.assert_allclose(per_class[c].cpu().numpy(), one.cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(F.multiclass_logauc(lg, tg, 4, average="macro").cpu().numpy(), per_class.mean().cpu().numpy(), rtol=1e-6)
    assert isinstance(TC.LogAUC(task="multilabel", num_labels=3), TC.MultilabelLogAUC)
