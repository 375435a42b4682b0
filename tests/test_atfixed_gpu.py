"""GPU: recall@precision / precision@recall / sensitivity@specificity / specificity@sensitivity end to end (curve kernels
+ device-side operating-point selection) vs goldens from the unmodified reference; exact and binned modes."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FAMS = [("recall_at_fixed_precision", "min_precision"), ("precision_at_fixed_recall", "min_recall"),
        ("sensitivity_at_specificity", "min_specificity"), ("specificity_at_sensitivity", "min_sensitivity")]
FLOORS = [0.0, 0.35, 0.6, 0.9, 1.0]
THRS = [("exact", None), ("int21", 21), ("list", [0.2, 0.5, 0.8])]


def _d(g, key):
    return torch.from_numpy(g[key]).to(DEV)


@pytest.mark.parametrize("fam,arg", FAMS)
@pytest.mark.parametrize("tname,thr", THRS)
def test_functionals(golden_atfixed, fam, arg, tname, thr):
    import metrics_b200.functional.classification as F

    g = golden_atfixed
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for floor in FLOORS:
            tag = f"{fam}/{floor}/{tname}"
            # logits: a sigmoid output within an ulp of a bin edge may move one count in binned mode
            for pre, call, tol in (
                ("b", lambda: getattr(F, f"binary_{fam}")(_d(g, "b/preds"), _d(g, "b/target"), **{arg: floor}, thresholds=thr), 1e-6),
                ("bl", lambda: getattr(F, f"binary_{fam}")(_d(g, "b/logits"), _d(g, "b/target")[:900], floor, thresholds=thr), 2e-3 if thr else 1e-6),
                ("mc", lambda: getattr(F, f"multiclass_{fam}")(_d(g, "mc/logits"), _d(g, "mc/target"), 5, floor, thresholds=thr), 2e-3 if thr else 1e-6),
                ("ml", lambda: getattr(F, f"multilabel_{fam}")(_d(g, "ml/preds"), _d(g, "ml/target"), 4, floor, thresholds=thr), 1e-6),
                ("mli", lambda: getattr(F, f"multilabel_{fam}")(_d(g, "ml/preds"), _d(g, "ml/target_ign"), 4, floor, thresholds=thr, ignore_index=-1), 1e-6),
            ):
                v, t = call()
                np.testing.assert_allclose(v.cpu().numpy(), g[f"{pre}/{tag}/value"], rtol=tol, atol=tol / 10, err_msg=f"{pre}/{tag}")
                np.testing.assert_allclose(t.cpu().numpy(), g[f"{pre}/{tag}/thr"], rtol=max(tol, 1e-6), atol=tol / 10, err_msg=f"{pre}/{tag}")


def test_modular_classes(golden_atfixed):
    import metrics_b200.classification as TC

    g = golden_atfixed
    m = TC.BinaryRecallAtFixedPrecision(min_precision=0.6).to(DEV)
    m2 = TC.MulticlassSpecificityAtSensitivity(num_classes=5, min_sensitivity=0.5, thresholds=30).to(DEV)
    for a, b in zip(_d(g, "b/preds").chunk(3), _d(g, "b/target").chunk(3)):
        m.update(a, b)
    for a, b in zip(_d(g, "mc/logits").chunk(3), _d(g, "mc/target").chunk(3)):
        m2.update(a, b)
    v, t = m.compute()
    np.testing.assert_allclose(v.cpu().numpy(), g["class/b_recall_at_p/value"], rtol=1e-6)
    np.testing.assert_allclose(t.cpu().numpy(), g["class/b_recall_at_p/thr"], rtol=1e-6)
    v, t = m2.compute()
    np.testing.assert_allclose(v.cpu().numpy(), g["class/mc_spec_at_sens/value"], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(t.cpu().numpy(), g["class/mc_spec_at_sens/thr"], rtol=2e-3, atol=2e-4)
    assert isinstance(TC.RecallAtFixedPrecision(task="binary", min_precision=0.5), TC.BinaryRecallAtFixedPrecision)
    assert isinstance(TC.SensitivityAtSpecificity("multilabel", 0.5, num_labels=3), TC.MultilabelSensitivityAtSpecificity)
