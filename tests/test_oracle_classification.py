"""CPU: the numpy oracle reproduces every golden vector produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import classification as oc
from tests.helpers import cfg1_inputs, cfg2_inputs, sha, stats_inputs, to_np


@pytest.mark.parametrize("C", [4, 37, 64, 1000, 1024, 2500])
@pytest.mark.parametrize("dt", ["f32", "bf16", "f16", "f64"])
def test_argmax_edge_semantics(golden_cls, C, dt):
    x = golden_cls[f"argmax/{dt}/C{C}/x"]
    np.testing.assert_array_equal(oc.argmax_dim1(x), golden_cls[f"argmax/{dt}/C{C}/y"])


@pytest.mark.parametrize("C", [5, 37, 130])
@pytest.mark.parametrize("ign", [None, -1, 0])
def test_confmat_cases(golden_cls, C, ign):
    tag = "none" if ign is None else str(ign)
    logits, labels = golden_cls[f"confmat/C{C}/logits"], golden_cls[f"confmat/C{C}/labels"]
    t = golden_cls[f"confmat/C{C}/ign{tag}/target"]
    np.testing.assert_array_equal(
        oc.multiclass_confusion_matrix(logits, t, C, ign), golden_cls[f"confmat/C{C}/ign{tag}/from_logits"])
    np.testing.assert_array_equal(
        oc.multiclass_confusion_matrix(labels, t, C, ign), golden_cls[f"confmat/C{C}/ign{tag}/from_labels"])


def test_confmat_multidim_and_uint8(golden_cls):
    np.testing.assert_array_equal(
        oc.multiclass_confusion_matrix(golden_cls["confmat/multidim/logits"], golden_cls["confmat/multidim/target"], 6),
        golden_cls["confmat/multidim/confmat"])
    np.testing.assert_array_equal(
        oc.multiclass_confusion_matrix(golden_cls["confmat/uint8/preds"], golden_cls["confmat/uint8/target"], 200),
        golden_cls["confmat/uint8/confmat"])


def test_cfg1_states_and_value(golden_cls):
    preds, target = cfg1_inputs()
    assert sha(preds) == str(golden_cls["cfg1/preds_sha256"])
    assert sha(target) == str(golden_cls["cfg1/target_sha256"])
    acc = [np.zeros(5, np.int64) for _ in range(4)]
    for i in range(100):
        for a, d in zip(acc, oc.multiclass_stat_scores(preds[i].numpy(), target[i].numpy(), 5, "macro")):
            a += d
    for a, s in zip(acc, ("tp", "fp", "tn", "fn")):
        np.testing.assert_array_equal(a, golden_cls[f"cfg1/{s}"])
    assert oc.accuracy_reduce(*acc, "macro") == golden_cls["cfg1/value"]
    assert float(golden_cls["cfg1/value"]) == pytest.approx(0.1986250877380371, abs=0)


def test_cfg2_argmax_and_confmat_digest(golden_cls):
    logits, target = cfg2_inputs()
    assert sha(logits) == str(golden_cls["cfg2/logits_sha256"])
    assert sha(target) == str(golden_cls["cfg2/target_sha256"])
    x = logits.float().numpy()
    am = oc.argmax_dim1(x)
    np.testing.assert_array_equal(am.astype(np.int16), golden_cls["cfg2/argmax_i16"])
    cm = oc.multiclass_confusion_matrix(x, target.numpy(), 1000)
    assert sha(torch.from_numpy(cm)) == str(golden_cls["cfg2/confmat_sha256"])
    assert int(golden_cls["cfg2/tied_rows"]) > 1000  # the 2.6 % tied-maximum rows are part of the workload


@pytest.mark.parametrize("C,N", [(5, 300), (1000, 4096)])
@pytest.mark.parametrize("avg", ["micro", "macro", "weighted", "none"])
@pytest.mark.parametrize("ign", [None, -1, 1])
def test_stat_scores_accuracy_fbeta(golden_cls, C, N, avg, ign):
    logits, target = stats_inputs(C, N)
    t = target.clone()
    if ign == -1:
        t[::5] = -1
    tag = f"stats/C{C}/{avg}/ign{'none' if ign is None else ign}"
    tp, fp, tn, fn = oc.multiclass_stat_scores(logits.numpy(), t.numpy(), C, avg, ign)
    stacked = np.stack([tp, fp, tn, fn, tp + fn], axis=-1)
    ref = golden_cls[f"{tag}/stat_scores"]
    if avg == "micro":
        np.testing.assert_array_equal(stacked, ref)
    elif avg == "none":
        np.testing.assert_array_equal(stacked, ref)
    elif avg == "macro":
        np.testing.assert_allclose(stacked.astype(np.float32).mean(0), ref, rtol=1e-6)
    np.testing.assert_allclose(oc.accuracy_reduce(tp, fp, tn, fn, avg), golden_cls[f"{tag}/accuracy"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(oc.fbeta_reduce(tp, fp, tn, fn, 1.0, avg), golden_cls[f"{tag}/f1"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(oc.fbeta_reduce(tp, fp, tn, fn, 2.0, avg), golden_cls[f"{tag}/fbeta2"], rtol=1e-6, atol=0)
