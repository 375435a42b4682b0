"""CPU: numpy oracle for the binary / multilabel stat-score & confusion-matrix family vs the reference goldens."""
import numpy as np
import pytest

from oracle import classification as oc

KINDS = ["probs", "logits", "labels"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ign", [None, -1])
def test_binary(golden_cls, kind, ign):
    g = golden_cls
    p = g[f"bin2/{kind}/preds"]
    t = g["bin2/target"] if ign is None else g["bin2/target_ign"]
    it = "none" if ign is None else str(ign)
    for mda in ("global", "samplewise"):
        tp, fp, tn, fn = oc.binary_stat_scores(p, t, 0.5, ign, mda == "samplewise")
        ref = g[f"bin2/{kind}/ign{it}/{mda}/stat_scores"]
        np.testing.assert_array_equal(np.stack([tp, fp, tn, fn, tp + fn], axis=0 if mda == "global" else 1), ref)
    tp, fp, tn, fn = oc.binary_stat_scores(p, t, 0.5, ign)
    np.testing.assert_array_equal(oc.confmat_from_counts(tp, fp, tn, fn), g[f"bin2/{kind}/ign{it}/confmat"])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ign", [None, -1])
def test_multilabel(golden_cls, kind, ign):
    g = golden_cls
    p = g[f"ml/{kind}/preds"]
    t = g["ml/target"] if ign is None else g["ml/target_ign"]
    it = "none" if ign is None else str(ign)
    for mda in ("global", "samplewise"):
        tp, fp, tn, fn = oc.multilabel_stat_scores(p, t, 6, 0.5, ign, mda == "samplewise")
        np.testing.assert_array_equal(np.stack([tp, fp, tn, fn, tp + fn], axis=-1), g[f"ml/{kind}/ign{it}/{mda}/none/stat_scores"])
    tp, fp, tn, fn = oc.multilabel_stat_scores(p, t, 6, 0.5, ign)
    np.testing.assert_array_equal(oc.confmat_from_counts(tp, fp, tn, fn), g[f"ml/{kind}/ign{it}/confmat"])


def test_group_fairness(golden_fairness):
    from oracle import classification as oc

    g = golden_fairness
    for c in range(int(g["n_cases"])):
        key = f"case{c}"
        _, ign = (int(v) for v in g[f"{key}/meta"])
        counts = oc.binary_groups_stat_scores(g[f"{key}/preds"], g[f"{key}/target"], g[f"{key}/groups"], 0.5, None if ign == -999 else ign)
        np.testing.assert_array_equal(counts, g[f"{key}/counts"], err_msg=key)
        want = dict(zip(str(g[f"{key}/fair_keys"]).split(","), g[f"{key}/fair_values"].tolist()))
        got = oc.fairness_ratios(counts)
        assert list(got) == list(want), key
        for k in want:
            assert (np.isnan(got[k]) and np.isnan(want[k])) or abs(float(got[k]) - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (key, k)
