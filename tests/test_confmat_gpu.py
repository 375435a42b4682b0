"""GPU: K1/K1b kernels (through the C-ABI) vs the oracle and the reference goldens — bit-exact integer states."""
import numpy as np
import pytest
import torch

from oracle import classification as oc
from tests.helpers import TORCH_DTYPES, cfg1_inputs, cfg2_inputs, sha, stats_inputs, to_np

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _native():
    from metrics_b200 import _native

    return _native


@pytest.mark.parametrize("C", [4, 37, 64, 1000, 1024, 2500])
@pytest.mark.parametrize("dt", ["f32", "bf16", "f16", "f64"])
def test_argmax_kernel_edge_semantics(golden_cls, C, dt):
    x = torch.from_numpy(golden_cls[f"argmax/{dt}/C{C}/x"]).to(TORCH_DTYPES[dt]).to(DEV)
    got = _native().argmax_rows(x).cpu().numpy()
    np.testing.assert_array_equal(got, golden_cls[f"argmax/{dt}/C{C}/y"])


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16", "f64"])
@pytest.mark.parametrize("N,C", [(1, 1000), (33, 1000), (257, 40), (100, 7), (64, 999), (50, 4100), (9, 32), (300, 33)])
def test_argmax_kernel_random_vs_oracle(dt, N, C):
    g = torch.Generator().manual_seed(N * 7919 + C)
    x = torch.randn(N, C, generator=g).to(TORCH_DTYPES[dt])
    # inject ties by coarse rounding on half of the rows
    x[::2] = (x[::2] * 2).round() / 2
    got = _native().argmax_rows(x.to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(got, oc.argmax_dim1(to_np(x) if dt != "f64" else x.numpy()))


def test_argmax_unaligned_view():
    x = torch.randn(65, 1001, device=DEV).bfloat16()
    v = x[:, 1:]  # contiguous() copy inside the wrapper -> aligned again; also test an odd width
    got = _native().argmax_rows(v).cpu().numpy()
    np.testing.assert_array_equal(got, oc.argmax_dim1(to_np(v)))
    y = torch.randn(10, 6, 5, 3, device=DEV)
    np.testing.assert_array_equal(_native().argmax_rows(y).cpu().numpy(), y.cpu().argmax(dim=1).numpy())


@pytest.mark.parametrize("C", [5, 37, 130])
@pytest.mark.parametrize("ign", [None, -1, 0])
def test_functional_confmat_vs_golden(golden_cls, C, ign):
    from metrics_b200.functional.classification import multiclass_confusion_matrix

    tag = "none" if ign is None else str(ign)
    logits = torch.from_numpy(golden_cls[f"confmat/C{C}/logits"]).to(DEV)
    labels = torch.from_numpy(golden_cls[f"confmat/C{C}/labels"]).to(DEV)
    t = torch.from_numpy(golden_cls[f"confmat/C{C}/ign{tag}/target"]).to(DEV)
    for dt in (torch.float32, torch.float64):
        got = multiclass_confusion_matrix(logits.to(dt), t, C, ignore_index=ign)
        assert got.dtype == torch.int64
        np.testing.assert_array_equal(got.cpu().numpy(), golden_cls[f"confmat/C{C}/ign{tag}/from_logits"])
    got = multiclass_confusion_matrix(labels, t, C, ignore_index=ign)
    np.testing.assert_array_equal(got.cpu().numpy(), golden_cls[f"confmat/C{C}/ign{tag}/from_labels"])
    got = multiclass_confusion_matrix(labels.to(torch.int32), t.to(torch.int32), C, ignore_index=ign)
    np.testing.assert_array_equal(got.cpu().numpy(), golden_cls[f"confmat/C{C}/ign{tag}/from_labels"])
    if ign is None:
        for norm in ("true", "pred", "all"):
            got = multiclass_confusion_matrix(logits, t, C, normalize=norm)
            np.testing.assert_allclose(got.cpu().numpy(), golden_cls[f"confmat/C{C}/norm_{norm}"], rtol=1e-6)


def test_confmat_multidim_uint8_empty(golden_cls):
    from metrics_b200.functional.classification import multiclass_confusion_matrix

    got = multiclass_confusion_matrix(
        torch.from_numpy(golden_cls["confmat/multidim/logits"]).to(DEV),
        torch.from_numpy(golden_cls["confmat/multidim/target"]).to(DEV), 6)
    np.testing.assert_array_equal(got.cpu().numpy(), golden_cls["confmat/multidim/confmat"])
    got = multiclass_confusion_matrix(
        torch.from_numpy(golden_cls["confmat/uint8/preds"]).to(DEV),
        torch.from_numpy(golden_cls["confmat/uint8/target"]).to(DEV), 200)
    np.testing.assert_array_equal(got.cpu().numpy(), golden_cls["confmat/uint8/confmat"])
    empty = multiclass_confusion_matrix(torch.zeros(0, 4, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV), 4)
    assert empty.shape == (4, 4) and int(empty.sum()) == 0


def test_confmat_small_c_privatised_path_large_n():
    from metrics_b200.functional.classification import multiclass_confusion_matrix

    g = torch.Generator().manual_seed(3)
    for C in (2, 5, 64):
        logits = torch.randn(20000, C, generator=g)
        target = torch.randint(0, C, (20000,), generator=g)
        got = multiclass_confusion_matrix(logits.to(DEV), target.to(DEV), C, validate_args=False)
        np.testing.assert_array_equal(got.cpu().numpy(), oc.multiclass_confusion_matrix(logits.numpy(), target.numpy(), C))
        labels = torch.randint(0, C, (20000,), generator=g)
        got = multiclass_confusion_matrix(labels.to(DEV), target.to(DEV), C, validate_args=False)
        np.testing.assert_array_equal(got.cpu().numpy(), oc.multiclass_confusion_matrix(labels.numpy(), target.numpy(), C))


def test_validation_flags_out_of_range_labels():
    from metrics_b200.functional.classification import multiclass_confusion_matrix

    logits = torch.randn(8, 4, device=DEV)
    bad_t = torch.tensor([0, 1, 2, 3, 4, 0, 1, 7], device=DEV)
    with pytest.raises(RuntimeError, match="Detected more unique values in `target` than expected"):
        multiclass_confusion_matrix(logits, bad_t, 4)
    bad_p = torch.tensor([0, 1, 2, 3, 9, 0, 1, 2], device=DEV)
    with pytest.raises(RuntimeError, match="Detected more unique values in `preds` than expected"):
        multiclass_confusion_matrix(bad_p, torch.zeros(8, dtype=torch.long, device=DEV), 4)
    with pytest.raises(ValueError, match="should be a float tensor"):
        multiclass_confusion_matrix(torch.zeros(8, 4, dtype=torch.long, device=DEV), bad_t, 4)
    # validate_args=False: offending rows are skipped, nothing is written out of bounds, no host sync
    got = multiclass_confusion_matrix(logits, bad_t, 4, validate_args=False)
    assert int(got.sum()) == 6


def test_cfg1_multiclass_accuracy_bit_exact(golden_cls):
    from metrics_b200.classification import MulticlassAccuracy

    preds, target = cfg1_inputs()
    preds, target = preds.to(DEV), target.to(DEV)
    for validate in (True, False):
        m = MulticlassAccuracy(num_classes=5, validate_args=validate).to(DEV)
        for i in range(100):
            m.update(preds[i], target[i])
        for s in ("tp", "fp", "tn", "fn"):
            np.testing.assert_array_equal(getattr(m, s).cpu().numpy(), golden_cls[f"cfg1/{s}"])
        val = m.compute()
        assert val.dtype == torch.float32
        # float32 ratio of exact integer states: tolerance 1e-6 relative (north_star); the reduction over the
        # 5 classes runs in a CUDA kernel whose summation order may differ from the CPU's by one ulp
        ref = float(golden_cls["cfg1/value"])
        assert ref == 0.1986250877380371
        assert abs(float(val) - ref) <= 1e-6 * ref


def test_cfg2_confmat_bit_exact_full_size(golden_cls):
    from metrics_b200.classification import MulticlassConfusionMatrix

    logits, target = cfg2_inputs()
    assert sha(logits) == str(golden_cls["cfg2/logits_sha256"])
    m = MulticlassConfusionMatrix(num_classes=1000, validate_args=False).to(DEV)
    dl, dt = logits.to(DEV), target.to(DEV)
    m.update(dl, dt)
    cm = m.compute().cpu()
    assert sha(cm) == str(golden_cls["cfg2/confmat_sha256"])
    np.testing.assert_array_equal(cm.sum(1).numpy(), golden_cls["cfg2/confmat_rowsum"])
    np.testing.assert_array_equal(cm.diag().numpy(), golden_cls["cfg2/confmat_diag"])
    np.testing.assert_array_equal(
        _native().argmax_rows(dl).cpu().numpy().astype(np.int16), golden_cls["cfg2/argmax_i16"])
    # size-independent properties: linearity over batches and over a row permutation
    m.update(dl, dt)
    assert torch.equal(m.confmat.cpu(), 2 * cm)
    perm = torch.randperm(65536, device=DEV)
    m2 = MulticlassConfusionMatrix(num_classes=1000, validate_args=False).to(DEV)
    m2.update(dl[perm], dt[perm])
    assert torch.equal(m2.confmat.cpu(), cm)
    assert int(cm.sum()) == 65536


@pytest.mark.parametrize("C,N", [(5, 300), (1000, 4096)])
@pytest.mark.parametrize("avg", ["micro", "macro", "weighted", "none"])
@pytest.mark.parametrize("ign", [None, -1, 1])
def test_stat_scores_accuracy_f1_vs_golden(golden_cls, C, N, avg, ign):
    from metrics_b200.functional.classification import (
        multiclass_accuracy,
        multiclass_f1_score,
        multiclass_fbeta_score,
        multiclass_stat_scores,
    )

    logits, target = stats_inputs(C, N)
    t = target.clone()
    if ign == -1:
        t[::5] = -1
    logits, t = logits.to(DEV), t.to(DEV)
    tag = f"stats/C{C}/{avg}/ign{'none' if ign is None else ign}"
    got = multiclass_stat_scores(logits, t, C, average=avg, ignore_index=ign)
    ref = golden_cls[f"{tag}/stat_scores"]
    if avg in ("micro", "none"):
        assert got.dtype == torch.int64
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
    else:
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-6)
    np.testing.assert_allclose(
        multiclass_accuracy(logits, t, C, average=avg, ignore_index=ign).cpu().numpy(), golden_cls[f"{tag}/accuracy"], rtol=1e-6)
    np.testing.assert_allclose(
        multiclass_f1_score(logits, t, C, average=avg, ignore_index=ign).cpu().numpy(), golden_cls[f"{tag}/f1"], rtol=1e-6)
    np.testing.assert_allclose(
        multiclass_fbeta_score(logits, t, 2.0, C, average=avg, ignore_index=ign).cpu().numpy(), golden_cls[f"{tag}/fbeta2"], rtol=1e-6)


@pytest.mark.parametrize("C,N", [(5, 300), (1000, 4096)])
@pytest.mark.parametrize("avg", ["micro", "macro"])
def test_stat_scores_class_states_over_batches(golden_cls, C, N, avg):
    from metrics_b200.classification import MulticlassF1Score, MulticlassStatScores

    logits, target = stats_inputs(C, N)
    logits, target = logits.to(DEV), target.to(DEV)
    mm = MulticlassStatScores(num_classes=C, average=avg).to(DEV)
    f1 = MulticlassF1Score(num_classes=C, average=avg).to(DEV)
    for cl, ct in zip(logits.chunk(4), target.chunk(4)):
        mm.update(cl, ct)
        f1.update(cl, ct)
    for s in ("tp", "fp", "tn", "fn"):
        np.testing.assert_array_equal(getattr(mm, s).cpu().numpy(), golden_cls[f"stats/C{C}/{avg}/class_state_{s}"])
    np.testing.assert_allclose(f1.compute().cpu().numpy(), golden_cls[f"stats/C{C}/{avg}/class_f1"], rtol=1e-6)
    # the workspace is self-cleaning
    assert int(mm._scratch.abs().sum()) == 0


def test_stat_scores_privatised_and_label_paths_vs_oracle():
    from metrics_b200.functional.classification import multiclass_stat_scores

    g = torch.Generator().manual_seed(5)
    for C, N in ((3, 50000), (100, 30000), (2048, 9000), (3000, 5000)):
        labels = torch.randint(0, C, (N,), generator=g)
        target = torch.randint(0, C, (N,), generator=g)
        target[::11] = -1
        for avg in ("micro", "none"):
            got = multiclass_stat_scores(labels.to(DEV), target.to(DEV), C, average=avg, ignore_index=-1, validate_args=False)
            tp, fp, tn, fn = oc.multiclass_stat_scores(labels.numpy(), target.numpy(), C, avg, -1)
            np.testing.assert_array_equal(got.cpu().numpy(), np.stack([tp, fp, tn, fn, tp + fn], axis=-1))


@pytest.mark.parametrize("average", ["macro", "micro"])
def test_large_stat_score_updates_with_deferred_fold_equal_chunked_updates(average):
    """Updates of >= 2^24 scores run the row kernel without the last-CTA fold and a one-CTA fold kernel behind it
    (csrc/sinks.cuh kDeferFold): several back-to-back large updates must leave exactly the states that the same rows leave when
    fed in small chunks (single-launch path), and the workspace must come back clean."""
    from metrics_b200.classification import MulticlassStatScores

    g = torch.Generator().manual_seed(99)
    big = MulticlassStatScores(num_classes=1000, average=average, validate_args=False).to(DEV)
    small = MulticlassStatScores(num_classes=1000, average=average, validate_args=False).to(DEV)
    for _ in range(4):
        lg = torch.randn(20000, 1000, generator=g).bfloat16().to(DEV)
        tg = torch.randint(0, 1000, (20000,), generator=g).to(DEV)
        big.update(lg, tg)  # 2e7 scores: deferred fold
        for lo in range(0, 20000, 4000):
            small.update(lg[lo:lo + 4000], tg[lo:lo + 4000])  # 4e6 scores: fold inside the row kernel
    for name in ("tp", "fp", "tn", "fn"):
        assert torch.equal(getattr(big, name), getattr(small, name)), name
    assert int(big._workspace(1000, torch.device(DEV)).abs().sum()) == 0


def test_failed_validation_leaves_the_state_untouched():
    """`validate_args=True`: the range check runs inside the counting kernel, but a batch with an out-of-range label must not
    leave its in-range rows in the state (the reference validates before it counts): a caller that catches the error and goes
    on sees exactly the state from before the bad batch."""
    from metrics_b200.classification import MulticlassConfusionMatrix, MulticlassStatScores

    g = torch.Generator().manual_seed(21)
    lg = torch.randn(500, 7, generator=g).to(DEV)
    good = torch.randint(0, 7, (500,), generator=g).to(DEV)
    bad = good.clone()
    bad[123] = 9
    for metric in (MulticlassConfusionMatrix(num_classes=7), MulticlassStatScores(num_classes=7, average=None)):
        metric = metric.to(DEV)
        metric.update(lg, good)
        before = {k: v.clone() for k, v in metric.metric_state.items()}
        with pytest.raises(RuntimeError, match="Detected more unique values"):
            metric.update(lg, bad)
        for k, v in metric.metric_state.items():
            assert torch.equal(v, before[k]), k
        metric.update(lg, good)  # and the metric keeps working
        for k, v in metric.metric_state.items():
            assert torch.equal(v, 2 * before[k]), k
