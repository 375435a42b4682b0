"""GPU: exact curve family (sort + scan kernels through the C-ABI) vs the reference goldens and the oracle.

Tolerance: integer counts (fps/tps) bit-exact; floating AUROC / AP within 1e-6 relative of the reference's fp32 output
(BASELINE.json north_star), thresholds exact for untransformed scores (1e-6 where a sigmoid/softmax was applied).
"""
import warnings

import numpy as np
import pytest
import torch

from oracle import curves as oc
from tests.helpers import MC_CASES, cfg3_inputs, cfg5_rank_batches, mc_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BIN_CASES = ["doc", "rand", "ties", "logits", "allpos", "allneg", "alltied", "one", "odd", "skew"]
RTOL = 1e-6


def _fc():
    import metrics_b200.functional.classification as fc

    return fc


@pytest.mark.parametrize("name", BIN_CASES)
def test_binary_functionals_vs_golden(golden_curves, name):
    fc, g = _fc(), golden_curves
    p = torch.from_numpy(g[f"bin/{name}/preds"]).to(DEV)
    t = torch.from_numpy(g[f"bin/{name}/target"]).to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        auroc, ap = fc.binary_auroc(p, t), fc.binary_average_precision(p, t)
        assert auroc.dtype == torch.float32 and ap.dtype == torch.float32 and auroc.ndim == 0
        np.testing.assert_allclose(auroc.cpu().numpy(), g[f"bin/{name}/auroc"], rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(ap.cpu().numpy(), g[f"bin/{name}/ap"], rtol=RTOL, atol=1e-7)
        fpr, tpr, thr = fc.binary_roc(p, t)
        np.testing.assert_allclose(fpr.cpu().numpy(), g[f"bin/{name}/roc_fpr"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(tpr.cpu().numpy(), g[f"bin/{name}/roc_tpr"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(thr.cpu().numpy(), g[f"bin/{name}/roc_thr"], rtol=RTOL if name == "logits" else 0, atol=0)
        pr, rc, th2 = fc.binary_precision_recall_curve(p, t)
        np.testing.assert_allclose(pr.cpu().numpy(), g[f"bin/{name}/prc_p"], rtol=RTOL, atol=0, equal_nan=True)
        np.testing.assert_allclose(rc.cpu().numpy(), g[f"bin/{name}/prc_r"], rtol=RTOL, atol=0, equal_nan=True)
        np.testing.assert_allclose(th2.cpu().numpy(), g[f"bin/{name}/prc_thr"], rtol=RTOL if name == "logits" else 0, atol=0)
        for mf in (0.5, 0.8):
            got = fc.binary_auroc(p, t, max_fpr=mf)
            np.testing.assert_allclose(got.cpu().numpy(), g[f"bin/{name}/auroc_maxfpr{mf}"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", BIN_CASES)
def test_binary_clf_curve_counts_bit_exact(golden_curves, name):
    from metrics_b200 import _native
    from metrics_b200.functional.classification.precision_recall_curve import _binary_clf_curve

    g = golden_curves
    p = torch.from_numpy(g[f"bin/{name}/preds"]).to(DEV)
    t = torch.from_numpy(g[f"bin/{name}/target"]).to(DEV)
    p = _native.sigmoid_if_logits(p)
    fps, tps, thr = _binary_clf_curve(p, t)
    assert fps.dtype == tps.dtype == thr.dtype == torch.float32
    np.testing.assert_array_equal(fps.cpu().numpy(), g[f"bin/{name}/clf_fps"])
    np.testing.assert_array_equal(tps.cpu().numpy(), g[f"bin/{name}/clf_tps"])


def test_binary_ignore_index_and_bf16(golden_curves):
    fc, g = _fc(), golden_curves
    p = torch.from_numpy(g["bin/rand/preds"]).to(DEV)
    t = torch.from_numpy(g["bin/ignore/target"]).to(DEV)
    np.testing.assert_allclose(fc.binary_auroc(p, t, ignore_index=-1).cpu().numpy(), g["bin/ignore/auroc"], rtol=RTOL)
    np.testing.assert_allclose(fc.binary_average_precision(p, t, ignore_index=-1).cpu().numpy(), g["bin/ignore/ap"], rtol=RTOL)
    pb = torch.from_numpy(g["bin/logits/preds"]).bfloat16().to(DEV)
    tb = torch.from_numpy(g["bin/logits/target"]).to(DEV)
    # bf16 sigmoid outputs collide massively: ranks depend on exact bf16 rounding -> looser but still tight
    np.testing.assert_allclose(fc.binary_auroc(pb, tb).cpu().numpy(), g["bin/bf16/auroc"], rtol=1e-3)
    np.testing.assert_allclose(fc.binary_average_precision(pb, tb).cpu().numpy(), g["bin/bf16/ap"], rtol=1e-3)
    assert fc.binary_roc(pb, tb)[2].dtype == torch.bfloat16


def test_binary_validation_errors():
    fc = _fc()
    p = torch.rand(10, device=DEV)
    with pytest.raises(RuntimeError, match="Detected the following values in `target`"):
        fc.binary_auroc(p, torch.full((10,), 3, device=DEV))
    with pytest.raises(ValueError, match="Expected argument `preds` to be an floating tensor"):
        fc.binary_auroc(torch.ones(10, dtype=torch.long, device=DEV), torch.ones(10, dtype=torch.long, device=DEV))
    assert fc.binary_auroc(p, torch.ones(10, dtype=torch.long, device=DEV), thresholds=10).ndim == 0  # binned mode works


def test_cfg3_collection_full_size(golden_curves):
    """BASELINE cfg3: MetricCollection([BinaryAUROC, BinaryAveragePrecision]), 1000 updates of 10 000, one compute."""
    from metrics_b200 import MetricCollection
    from metrics_b200.classification import BinaryAUROC, BinaryAveragePrecision

    preds, target = cfg3_inputs()
    preds, target = preds.to(DEV), target.to(DEV)
    mc = MetricCollection([BinaryAUROC(validate_args=False), BinaryAveragePrecision(validate_args=False)]).to(DEV)
    for i in range(1000):
        mc.update(preds[i], target[i])
    assert mc.compute_groups == {0: ["BinaryAUROC", "BinaryAveragePrecision"]}
    res = mc.compute()
    np.testing.assert_allclose(res["BinaryAUROC"].cpu().numpy(), golden_curves["cfg3/auroc"], rtol=RTOL)
    np.testing.assert_allclose(res["BinaryAveragePrecision"].cpu().numpy(), golden_curves["cfg3/ap"], rtol=RTOL)
    # size-independent properties at full size: AUROC(1 - p) == 1 - AUROC(p); label flip does the same
    from metrics_b200.functional.classification import binary_auroc

    flat_p, flat_t = preds.reshape(-1), target.reshape(-1)
    a = float(binary_auroc(flat_p, flat_t, validate_args=False))
    assert abs(float(binary_auroc(1.0 - flat_p, flat_t, validate_args=False)) - (1.0 - a)) < 2e-7
    assert abs(float(binary_auroc(flat_p, 1 - flat_t, validate_args=False)) - (1.0 - a)) < 2e-7


def test_sort_scan_random_sizes_vs_oracle():
    """Kernel (counts bit-exact, scalars vs exact fp64 oracle) on ragged sizes around tile boundaries, heavy ties."""
    from metrics_b200 import _native

    g = torch.Generator().manual_seed(99)
    for n in (2, 3, 31, 2047, 2048, 2049, 4095, 4096, 4097, 12289, 100003):
        p = (torch.rand(n, generator=g) * 200).floor() / 200 if n % 2 else torch.randn(n, generator=g)
        t = torch.randint(0, 2, (n,), generator=g)
        auroc, ap, counts, (fps, tps, thr) = _native.curve_evaluate(p.to(DEV), t.to(DEV), 1, 1, want_curve=True)
        rf, rt, rth = oc.binary_clf_curve(p.numpy(), t.numpy())
        u = int(counts[0, 2])
        assert u == rf.size and int(counts[0, 0]) == int(t.sum()) and int(counts[0, 1]) == n - int(t.sum())
        np.testing.assert_array_equal(fps[0, :u].cpu().numpy().astype(np.int64), rf)
        np.testing.assert_array_equal(tps[0, :u].cpu().numpy().astype(np.int64), rt)
        np.testing.assert_array_equal(thr[0, :u].cpu().numpy(), rth)
        np.testing.assert_allclose(float(auroc[0]), oc.binary_auroc_exact(p.numpy(), t.numpy()), rtol=2e-7, atol=1e-7)
        np.testing.assert_allclose(float(ap[0]), oc.binary_average_precision_exact(p.numpy(), t.numpy()), rtol=2e-7, atol=1e-7)


def test_sort_is_deterministic_and_handles_special_values():
    from metrics_b200 import _native

    p = torch.tensor([0.5, float("inf"), -0.0, 0.0, float("-inf"), 0.5, 1e-40, -1e-40, 2.0, float("nan")])
    t = torch.tensor([1, 0, 1, 0, 1, 0, 1, 0, 1, 0])
    out1 = _native.curve_evaluate(p.to(DEV), t.to(DEV), 1, 1, want_curve=True)
    out2 = _native.curve_evaluate(p.to(DEV), t.to(DEV), 1, 1, want_curve=True)
    assert torch.equal(out1[0], out2[0]) and torch.equal(out1[1], out2[1])
    u = int(out1[2][0, 2])
    thr = out1[3][2][0, :u].cpu()
    assert torch.isnan(thr[0]) and thr[1] == float("inf") and thr[-1] == float("-inf")  # NaN first, like torch.argsort(desc)
    assert u == 8  # {nan, inf, 2, 0.5 (x2), 1e-40, +-0 (one tie group), -1e-40, -inf}


@pytest.mark.parametrize("C,N,kind", MC_CASES)
def test_multiclass_auroc_ap_vs_golden(golden_curves, C, N, kind):
    fc = _fc()
    p, t = mc_inputs(C, N, kind)
    p, t = p.to(DEV), t.to(DEV)
    key = f"mc/C{C}_{kind}"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for avg in ("macro", "weighted", "none"):
            got = fc.multiclass_auroc(p, t, C, average=avg)
            np.testing.assert_allclose(got.cpu().numpy(), golden_curves[f"{key}/auroc_{avg}"], rtol=2e-6, atol=1e-7, equal_nan=True)
            got = fc.multiclass_average_precision(p, t, C, average=avg)
            np.testing.assert_allclose(got.cpu().numpy(), golden_curves[f"{key}/ap_{avg}"], rtol=2e-6, atol=1e-7, equal_nan=True)


def test_multiclass_curves_and_ignore_vs_golden(golden_curves):
    fc, g = _fc(), golden_curves
    for kind in ("probs", "logits"):
        p, t = mc_inputs(5, 400, kind)
        key = f"mc/C5_{kind}"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fpr, tpr, thr = fc.multiclass_roc(p.to(DEV), t.to(DEV), 5)
            pr, rc, th2 = fc.multiclass_precision_recall_curve(p.to(DEV), t.to(DEV), 5)
            tol = dict(rtol=RTOL, atol=1e-7, equal_nan=True)
            for c in range(5):
                np.testing.assert_allclose(fpr[c].cpu().numpy(), g[f"{key}/roc_fpr{c}"], **tol)
                np.testing.assert_allclose(tpr[c].cpu().numpy(), g[f"{key}/roc_tpr{c}"], **tol)
                np.testing.assert_allclose(thr[c].cpu().numpy(), g[f"{key}/roc_thr{c}"], **tol)
                np.testing.assert_allclose(pr[c].cpu().numpy(), g[f"{key}/prc_p{c}"], **tol)
                np.testing.assert_allclose(rc[c].cpu().numpy(), g[f"{key}/prc_r{c}"], **tol)
            t3 = torch.from_numpy(g[f"{key}/ignore_target"]).to(DEV)
            got = fc.multiclass_auroc(p.to(DEV), t3, 5, average="none", ignore_index=-1)
            np.testing.assert_allclose(got.cpu().numpy(), g[f"{key}/ignore_auroc"], rtol=2e-6, atol=1e-7)


def test_cfg5_single_rank_modular_vs_golden(golden_curves):
    from metrics_b200.classification import MulticlassAUROC, MulticlassAveragePrecision

    m_auc = MulticlassAUROC(num_classes=1000, validate_args=False).to(DEV)
    m_ap = MulticlassAveragePrecision(num_classes=1000, validate_args=False).to(DEV)
    for lg, tg in cfg5_rank_batches(0, 2):
        m_auc.update(lg.to(DEV), tg.to(DEV))
        m_ap.update(lg.to(DEV), tg.to(DEV))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_allclose(m_auc.compute().cpu().numpy(), golden_curves["mc/cfg5_rank0_2batches/auroc"], rtol=2e-6)
        np.testing.assert_allclose(m_ap.compute().cpu().numpy(), golden_curves["mc/cfg5_rank0_2batches/ap"], rtol=2e-6)


def _same_evaluation(a, b):
    assert torch.equal(a[2], b[2])
    assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32)) and torch.equal(a[1].view(torch.int32), b[1].view(torch.int32))
    for c in range(a[2].shape[0]):
        u = int(a[2][c, 2])
        for x, y in zip(a[3], b[3]):
            assert torch.equal(x[c, :u].view(torch.int32), y[c, :u].view(torch.int32))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_label_in_key_path_is_bit_identical_to_the_pair_path(dtype):
    """`mb200_curve_evaluate_nonneg` (non-negative or NaN scores: 31-bit keys, label in bit 0, 4-byte records) against the general
    (key, label) pair sort: every output bit-identical — binary and one-vs-rest, ties, NaN, +-0, exact 0 and 1, ragged sizes."""
    from metrics_b200 import _native

    g = torch.Generator().manual_seed(7)
    for n in (1, 2, 33, 2049, 12289, 300001):
        p = (torch.rand(n, generator=g) * 64).floor() / 64 if n % 2 else torch.rand(n, generator=g)
        p[::7] = 0.0
        p[3::11] = 1.0
        p[5::13] = float("nan")
        p[6::17] = -0.0
        t = torch.randint(0, 2, (n,), generator=g)
        pd, td = p.to(DEV).to(dtype), t.to(DEV)
        _same_evaluation(_native.curve_evaluate(pd, td, 1, 1, want_curve=True, unit_range=True),
                         _native.curve_evaluate(pd, td, 1, 1, want_curve=True, unit_range=False))
    for n, c in ((5, 3), (1000, 7), (4099, 33), (20000, 100)):
        p = torch.softmax(torch.randn(n, c, generator=g) * 3, 1)
        p[::5, 0] = float("nan")
        p[1::9] = 0.0
        t = torch.randint(0, c, (n,), generator=g)
        pd, td = p.to(DEV).to(dtype), t.to(DEV)
        _same_evaluation(_native.curve_evaluate(pd, td, c, want_curve=True, unit_range=True),
                         _native.curve_evaluate(pd, td, c, want_curve=True, unit_range=False))


@pytest.mark.raw_abi
def test_label_in_key_path_refuses_negative_scores():
    """Default mode (unit_range=None) speculates, sees MB200_FLAG_PREDS_RANGE and re-evaluates on the general path; the raw
    `_nonneg` entry raises the flag for a caller that broke its promise (and only then: 1.5 and +inf have 31-bit keys)."""
    from metrics_b200 import _native

    g = torch.Generator().manual_seed(8)
    n = 5000
    for bad in (-1e-30, -1e-42, 1.5, float("-inf"), float("inf")):
        p = torch.rand(n, generator=g)
        p[n // 2] = bad
        t = torch.randint(0, 2, (n,), generator=g)
        pd, td = p.to(DEV), t.to(DEV)
        _same_evaluation(_native.curve_evaluate(pd, td, 1, 1, want_curve=True),
                         _native.curve_evaluate(pd, td, 1, 1, want_curve=True, unit_range=False))
        lib = _native.lib()
        nbytes = int(lib.mb200_curve_workspace_bytes_for(1, n, _native.tag(pd)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
        out = [torch.empty(1, dtype=torch.float32, device=DEV), torch.empty(1, dtype=torch.float32, device=DEV),
               torch.empty((1, 3), dtype=torch.int64, device=DEV)]
        flag = torch.zeros(1, dtype=torch.int32, device=DEV)
        with _native.on_device(pd.device):
            rc = lib.mb200_curve_evaluate_nonneg(pd.data_ptr(), _native.tag(pd), td.data_ptr(), _native.tag(td), n, 1, 1, ws.data_ptr(),
                                               nbytes, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), None, None, None,
                                               flag.data_ptr(), _native.stream_handle(pd.device))
        assert rc == 0 and bool(int(flag) & _native.FLAG_PREDS_RANGE) == (bad < 0)


@pytest.mark.parametrize("unit", [True, False])
def test_chained_scan_long_tie_runs(unit):
    """Tie groups longer than a scan tile (4096): tiles without any group end must hand the previous TP/FP through — constant
    scores, two-valued scores with runs of 10^4, a single odd score in the middle, and the same per class."""
    from metrics_b200 import _native

    g = torch.Generator().manual_seed(21)
    cases = []
    for n in (4096, 4097, 20000, 70001):
        cases.append(torch.full((n,), 0.25))
        p = torch.full((n,), 0.5)
        p[n // 2:] = 0.125
        cases.append(p)
        p = torch.full((n,), 0.75)
        p[n // 3] = 0.5
        cases.append(p)
        p = (torch.rand(n, generator=g) * 3).floor() / 4
        cases.append(p)
    for p in cases:
        n = p.numel()
        t = torch.randint(0, 2, (n,), generator=g)
        auroc, ap, counts, (fps, tps, thr) = _native.curve_evaluate(p.to(DEV), t.to(DEV), 1, 1, want_curve=True, unit_range=unit)
        rf, rt, rth = oc.binary_clf_curve(p.numpy(), t.numpy())
        u = int(counts[0, 2])
        assert u == rf.size
        np.testing.assert_array_equal(fps[0, :u].cpu().numpy().astype(np.int64), rf)
        np.testing.assert_array_equal(tps[0, :u].cpu().numpy().astype(np.int64), rt)
        np.testing.assert_array_equal(thr[0, :u].cpu().numpy(), rth)
        np.testing.assert_allclose(float(auroc[0]), oc.binary_auroc_exact(p.numpy(), t.numpy()), rtol=2e-7, atol=1e-7)
        np.testing.assert_allclose(float(ap[0]), oc.binary_average_precision_exact(p.numpy(), t.numpy()), rtol=2e-7, atol=1e-7)
    # per class: [n, 3] with a constant column, a two-valued column and a random one
    n = 30011
    p = torch.stack([torch.full((n,), 0.3), (torch.arange(n) < n // 2).float() * 0.5, torch.rand(n, generator=g)], 1)
    t = torch.randint(0, 3, (n,), generator=g)
    auroc, ap, counts, (fps, tps, thr) = _native.curve_evaluate(p.to(DEV), t.to(DEV), 3, want_curve=True, unit_range=unit)
    for c in range(3):
        rf, rt, rth = oc.binary_clf_curve(p[:, c].numpy(), (t == c).long().numpy())
        u = int(counts[c, 2])
        assert u == rf.size
        np.testing.assert_array_equal(fps[c, :u].cpu().numpy().astype(np.int64), rf)
        np.testing.assert_array_equal(tps[c, :u].cpu().numpy().astype(np.int64), rt)
        np.testing.assert_allclose(float(auroc[c]), oc.binary_auroc_exact(p[:, c].numpy(), (t == c).long().numpy()), rtol=2e-7, atol=1e-7)


@pytest.mark.raw_abi
def test_evaluate_keys_label_in_key_path_matches_the_pair_path():
    """`mb200_curve_evaluate_keys_nonneg` (the class-sharded exchange of metric states) against `mb200_curve_evaluate_keys`:
    same packed keys, same targets -> bit-identical scalars and counts, for class offsets and ragged sizes."""
    from metrics_b200 import _native

    g = torch.Generator().manual_seed(77)
    for n, c, first in ((1, 2, 0), (4097, 5, 0), (20000, 7, 3), (333, 40, 10)):
        p = torch.softmax(torch.randn(n, c, generator=g) * 2, 1).to(DEV)
        p[::9, 0] = float("nan")
        t = torch.randint(0, c + first, (n,), generator=g).to(DEV)
        a = _native.curve_evaluate_keys(_native.curve_pack_keys(p), t, first)
        b = _native.curve_evaluate_keys(_native.curve_pack_keys(p), t, first, nonneg=True)
        for x, y in zip(a, b):
            assert torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x, y.view(torch.int32) if y.dtype == torch.float32 else y)
