"""CPU: the reference CPU op chains used as timing baselines (oracle/torch_cpu_chain.py) reproduce the reference's results
(goldens from the unmodified reference), so the numbers timed beside the kernels are the right computation."""
import numpy as np
import torch

from oracle.torch_cpu_chain import (
    macro_accuracy_cpu,
    multiclass_auroc_compute_cpu,
    multiclass_confmat_update_cpu,
    multiclass_stat_scores_update_cpu,
)
from tests.helpers import cfg1_inputs


def test_cfg1_chain(golden_cls):
    p, t = cfg1_inputs()
    st = [torch.zeros(5, dtype=torch.long) for _ in range(4)]
    for i in range(100):
        multiclass_stat_scores_update_cpu(*st, p[i], t[i], 5)
    for name, got in zip(("tp", "fp", "tn", "fn"), st):
        np.testing.assert_array_equal(got.numpy(), golden_cls[f"cfg1/{name}"])
    np.testing.assert_allclose(float(macro_accuracy_cpu(*st)), golden_cls["cfg1/value"], rtol=1e-6)


def test_multiclass_auroc_chain(golden_curves):
    g = golden_curves
    lg, tg = torch.from_numpy(g["mc/C5_logits/preds"]), torch.from_numpy(g["mc/C5_logits/target"])
    got = multiclass_auroc_compute_cpu(torch.softmax(lg, 1), tg, 5)
    np.testing.assert_allclose(float(got), g["mc/C5_logits/auroc_macro"], rtol=1e-6)


def test_confmat_chain_small():
    g = torch.Generator().manual_seed(3)
    lg = torch.randn(512, 7, generator=g).bfloat16()
    tg = torch.randint(0, 7, (512,), generator=g)
    cm = torch.zeros(7, 7, dtype=torch.long)
    multiclass_confmat_update_cpu(cm, lg, tg, 7)
    pred = lg.float().argmax(1)
    exp = torch.zeros(7, 7, dtype=torch.long)
    for a, b in zip(tg.tolist(), pred.tolist()):
        exp[a, b] += 1
    assert torch.equal(cm, exp)


def test_exact_curve_chain_reproduces_the_reference_goldens(golden_fuzz2):
    """The device-agnostic restatement of the exact curve functionals (the arbiter of tests/test_zzz_fuzz2_gpu.py when
    baseline/_ref did not travel) against every matching case of the second fuzz draw, produced by the unmodified reference on
    CPU: same lengths, same values."""
    import json

    import numpy as np
    import torch

    from oracle.torch_cpu_chain import exact_curve_functional_chain
    from tests.fuzz_cases import _DT, _flatten, n_cases

    hits = 0
    for k in range(n_cases("fuzz2")):
        spec = json.loads(str(golden_fuzz2[f"{k}/spec"]))
        kw = spec["kwargs"]
        if spec["fn"] not in ("binary_roc", "binary_precision_recall_curve", "multiclass_roc", "multiclass_precision_recall_curve"):
            continue
        if kw.get("thresholds") is not None or kw.get("average") is not None or spec["preds_dtype"] in ("float16", "bfloat16"):
            continue
        preds = torch.from_numpy(golden_fuzz2[f"{k}/preds"]).to(_DT[spec["preds_dtype"]])
        target = torch.from_numpy(golden_fuzz2[f"{k}/target"])
        got = _flatten(exact_curve_functional_chain(spec["fn"], preds, target, num_classes=kw.get("num_classes"),
                                                    ignore_index=kw.get("ignore_index")))
        assert len(got) == int(golden_fuzz2[f"{k}/n_out"])
        for i, t in enumerate(got):
            exp = golden_fuzz2[f"{k}/out{i}"]
            assert tuple(t.shape) == exp.shape, (k, spec["fn"], i)
            np.testing.assert_allclose(t.numpy(), exp, rtol=1e-6, atol=1e-7, equal_nan=True)
        hits += 1
    assert hits >= 10
