"""GPU: collection-level fusion (csrc/fused.cu, K11).  `MetricCollection([MulticlassF1Score, MulticlassAUROC]).update` through
the one fused kernel must leave exactly the states the two members' own updates leave (which the reference goldens pin in
test_confmat_gpu.py / test_curves_gpu.py): integer tp/fp/tn/fn bit-equal, stored probabilities bit-equal to K6's (and thereby
to ATen's CUDA softmax, test_normalize_aten_gpu.py), results equal."""
import os

import pytest
import torch

from metrics_b200 import MetricCollection
from metrics_b200.classification import MulticlassAUROC, MulticlassAveragePrecision, MulticlassF1Score, MulticlassAccuracy

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _collection(c, average="macro"):
    return MetricCollection([MulticlassF1Score(num_classes=c, average=average, validate_args=False),
                             MulticlassAccuracy(num_classes=c, average=average, validate_args=False),
                             MulticlassAUROC(num_classes=c, validate_args=False),
                             MulticlassAveragePrecision(num_classes=c, validate_args=False)]).to(DEV)


def _run(c, batches, fused, average="macro"):
    os.environ["MB200_COLLECTION_FUSION"] = "1" if fused else "0"
    try:
        mc = _collection(c, average)
        for lg, tg in batches:
            mc.update(lg, tg)
        f1, auroc = mc["MulticlassF1Score"], mc["MulticlassAUROC"]
        states = [f1.tp.clone(), f1.fp.clone(), f1.tn.clone(), f1.fn.clone(), torch.cat(auroc.preds), torch.cat(auroc.target)]
        return states, {k: v.clone() for k, v in mc.compute().items()}, len(auroc.preds)
    finally:
        os.environ.pop("MB200_COLLECTION_FUSION", None)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("c,n", [(5, 300), (33, 1000), (1000, 4096), (1024, 257)])
@pytest.mark.parametrize("average", ["macro", "micro"])
def test_fused_update_equals_member_updates(dtype, c, n, average):
    g = torch.Generator().manual_seed(c * 7 + n)
    batches = []
    for b in range(3):
        lg = (torch.randn(n, c, generator=g) * 3)
        if b == 1:
            lg = (lg * 2).round() / 2  # many ties for the maximum
        batches.append((lg.to(dtype).to(DEV), torch.randint(0, c, (n,), generator=g).to(DEV)))
    s1, r1, k1 = _run(c, batches, True, average)
    s0, r0, k0 = _run(c, batches, False, average)
    assert k1 == k0 == 3
    for a, b in zip(s1, s0):
        assert a.dtype == b.dtype and torch.equal(a, b)
    for k in r0:
        assert torch.equal(r1[k].nan_to_num(7.0), r0[k].nan_to_num(7.0)), k


def test_probability_batches_are_kept_as_they_are():
    g = torch.Generator().manual_seed(3)
    probs = torch.softmax(torch.randn(500, 9, generator=g), 1).to(DEV)
    tgt = torch.randint(0, 9, (500,), generator=g).to(DEV)
    logits = torch.randn(500, 9, generator=g).to(DEV)
    s1, r1, _ = _run(9, [(logits, tgt), (probs, tgt), (logits, tgt)], True)
    s0, r0, _ = _run(9, [(logits, tgt), (probs, tgt), (logits, tgt)], False)
    for a, b in zip(s1, s0):
        assert torch.equal(a, b)
    assert torch.equal(s1[4][500:1000], probs)  # the vote said "not logits": the scores themselves were stored


def test_nan_and_infinite_rows_follow_argmax_semantics():
    c = 40
    lg = torch.randn(64, c)
    lg[0, 7] = float("nan")
    lg[1, [3, 9]] = float("nan")          # the first NaN wins
    lg[2] = float("-inf")                 # all equal: class 0
    lg[3, 5] = float("inf")
    lg[4, [11, 12]] = 4.0                 # tie: lowest index
    lg[5, 0], lg[5, 1] = 0.0, -0.0        # -0 == +0
    tgt = torch.randint(0, c, (64,))
    s1, _, _ = _run(c, [(lg.to(DEV), tgt.to(DEV))], True)
    s0, _, _ = _run(c, [(lg.to(DEV), tgt.to(DEV))], False)
    for a, b in zip(s1[:4], s0[:4]):
        assert torch.equal(a, b)
    assert torch.equal(s1[4].nan_to_num(7.0), s0[4].nan_to_num(7.0))


def test_fusion_steps_aside_when_it_does_not_apply():
    mc = MetricCollection([MulticlassF1Score(num_classes=4, validate_args=False, ignore_index=1),
                           MulticlassAUROC(num_classes=4, validate_args=False)]).to(DEV)
    lg, tg = torch.randn(50, 4, device=DEV), torch.randint(0, 4, (50,), device=DEV)
    mc.update(lg, tg)
    mc.update(lg, tg)
    assert mc._fusion_plan() is None and len(mc["MulticlassAUROC"].preds) == 2
