"""GPU: `compute_on_cpu=True` with kernel-backed list-state metrics.  The kwarg parks list states in host memory after every
update (reference metric.py:478-479); there is no CPU arithmetic in this package, so `compute` stages them back on the metric's
device for the evaluation only.  Round 1 accepted the kwarg and then failed at `compute()` on every curve metric."""
import pytest
import torch

from metrics_b200.classification import BinaryAUROC, MulticlassAUROC

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("cls,kw,c", [(BinaryAUROC, {}, None), (MulticlassAUROC, {"num_classes": 7}, 7)])
def test_compute_on_cpu_matches_the_default(cls, kw, c):
    g = torch.Generator().manual_seed(5)
    parked, plain = cls(validate_args=False, compute_on_cpu=True, **kw).to(DEV), cls(validate_args=False, **kw).to(DEV)
    for _ in range(4):
        p = torch.randn(500, c, generator=g) if c else torch.randn(500, generator=g)
        t = torch.randint(0, c or 2, (500,), generator=g)
        parked.update(p.to(DEV), t.to(DEV))
        plain.update(p.to(DEV), t.to(DEV))
    assert all(v.device.type == "cpu" for v in parked.preds)  # parked between updates
    assert torch.equal(parked.compute(), plain.compute())
    assert all(v.device.type == "cpu" for v in parked.preds)  # and again after the evaluation
    parked.update(p.to(DEV), t.to(DEV))
    plain.update(p.to(DEV), t.to(DEV))
    assert torch.equal(parked.compute(), plain.compute())
