"""GPU: arena-backed list states of the binary exact curve metrics (metrics_b200/utilities/arena.py).  `update` writes each
batch's formatted scores and targets straight into growing buffers (one launch), the list states stay lists of tensors with
the reference's content, and `compute` uses the buffers without concatenating.  Everything must equal the generic path
(separate format kernel + list append + cat), which the reference goldens pin in test_curves_gpu.py."""
import pytest
import torch

from metrics_b200 import MetricCollection
from metrics_b200.classification import BinaryAUROC, BinaryAveragePrecision, BinaryPrecisionRecallCurve, BinaryROC
from metrics_b200.utilities.arena import ArenaList

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batches(seed, sizes, dtype=torch.float32, logits=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in sizes:
        p = torch.randn(n, generator=g) * 3 if logits else torch.rand(n, generator=g)
        out.append((p.to(dtype).to(DEV), torch.randint(0, 2, (n,), generator=g).to(DEV)))
    return out


def _generic(metric_cls, batches, monkeypatch, **kw):
    m = metric_cls(validate_args=False, **kw).to(DEV)
    monkeypatch.setattr(m, "_append_to_arena", lambda p, t: False)
    for p, t in batches:
        m.update(p, t)
    return m


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.float64])
def test_arena_states_equal_generic_states(dtype, monkeypatch):
    sizes = [10, 1000, 70000, 3, 40000, 65536, 1, 20000] * 3  # small-kernel and large-kernel batches, several growths
    batches = _batches(1, sizes, dtype)
    a = BinaryAUROC(validate_args=False).to(DEV)
    for p, t in batches:
        a.update(p, t)
    g = _generic(BinaryAUROC, batches, monkeypatch)
    assert isinstance(a.preds, ArenaList) and a.preds.packed() is not None and len(a.preds) == len(g.preds) == len(sizes)
    for x, y in zip(a.preds, g.preds):
        assert x.dtype == y.dtype and torch.equal(x, y)
    for x, y in zip(a.target, g.target):
        assert torch.equal(x, y)
    assert a.preds.packed().data_ptr() == a.preds[0].data_ptr()  # compute() reads the buffer itself
    assert torch.equal(a.compute(), g.compute())


def test_collection_group_shares_the_arena_and_reset_starts_over(monkeypatch):
    mc = MetricCollection([BinaryAUROC(validate_args=False), BinaryAveragePrecision(validate_args=False),
                           BinaryROC(validate_args=False)]).to(DEV)
    ref = MetricCollection([BinaryAUROC(validate_args=False), BinaryAveragePrecision(validate_args=False),
                            BinaryROC(validate_args=False)], compute_groups=False).to(DEV)
    for m in ref.values(copy_state=False):
        monkeypatch.setattr(m, "_append_to_arena", lambda p, t: False)
    for epoch in range(2):
        batches = _batches(10 + epoch, [5000] * 40, logits=bool(epoch))
        for p, t in batches:
            mc.update(p, t)
            ref.update(p, t)
        got, want = mc.compute(), ref.compute()
        assert torch.equal(got["BinaryAUROC"], want["BinaryAUROC"]) and torch.equal(got["BinaryAveragePrecision"], want["BinaryAveragePrecision"])
        for x, y in zip(got["BinaryROC"], want["BinaryROC"]):
            assert torch.equal(x, y)
        leader = list(mc.values(copy_state=False))[0]
        assert len(mc.compute_groups) == 1  # equal states were recognised: only the leader updates
        assert all(m.preds is leader.preds for m in mc.values(copy_state=False))
        assert isinstance(leader.preds, ArenaList) and leader.preds.packed() is not None
        mc.reset()
        ref.reset()
        assert len(leader.preds) == 0 and leader.preds.buffer is None


def test_falls_back_when_the_list_is_changed_behind_its_back(monkeypatch):
    batches = _batches(3, [1000, 2000, 3000])
    m = BinaryPrecisionRecallCurve(validate_args=False).to(DEV)
    m.update(*batches[0])
    m.preds.append(torch.sigmoid(batches[1][0]))  # an outside append: the arena steps aside, results stay right
    m.target.append(batches[1][1])
    m.update(*batches[2])
    g = _generic(BinaryPrecisionRecallCurve, batches, monkeypatch)
    assert m.preds.packed() is None
    for x, y in zip(m.compute(), g.compute()):
        assert torch.equal(x, y)


def test_dtype_change_and_device_round_trip(monkeypatch):
    b32, b16 = _batches(4, [4096, 4096]), _batches(5, [4096], torch.float16)
    m = BinaryAUROC(validate_args=False).to(DEV)
    for p, t in b32 + b16:  # the float16 batch cannot live in the float32 buffer: generic append from then on
        m.update(p, t)
    assert len(m.preds) == 3 and m.preds[2].dtype == torch.float16
    m2 = BinaryAUROC(validate_args=False).to(DEV)
    for p, t in b32:
        m2.update(p, t)
    v = m2.compute()
    m2 = m2.to("cpu").to(DEV)  # the move rebuilds plain lists; values unchanged, later updates take the generic path
    m2.update(*b32[0])
    g = _generic(BinaryAUROC, b32 + [b32[0]], monkeypatch)
    assert torch.equal(m2.compute(), g.compute()) and torch.isfinite(v)


def test_forward_keeps_working():
    m = BinaryAUROC(validate_args=False).to(DEV)
    p, t = _batches(6, [5000])[0]
    batch_value = m(p, t)  # full-state forward: snapshot, batch update, restore
    m.update(p, t)
    assert torch.isfinite(batch_value) and torch.equal(m.compute(), batch_value)  # same data twice: same AUROC
