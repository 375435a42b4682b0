"""GPU: the second binding of the C-ABI — `torch.ops.metrics_b200.*` (csrc/torch_ops/ops.cpp, TORCH_LIBRARY) — gives exactly
the results of the ctypes binding, raises dispatcher-style errors, and traces under fake tensors / torch.compile."""
import pytest
import torch

from metrics_b200 import _native, torch_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    torch_ops.load()
    return torch.ops.metrics_b200


def test_confmat_and_stat_scores_ops_match_ctypes(ops):
    from metrics_b200.functional.classification.stat_scores import stat_scores_workspace

    g = torch.Generator().manual_seed(0)
    lg = torch.randn(4096, 1000, generator=g).bfloat16().to(DEV)
    tg = torch.randint(0, 1000, (4096,), generator=g).to(DEV)
    a, b = (torch.zeros(1000, 1000, dtype=torch.long, device=DEV) for _ in range(2))
    _native.multiclass_confmat_update_(a, lg, tg, 1000, None)
    ops.confmat_update_(b, lg, tg, 1000)
    assert torch.equal(a, b) and int(b.sum()) == 4096
    ops.confmat_update_(b, lg, tg, 1000, 5)  # ignore_index
    sa = [torch.zeros(1000, dtype=torch.long, device=DEV) for _ in range(4)]
    sb = [torch.zeros(1000, dtype=torch.long, device=DEV) for _ in range(4)]
    _native.multiclass_stat_scores_update_(*sa, stat_scores_workspace(1000, torch.device(DEV)), lg, tg, 1000, None, False)
    ops.stat_scores_update_(*sb, stat_scores_workspace(1000, torch.device(DEV)), lg, tg, 1000)
    assert all(torch.equal(x, y) for x, y in zip(sa, sb))


def test_curve_and_regression_ops_match_ctypes(ops):
    g = torch.Generator().manual_seed(1)
    p = torch.rand(100_000, generator=g).to(DEV)
    t = torch.randint(0, 2, (100_000,), generator=g).to(DEV)
    auroc, ap, counts, curve = _native.curve_evaluate(p, t, 1, 1, want_curve=True)
    o = ops.curve_evaluate(p, t, 1, 1, True)
    u = int(counts[0, 2])
    assert torch.equal(o[0], auroc) and torch.equal(o[1], ap) and torch.equal(o[2], counts)
    assert all(torch.equal(x[0, :u], y[0, :u]) for x, y in zip(o[3:], curve))
    q = torch.rand(100_000, generator=g).to(DEV)
    assert torch.equal(ops.regression_sums(p, q, 0), _native.regression_sums(p, q, 0))
    x = (torch.randn(64, 10, generator=g) * 3).to(DEV)
    assert torch.equal(ops.normalize_logits_if_needed(x, "softmax"), _native.softmax_if_logits(x))
    assert torch.equal(ops.normalize_logits_if_needed(x[:, 0].contiguous(), "sigmoid"), _native.sigmoid_if_logits(x[:, 0].contiguous()))


def test_metric_classes_run_on_the_operator_binding(ops, monkeypatch):
    from metrics_b200 import MetricCollection
    from metrics_b200.classification import MulticlassAUROC, MulticlassConfusionMatrix, MulticlassF1Score
    from metrics_b200.regression import MeanSquaredError

    g = torch.Generator().manual_seed(2)
    lg = torch.randn(2048, 37, generator=g).to(DEV)
    tg = torch.randint(0, 37, (2048,), generator=g).to(DEV)

    def run():
        mc = MetricCollection([MulticlassConfusionMatrix(37, validate_args=False), MulticlassF1Score(37, validate_args=False),
                               MulticlassAUROC(37, validate_args=False)]).to(DEV)
        mc.update(lg, tg)
        mc.update(lg, tg)
        mse = MeanSquaredError().to(DEV)
        mse.update(lg[:, 0], lg[:, 1])
        return {**mc.compute(), "mse": mse.compute()}

    want = run()
    monkeypatch.setattr(_native, "_TORCH_BINDING", True)
    got = run()
    assert want.keys() == got.keys() and all(torch.equal(want[k], got[k]) for k in want)


def test_dispatcher_errors_and_cpu_tensors(ops):
    with pytest.raises(RuntimeError):  # no CPU dispatch key registered
        ops.confmat_update_(torch.zeros(3, 3, dtype=torch.long), torch.randn(4, 3), torch.tensor([0, 1, 2, 0]), 3)
    with pytest.raises(RuntimeError, match="contiguous int64"):
        ops.confmat_update_(torch.zeros(3, 3, device=DEV), torch.randn(4, 3, device=DEV), torch.tensor([0, 1, 2, 0], device=DEV), 3)


def test_fake_tensor_tracing_and_compile(ops):
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():
        p = torch.empty(1000, device=DEV)
        t = torch.empty(1000, dtype=torch.long, device=DEV)
        out = ops.curve_evaluate(p, t, 1, 1, True)
        assert out[0].shape == (1,) and out[3].shape == (1, 1000) and out[2].dtype == torch.int64

    def step(confmat, lg, tg):
        torch.ops.metrics_b200.confmat_update_(confmat, lg, tg, 10)
        return confmat.sum()

    cm = torch.zeros(10, 10, dtype=torch.long, device=DEV)
    lg, tg = torch.randn(256, 10, device=DEV), torch.randint(0, 10, (256,), device=DEV)
    total = torch.compile(step, fullgraph=True, backend="eager")(cm, lg, tg)
    assert int(total) == 256


def test_opcheck_schema_and_fake_registration(ops):
    p = torch.rand(512, device=DEV)
    t = torch.randint(0, 2, (512,), device=DEV)
    torch.library.opcheck(ops.curve_evaluate.default, (p, t, 1, 1, False), test_utils=("test_schema", "test_faketensor"))
    torch.library.opcheck(ops.regression_sums.default, (p, p.flip(0), 0), test_utils=("test_schema", "test_faketensor"))
