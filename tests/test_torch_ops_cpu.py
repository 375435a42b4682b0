"""CPU: the operator binding (csrc/torch_ops/ops.cpp) is built in-tree, registers every schema the design names, has no CPU
dispatch (it must fail loudly there), and traces under fake tensors (shape propagation needs no device)."""
import pytest
import torch

from metrics_b200 import torch_ops

SCHEMAS = {
    "confmat_update_": "metrics_b200::confmat_update_(Tensor(a!) confmat, Tensor preds, Tensor target, int num_classes, int? ignore_index=None, Tensor? err_flag=None) -> ()",
    "stat_scores_update_": None, "stats_softmax_update_": None, "normalize_logits_if_needed": None, "curve_evaluate": None,
    "binned_curve_update_": None, "regression_sums": None,
}


def test_library_is_built_and_registers_the_operators():
    assert torch_ops.available(), "run `python -c 'import __graft_entry__ as g; g.build()'`"
    torch_ops.load()
    for name, schema in SCHEMAS.items():
        op = getattr(torch.ops.metrics_b200, name)
        if schema is not None:
            assert str(op.default._schema) == schema


def test_no_cpu_dispatch():
    torch_ops.load()
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.metrics_b200.confmat_update_(torch.zeros(3, 3, dtype=torch.long), torch.randn(4, 3), torch.tensor([0, 1, 2, 0]), 3)


def test_fake_tensors_propagate_shapes_without_a_device():
    from torch._subclasses.fake_tensor import FakeTensorMode

    torch_ops.load()
    with FakeTensorMode():
        p = torch.empty(1000, 7)
        t = torch.empty(1000, dtype=torch.long)
        out = torch.ops.metrics_b200.curve_evaluate(p, t, 7, 1, True)
        assert out[0].shape == (7,) and out[3].shape == (7, 1000) and out[2].dtype == torch.int64
        probs = torch.ops.metrics_b200.normalize_logits_if_needed(p, "softmax")
        assert probs.shape == p.shape
