"""CPU: the operator-level helpers user code imports from the reference's `utilities` package
(reference tests/unittests/utilities/test_utilities.py is the model for the cases)."""
import pytest
import torch

from metrics_b200 import Metric
from metrics_b200.utilities import check_forward_full_state_property, class_reduce, reduce
from metrics_b200.utilities.checks import _allclose_recursive, is_overridden
from metrics_b200.utilities.compute import _auc_compute, auc, normalize_logits_if_needed
from metrics_b200.utilities.data import (
    _bincount,
    _cumsum,
    _flatten,
    _flatten_dict,
    _flexible_bincount,
    allclose,
    select_topk,
    to_categorical,
    to_onehot,
)
from metrics_b200._native import NativeLibraryError
from tests.dummies import DummySum


def test_reduce_and_class_reduce():
    x = torch.rand(5, 4, 3)
    assert torch.allclose(reduce(x, "elementwise_mean"), x.mean())
    assert torch.allclose(reduce(x, "sum"), x.sum())
    assert reduce(x, "none") is x and reduce(x, None) is x
    with pytest.raises(ValueError, match="Reduction parameter unknown."):
        reduce(x, "error_reduction")
    num, denom, w = torch.tensor([1.0, 0.0, 3.0]), torch.tensor([2.0, 0.0, 4.0]), torch.tensor([2.0, 0.0, 4.0])
    assert torch.allclose(class_reduce(num, denom, w, "micro"), torch.tensor(4 / 6))
    assert torch.allclose(class_reduce(num, denom, w, "macro"), torch.tensor((0.5 + 0 + 0.75) / 3))
    assert torch.allclose(class_reduce(num, denom, w, "weighted"), torch.tensor(0.5 * 2 / 6 + 0.75 * 4 / 6))
    assert torch.allclose(class_reduce(num, denom, w, "none"), torch.tensor([0.5, 0.0, 0.75]))
    assert torch.allclose(class_reduce(num, denom, w, None), torch.tensor([0.5, 0.0, 0.75]))
    with pytest.raises(ValueError, match="Reduction parameter nope unknown"):
        class_reduce(num, denom, w, "nope")


def test_onehot_topk_categorical():
    labels = torch.tensor([[0, 1, 2, 3, 4], [5, 6, 7, 8, 9]])
    oh = to_onehot(labels)
    assert oh.shape == (2, 10, 5) and oh.dtype == labels.dtype
    assert torch.equal(oh.argmax(1), labels) and int(oh.sum()) == 10
    assert to_onehot(torch.tensor([1, 2, 3])).tolist() == [[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    assert to_onehot(torch.tensor([1, 0]), num_classes=3).shape == (2, 3)
    x = torch.tensor([[1.1, 2.0, 3.0], [2.0, 1.0, 0.5]])
    assert select_topk(x, topk=2).tolist() == [[0, 1, 1], [1, 1, 0]] and select_topk(x, 2).dtype == torch.int32
    assert select_topk(x, topk=1).tolist() == [[0, 0, 1], [1, 0, 0]]
    assert select_topk(x.half(), topk=2).tolist() == [[0, 1, 1], [1, 1, 0]]
    assert select_topk(x.T, topk=2, dim=0).T.tolist() == [[0, 1, 1], [1, 1, 0]]
    probs = torch.tensor([[0.2, 0.5], [0.9, 0.1]])
    assert to_categorical(probs).tolist() == [1, 0]


def test_bincount_cumsum_flatten_allclose():
    x = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2, 2])
    assert _bincount(x, minlength=3).tolist() == [3, 2, 4]
    assert _bincount(x).tolist() == [3, 2, 4]
    assert _bincount(x, minlength=5).tolist() == [3, 2, 4, 0, 0]
    big = torch.randint(100, (1000,))
    assert torch.equal(_bincount(big, minlength=100), torch.bincount(big, minlength=100))
    torch.use_deterministic_algorithms(True)
    try:
        assert torch.equal(_bincount(big, minlength=100), torch.bincount(big, minlength=100))
    finally:
        torch.use_deterministic_algorithms(False)
    assert _flexible_bincount(torch.tensor([7, 7, -2, 40, 40, 40])).tolist() == [1, 2, 3]
    assert _cumsum(torch.arange(5), dim=0).tolist() == [0, 1, 3, 6, 10]
    assert _cumsum(torch.ones(3, dtype=torch.int32), dtype=torch.float64).dtype == torch.float64
    assert _flatten([[1, 2, 3], [4, 5]]) == [1, 2, 3, 4, 5]
    assert _flatten_dict({"a": {"b": 1, "c": 2}, "d": 3}) == ({"b": 1, "c": 2, "d": 3}, False)
    assert _flatten_dict({"a": {"b": 1, "c": 2}, "b": 3}) == ({"b": 3, "c": 2}, True)
    assert allclose(torch.ones(3), torch.ones(3, dtype=torch.float64))
    assert _allclose_recursive({"a": [torch.ones(2), 1.0]}, {"a": [torch.ones(2), 1.0]})


@pytest.mark.parametrize(
    ("x", "y", "expected"),
    [([0, 1], [0, 1], 0.5), ([1, 0], [0, 1], 0.5), ([1, 0, 0], [0, 1, 1], 0.5), ([0, 1], [1, 1], 1.0),
     ([0, 0.5, 1], [0, 0.5, 1], 0.5)],
)
def test_auc(x, y, expected):
    x, y = torch.tensor(x, dtype=torch.float32), torch.tensor(y, dtype=torch.float32)
    assert float(auc(x, y)) == pytest.approx(expected)
    assert float(auc(x.flip(0), y.flip(0))) == pytest.approx(expected)
    assert float(auc(x[None], y[None], reorder=True)) == pytest.approx(expected)


def test_auc_rejects_bad_inputs():
    with pytest.raises(ValueError, match="neither increasing or decreasing"):
        _auc_compute(torch.tensor([0.0, 2.0, 1.0]), torch.tensor([0.0, 1.0, 2.0]))
    assert float(_auc_compute(torch.tensor([0.0, 2.0, 1.0]), torch.tensor([0.0, 2.0, 1.0]), reorder=True)) == pytest.approx(2.0)
    with pytest.raises(ValueError, match="to be 1d"):
        auc(torch.rand(2, 3), torch.rand(2, 3))
    with pytest.raises(ValueError, match="same number of elements"):
        auc(torch.rand(3), torch.rand(4))


def test_normalize_logits_has_no_cpu_path():
    with pytest.raises(NativeLibraryError):
        normalize_logits_if_needed(torch.tensor([-1.0, 0.0, 1.0]), "sigmoid")
    with pytest.raises(ValueError, match="sigmoid"):
        normalize_logits_if_needed(torch.tensor([-1.0]), "tanh")


def test_is_overridden():
    class Child(DummySum):
        def update(self, x):
            super().update(x)

    assert is_overridden("update", Child(), DummySum) and not is_overridden("compute", Child(), DummySum)
    assert is_overridden("update", DummySum(), Metric)
    assert not is_overridden("nope", DummySum(), Metric)
    with pytest.raises(ValueError, match="parent should define"):
        is_overridden("x", type("T", (), {"x": lambda self: 0})(), Metric)


def test_check_forward_full_state_property(capsys):
    class Independent(DummySum):
        pass

    class Dependent(DummySum):
        def update(self, x):
            super().update(x)
            if self.x > 3:  # later states depend on earlier ones
                self.reset()

    check_forward_full_state_property(Independent, input_args={"x": 1.0}, num_update_to_compare=[5, 10], reps=2)
    out = capsys.readouterr().out
    assert "Full state for 5 steps took" in out and "Partial state for 10 steps took" in out
    assert "Recommended setting `full_state_update=" in out
    check_forward_full_state_property(Dependent, input_args={"x": 1.0}, num_update_to_compare=[10], reps=1)
    assert capsys.readouterr().out.strip() == "Recommended setting `full_state_update=True`"
