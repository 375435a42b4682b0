"""CPU: the Metric / MetricCollection runtime (API parity with reference metric.py / collections.py)."""
import pickle
from copy import deepcopy
from unittest.mock import Mock

import pytest
import torch

from metrics_b200 import CompositionalMetric, Metric, MetricCollection
from metrics_b200.utilities.exceptions import TorchMetricsUserError
from tests.dummies import DummyCat, DummyFullState, DummyIntStates, DummyKw, DummyMean, DummyNone, DummySum


def test_constructor_kwargs_are_validated():
    with pytest.raises(ValueError, match="Unexpected keyword arguments: `foo`"):
        DummySum(foo=1)
    with pytest.raises(ValueError, match="`compute_on_cpu` to be an `bool`"):
        DummySum(compute_on_cpu=None)
    with pytest.raises(ValueError, match="`dist_sync_fn` to be an callable"):
        DummySum(dist_sync_fn=3)
    m = DummySum(sync_on_compute=False, compute_with_cache=False, dist_sync_on_step=True)
    assert (m.sync_on_compute, m.compute_with_cache, m.dist_sync_on_step) == (False, False, True)


def test_add_state_rules():
    m = DummySum()
    with pytest.raises(ValueError, match="state variable must be a tensor or any empty list"):
        m.add_state("bad", [torch.tensor(1)], "sum")
    with pytest.raises(ValueError, match="state variable must be a tensor or any empty list"):
        m.add_state("bad", 3, "sum")
    with pytest.raises(ValueError, match="`dist_reduce_fx` must be callable or one of"):
        m.add_state("bad", torch.tensor(0), "xyz")
    m.add_state("ok", torch.tensor(1), lambda x: x.sum())
    assert callable(m._reductions["ok"]) and m._persistent["ok"] is False
    assert set(m.metric_state) == {"x", "ok"}


def test_update_compute_cache_reset():
    m = DummySum()
    with pytest.warns(UserWarning, match="was called before the ``update`` method"):
        m.compute()
    assert not m.update_called and m.update_count == 0
    m.update(2.0)
    m.update(3.0)
    assert m.update_called and m.update_count == 2
    out = m.compute()
    assert float(out) == 5.0 and m._computed is not None
    out += 100  # returned values never alias the state
    assert float(m.x) == 5.0
    m.update(1.0)
    assert m._computed is None and float(m.compute()) == 6.0
    m.reset()
    assert float(m.x) == 0.0 and m.update_count == 0 and m._computed is None
    nc = DummySum(compute_with_cache=False)
    nc.update(1.0)
    nc.compute()
    assert nc._computed is None


@pytest.mark.parametrize("cls", [DummySum, DummyFullState])
def test_forward_returns_batch_value_and_accumulates(cls):
    m = cls()
    assert float(m(2.0)) == 2.0
    assert float(m(5.0)) == 5.0
    assert float(m.compute()) == 7.0 and m.update_count == 2
    assert m._forward_cache is not None


def test_forward_with_list_state_and_mean_state():
    c = DummyCat()
    v, _ = c(torch.tensor([1.0, 2.0]))
    assert v.tolist() == [1.0, 2.0]
    v, _ = c(torch.tensor([3.0]))
    assert float(v) == 3.0  # one-element results are squeezed (`_squeeze_if_scalar`), like the reference
    assert c.compute()[0].tolist() == [1.0, 2.0, 3.0]
    mean = DummyMean()
    mean(2.0)
    mean(4.0)
    assert float(mean.m) == 3.0  # running mean merge rule


def test_forward_while_synced_is_an_error():
    m = DummySum()
    m._is_synced = True
    with pytest.raises(TorchMetricsUserError, match="shouldn't be synced when performing ``forward``"):
        m(1.0)


def test_merge_state():
    a, b = DummySum(), DummySum()
    a.update(1.0)
    b.update(2.0)
    a.merge_state(b)
    assert float(a.compute()) == 3.0
    a.merge_state({"x": torch.tensor(4.0)})
    assert float(a.x) == 7.0
    with pytest.raises(ValueError, match="Expected incoming state to be a dict or an instance of Metric"):
        a.merge_state(3)
    with pytest.raises(ValueError, match="Expected incoming state to be an instance of DummySum"):
        a.merge_state(DummyMean())
    with pytest.raises(RuntimeError, match="``merge_state`` is not supported"):
        DummyFullState().merge_state(DummyFullState())
    c1, c2 = DummyCat(), DummyCat()
    c1.update(torch.tensor([1.0]))
    c2.update(torch.tensor([2.0]))
    c1.merge_state(c2)
    assert c1.compute()[0].tolist() == [2.0, 1.0]  # incoming ("global") first, as in the reference


def test_sync_calls_custom_fn_once_per_state_and_unsync_restores():
    fn = Mock(side_effect=lambda t, group=None: [t, t + 1])
    m = DummyIntStates(n=3, dist_sync_fn=fn, distributed_available_fn=lambda: True)
    m.update([1, 2, 3])
    out = m.compute()
    assert fn.call_count == 6  # one call per state tensor
    assert out[0].tolist() == [3, 5, 7]  # x + (x + 1)
    assert out[4].tolist() == [2, 3, 4] and out[5].tolist() == [1, 2, 3]
    assert m.tp.tolist() == [1, 2, 3] and not m._is_synced  # local state restored by unsync
    m.sync(dist_sync_fn=fn, distributed_available=lambda: True)
    assert m._is_synced and m.tp.tolist() == [3, 5, 7]
    with pytest.raises(TorchMetricsUserError, match="already been synced"):
        m.sync(dist_sync_fn=fn, distributed_available=lambda: True)
    m.unsync()
    assert m.tp.tolist() == [1, 2, 3]
    with pytest.raises(TorchMetricsUserError, match="already been un-synced"):
        m.unsync()


def test_sync_list_and_none_states_with_custom_fn():
    fn = lambda t, group=None: [t, t * 10]  # noqa: E731
    c = DummyCat(dist_sync_fn=fn, distributed_available_fn=lambda: True)
    c.update(torch.tensor([1.0, 2.0]))
    c.update(torch.tensor([3.0]))
    vals, ids = c.compute()
    assert vals.tolist() == [1.0, 2.0, 3.0, 10.0, 20.0, 30.0]
    assert len(c.vals) == 2  # list state restored
    n = DummyNone(dist_sync_fn=fn, distributed_available_fn=lambda: True)
    n.update([1.0, 2.0])
    t, l = n.compute()
    assert t.shape == (2, 2) and t[1].tolist() == [10.0, 20.0]
    assert isinstance(l, list) and len(l) == 2


def test_state_dict_persistence_and_loading():
    m = DummySum()
    m.update(3.0)
    assert "x" not in m.state_dict()
    m.persistent(True)
    sd = m.state_dict(prefix="p.")
    assert float(sd["p.x"]) == 3.0
    fresh = DummySum()
    fresh.persistent(True)
    fresh.load_state_dict({"x": torch.tensor(9.0)})
    assert float(fresh.x) == 9.0
    c = DummyCat()
    c.persistent(True)
    c.update(torch.tensor([1.0]))
    assert isinstance(c.state_dict()["vals"], list)


def test_pickle_clone_deepcopy_hash():
    m = DummySum()
    m.update(2.0)
    for other in (pickle.loads(pickle.dumps(m)), m.clone(), deepcopy(m)):
        other.update(1.0)
        assert float(other.compute()) == 3.0 and float(m.x) == 2.0
    assert hash(m) != hash(m.clone())
    assert isinstance(hash(DummyCat()), int)


def test_const_attributes_and_iter():
    m = DummySum()
    for name in ("higher_is_better", "is_differentiable", "full_state_update"):
        with pytest.raises(RuntimeError, match=f"Can't change const `{name}`"):
            setattr(m, name, True)
    with pytest.raises(TypeError):
        iter(m)


def test_dtype_device_rules():
    m = DummySum()
    assert m.device == torch.device("cpu") and m.dtype == torch.float32
    m.double()
    m.half()
    assert m.x.dtype == torch.float32  # plain casts are no-ops by design
    m.set_dtype(torch.float64)
    assert m.x.dtype == torch.float64 and m.dtype == torch.float64
    m.reset()
    assert m.x.dtype == torch.float64  # defaults follow


def test_filter_kwargs_and_compositional_metrics():
    k = DummyKw()
    assert k._filter_kwargs(preds=1, target=2, other=3) == {"preds": 1, "target": 2}
    a, b = DummySum(), DummySum()
    comp = a + b * 2
    assert isinstance(comp, CompositionalMetric)
    comp.update(3.0)
    assert float(comp.compute()) == 9.0
    assert float((a - 1.0).compute()) == 2.0
    assert float(abs(-a).compute()) == 3.0
    assert bool((a == 3.0).compute()) and bool((a >= b).compute())
    assert float((a / 2).compute()) == 1.5 and float((2 ** a).compute()) == 8.0
    comp.reset()
    assert float(a.x) == 0.0 and float(b.x) == 0.0
    assert float(comp(1.0)) == 3.0  # forward fans out too


def test_collection_basics_prefix_postfix_and_kwargs_filtering():
    mc = MetricCollection({"s": DummySum(), "k": DummyKw()}, prefix="val_", postfix="_ep")
    with pytest.raises(TypeError):
        mc.update(1.0)  # DummyKw needs preds/target
    mc = MetricCollection([DummyKw()], prefix="val_")
    mc.update(preds=torch.tensor([1.0, 2.0]), target=torch.tensor([0.0, 0.0]), extra=5)
    assert float(mc.compute()["val_DummyKw"]) == 3.0
    assert list(mc.keys()) == ["val_DummyKw"] and list(mc.keys(keep_base=True)) == ["DummyKw"]
    assert isinstance(mc["val_DummyKw"], DummyKw)
    clone = mc.clone(prefix="test_")
    assert list(clone.keys()) == ["test_DummyKw"]
    with pytest.raises(ValueError, match="Encountered two metrics both named DummySum"):
        MetricCollection([DummySum(), DummySum()])
    with pytest.raises(ValueError, match="Expected input `prefix` to be a string"):
        MetricCollection([DummySum()], prefix=3)
    with pytest.raises(ValueError, match="Unknown input to MetricCollection"):
        MetricCollection(5)


def test_collection_compute_groups_share_state_and_update_once():
    class A(DummyIntStates):
        pass

    class B(DummyIntStates):
        def compute(self):
            return self.tp.sum()

    class C(DummySum):
        def update(self, x):
            self.x += torch.as_tensor(x, dtype=self.x.dtype).sum()

    mc = MetricCollection([A(n=3), B(n=3), C()])
    assert mc.compute_groups == {0: ["A"], 1: ["B"], 2: ["C"]}
    calls = {"B": 0}
    orig = mc["B"].update

    def counted(*a, **k):
        calls["B"] += 1
        return orig(*a, **k)

    mc._modules["B"].update = counted
    mc.update(torch.tensor([1, 2, 3]))
    assert mc.compute_groups == {0: ["A", "B"], 1: ["C"]}
    mc.update(torch.tensor([1, 1, 1]))
    mc.update(torch.tensor([1, 1, 1]))
    assert calls["B"] == 1  # only the group leader updates after the groups are formed
    res = mc.compute()
    assert int(res["B"]) == 12 and res["A"][0].tolist() == [3, 4, 5]
    assert mc._modules["B"].update_count == 3
    # reading a member copies its state; the next update re-links it
    member = mc["B"]
    member.tp += 100
    assert int(mc._modules["A"].tp.sum()) == 12
    mc.update(torch.tensor([0, 0, 0]))
    assert int(mc.compute()["B"]) == 12
    mc.reset()
    assert int(mc._modules["B"].tp.sum()) == 0
    off = MetricCollection([A(n=3), B(n=3)], compute_groups=False)
    off.update(torch.tensor([1, 2, 3]))
    assert off.compute_groups == {}
    fixed = MetricCollection([A(n=3), B(n=3)], compute_groups=[["A", "B"]])
    fixed.update(torch.tensor([1, 2, 3]))
    assert int(fixed.compute()["B"]) == 6
    with pytest.raises(ValueError, match="does not match a metric in the collection"):
        MetricCollection([A(n=3)], compute_groups=[["nope"]])


def test_collection_nested_and_dict_results():
    class D(DummySum):
        def compute(self):
            return {"a": self.x, "b": self.x * 2}

    inner = MetricCollection([DummySum()], prefix="in_")
    mc = MetricCollection([D(), inner])
    mc.update(2.0)
    res = mc.compute()
    assert set(res) == {"a", "b", "in_DummySum"}
    assert float(res["b"]) == 4.0
    fwd = mc(1.0)
    assert float(fwd["a"]) == 1.0
    assert set(mc.metric_state) == {"D", "in_DummySum"}  # nested members are registered under their renamed key


def test_collection_can_be_scripted():
    """`torch.jit.script(MetricCollection(...))` works (reference tests/unittests/bases/test_collections.py:42-50): string
    class-level annotations on the collection would make TorchScript's annotation resolution fail."""
    import warnings

    from metrics_b200 import MetricCollection
    from tests.dummies import DummyMean, DummySum

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scripted = torch.jit.script(MetricCollection({"a": DummySum(), "b": DummyMean()}))
    assert scripted is not None


def test_every_metric_class_can_be_scripted():
    """The reference's MetricTester scripts every metric it tests (tests/unittests/_helpers/testers.py:144); bare class-level
    annotations turned into strings by `from __future__ import annotations` break TorchScript's type resolution."""
    import warnings

    import metrics_b200.classification as TC
    import metrics_b200.regression as TR
    from metrics_b200.detection import MeanAveragePrecision

    floors = ("AtFixed", "SensitivityAt", "SpecificityAt")
    made = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name in dir(TC):
            cls = getattr(TC, name)
            prefix = next((p for p in ("Binary", "Multiclass", "Multilabel") if name.startswith(p)), None)
            if not isinstance(cls, type) or prefix is None:
                continue
            args = () if prefix == "Binary" else (3,)
            if "FBeta" in name:
                args = (2.0,) + args
            if any(f in name for f in floors):
                args = args + (0.5,)
            if name in ("BinaryFairness", "BinaryGroupStatRates"):
                args = (2,)
            torch.jit.script(cls(*args))
            made += 1
        for name in dir(TR):
            cls = getattr(TR, name)
            if isinstance(cls, type) and issubclass(cls, torch.nn.Module) and cls.__module__.startswith("metrics_b200.regression"):
                special = {"MinkowskiDistance": {"p": 2.0}, "CriticalSuccessIndex": {"threshold": 0.5}}
                torch.jit.script(cls(**special.get(name, {})))
                made += 1
        torch.jit.script(MeanAveragePrecision())
    assert made >= 75  # 67 classification classes + 11 regression metrics at the time of writing


def test_classwise_wrapper_names_results_and_is_transparent_to_compute_groups():
    """Reference wrappers/classwise.py:32-237 and tests/unittests/bases/test_collections.py:574-604."""
    from metrics_b200 import ClasswiseWrapper

    class PerClass(DummyIntStates):
        def compute(self):
            return self.tp

    class PerClassTwice(DummyIntStates):
        def compute(self):
            return 2 * self.tp

    w = ClasswiseWrapper(PerClass(n=3))
    assert list(w(torch.tensor([1, 2, 3]))) == ["perclass_0", "perclass_1", "perclass_2"]
    w.update(torch.tensor([1, 1, 1]))
    assert [int(v) for v in w.compute().values()] == [2, 3, 4]
    assert w.tp is w.metric.tp  # states are the wrapped metric's
    w.reset()
    assert int(w.metric.tp.sum()) == 0
    named = ClasswiseWrapper(PerClass(n=2), labels=["tree", "bush"], prefix="f_", postfix="_x")
    assert list(named(torch.tensor([1, 2]))) == ["f_tree_x", "f_bush_x"]
    for bad in ({"metric": 3}, {"metric": PerClass(), "labels": "ab"}, {"metric": PerClass(), "prefix": 1},
                {"metric": PerClass(), "postfix": 1}):
        with pytest.raises(ValueError, match="Expected argument"):
            ClasswiseWrapper(**bad)

    members = {"a": ClasswiseWrapper(PerClass(n=3), prefix="a"), "b": ClasswiseWrapper(PerClassTwice(n=3), prefix="b")}
    mc = MetricCollection(members, compute_groups=[["a", "b"]], prefix="val/")
    assert mc.compute_groups == {0: ["a", "b"]}
    mc.update(torch.tensor([1, 2, 3]))
    mc.update(torch.tensor([1, 2, 3]))
    res = mc.compute()
    assert {k: int(v) for k, v in res.items()} == {"val/a0": 2, "val/a1": 4, "val/a2": 6, "val/b0": 4, "val/b1": 8, "val/b2": 12}
    auto = MetricCollection({"a": ClasswiseWrapper(PerClass(n=3), prefix="a"), "b": ClasswiseWrapper(PerClassTwice(n=3), prefix="b")})
    auto.update(torch.tensor([1, 2, 3]))
    assert auto.compute_groups == {0: ["a", "b"]}
    assert int(auto(torch.tensor([1, 1, 1]))["b2"]) == 2


def test_compute_groups_compare_states_across_dtypes():
    """Two members whose same-named states differ in dtype (long vs float32) must not raise when the collection looks for
    compute groups: the comparison casts like the reference's `allclose` helper (utilities/data.py:242-246)."""
    import torch

    from metrics_b200 import Metric, MetricCollection

    class CountLong(Metric):
        full_state_update = False

        def __init__(self, **kw):
            super().__init__(**kw)
            self.add_state("total", torch.tensor(0), dist_reduce_fx="sum")

        def update(self, x):
            self.total += int(x.numel())

        def compute(self):
            return self.total

    class CountFloat(CountLong):
        def __init__(self, **kw):
            Metric.__init__(self, **kw)
            self.add_state("total", torch.tensor(0.0), dist_reduce_fx="sum")

    mc = MetricCollection({"a": CountLong(), "b": CountFloat()})
    mc.update(torch.zeros(5))
    mc.update(torch.zeros(3))
    out = mc.compute()
    assert int(out["a"]) == 8 and float(out["b"]) == 8.0


def test_compute_on_cpu_parks_lists_and_stages_them_for_compute():
    """`compute_on_cpu=True`: list states live in host memory between updates; `compute` sees them on the metric's device
    (here a fake "device" = the meta of the metric's own device attribute: the staging logic is device-agnostic) and leaves
    them parked afterwards."""
    import torch

    from metrics_b200 import Metric

    seen = {}

    class Cat(Metric):
        full_state_update = False

        def __init__(self, **kw):
            super().__init__(**kw)
            self.add_state("xs", [], dist_reduce_fx="cat")

        def update(self, x):
            self.xs.append(x)

        def compute(self):
            seen["devices"] = {v.device.type for v in self.xs}
            return torch.cat(self.xs).sum()

    m = Cat(compute_on_cpu=True)
    m.update(torch.ones(3))
    m.update(torch.ones(2))
    assert all(v.device.type == "cpu" for v in m.xs)
    m._device = torch.device("meta")  # pretend the metric lives elsewhere: compute must stage the parked lists there
    try:
        m.compute()
    except (RuntimeError, NotImplementedError):
        pass  # meta tensors cannot be summed to a value; what matters is where compute saw them
    assert seen["devices"] == {"meta"}
    assert all(v.device.type == "cpu" for v in m.xs)  # parked again
