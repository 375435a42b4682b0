"""Pure-torch metrics used to exercise the runtime on CPU (no kernels involved)."""
import torch

from metrics_b200 import Metric


class DummySum(Metric):
    full_state_update = False

    def __init__(self, **kw):
        super().__init__(**kw)
        self.add_state("x", torch.tensor(0.0), dist_reduce_fx="sum")

    def update(self, x):
        self.x += torch.as_tensor(x, dtype=self.x.dtype)

    def compute(self):
        return self.x


class DummyFullState(DummySum):
    full_state_update = True


class DummyIntStates(Metric):
    """tp/fp/tn/fn-like integer vector states (bucketed all-reduce path)."""

    full_state_update = False

    def __init__(self, n=4, **kw):
        super().__init__(**kw)
        for name in ("tp", "fp", "tn", "fn"):
            self.add_state(name, torch.zeros(n, dtype=torch.long), dist_reduce_fx="sum")
        self.add_state("hi", torch.zeros(n, dtype=torch.long), dist_reduce_fx="max")
        self.add_state("lo", torch.full((n,), 10**6, dtype=torch.long), dist_reduce_fx="min")

    def update(self, v):
        v = torch.as_tensor(v, dtype=torch.long)
        self.tp += v
        self.fp += 2 * v
        self.tn += 3 * v
        self.fn += 4 * v
        self.hi = torch.maximum(self.hi, v)
        self.lo = torch.minimum(self.lo, v)

    def compute(self):
        return torch.stack([self.tp, self.fp, self.tn, self.fn, self.hi, self.lo])


class DummyCat(Metric):
    full_state_update = False

    def __init__(self, **kw):
        super().__init__(**kw)
        self.add_state("vals", [], dist_reduce_fx="cat")
        self.add_state("ids", [], dist_reduce_fx="cat")

    def update(self, x, ids=None):
        x = torch.as_tensor(x)
        self.vals.append(x)
        self.ids.append(torch.arange(x.shape[0]) if ids is None else torch.as_tensor(ids))

    def compute(self):
        from metrics_b200.utilities.data import dim_zero_cat

        return dim_zero_cat(self.vals), dim_zero_cat(self.ids)


class DummyMean(Metric):
    full_state_update = False

    def __init__(self, **kw):
        super().__init__(**kw)
        self.add_state("m", torch.tensor(0.0), dist_reduce_fx="mean")

    def update(self, x):
        self.m = torch.as_tensor(x, dtype=torch.float32)

    def compute(self):
        return self.m


class DummyNone(Metric):
    full_state_update = False

    def __init__(self, **kw):
        super().__init__(**kw)
        self.add_state("t", torch.tensor([0.0, 0.0]), dist_reduce_fx=None)
        self.add_state("l", [], dist_reduce_fx=None)

    def update(self, x):
        x = torch.as_tensor(x, dtype=torch.float32)
        self.t = self.t + x
        self.l.append(x.clone())

    def compute(self):
        return self.t, self.l


class DummyKw(Metric):
    full_state_update = False

    def __init__(self, **kw):
        super().__init__(**kw)
        self.add_state("s", torch.tensor(0.0), dist_reduce_fx="sum")

    def update(self, preds, target, weight=None):
        self.s += (preds - target).abs().sum() * (1.0 if weight is None else weight)

    def compute(self):
        return self.s
