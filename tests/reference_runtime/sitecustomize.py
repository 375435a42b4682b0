"""Imported by every interpreter started with this directory on PYTHONPATH (see run.sh): aliases `torchmetrics` to
`metrics_b200` so that the REFERENCE's own runtime tests exercise our Metric / MetricCollection runtime; classes outside the
scope that those test modules import are stubbed."""
import importlib
import pkgutil
import sys
import types

import os

_TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(_TESTS))  # repo root: metrics_b200
sys.path.insert(0, os.path.join(_TESTS, "golden", "_standins"))  # lightning_utilities stand-in used by the reference's helpers
import torch  # noqa: E402

import metrics_b200  # noqa: E402
from metrics_b200 import Metric  # noqa: E402

sys.modules["torchmetrics"] = metrics_b200
for m in pkgutil.walk_packages(metrics_b200.__path__, "metrics_b200."):
    try:
        mod = importlib.import_module(m.name)
    except Exception:
        continue
    sys.modules["torchmetrics." + m.name[len("metrics_b200."):]] = mod


class SumMetric(Metric):
    full_state_update = False
    def __init__(self, **kw):
        super().__init__(**kw)
        self.add_state("sum_value", torch.tensor(0.0), dist_reduce_fx="sum")

    def update(self, value):
        self.sum_value = self.sum_value + torch.as_tensor(value, dtype=torch.float32).sum()

    def compute(self):
        return self.sum_value


class MeanMetric(Metric):
    full_state_update = False
    def __init__(self, **kw):
        super().__init__(**kw)
        self.add_state("mean_value", torch.tensor(0.0), dist_reduce_fx="sum")
        self.add_state("weight", torch.tensor(0.0), dist_reduce_fx="sum")

    def update(self, value, weight=1.0):
        value = torch.as_tensor(value, dtype=torch.float32)
        weight = torch.broadcast_to(torch.as_tensor(weight, dtype=torch.float32), value.shape)
        self.mean_value = self.mean_value + (value * weight).sum()
        self.weight = self.weight + weight.sum()

    def compute(self):
        return self.mean_value / self.weight


class _Absent(Metric):
    def __init__(self, *a, **k):
        raise NotImplementedError("not part of metrics_b200's scope")

    def update(self):
        pass

    def compute(self):
        pass


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_mod("torchmetrics.aggregation", SumMetric=SumMetric, MeanMetric=MeanMetric)
_mod("torchmetrics.clustering", AdjustedRandScore=_Absent)
_mod("torchmetrics.image", StructuralSimilarityIndexMeasure=_Absent)
sys.modules["torchmetrics.regression"].PearsonCorrCoef = _Absent

# names some reference test modules import at module level but that are outside the scope (wrappers)
if not hasattr(metrics_b200, "ClasswiseWrapper"):
    metrics_b200.ClasswiseWrapper = _Absent

# optional: torch-CPU stand-ins for the kernel wrappers, so that the reference's CPU-tensor unit tests reach our host layer
if os.environ.get("MB200_REF_CPU_KERNELS") == "1":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_kernels  # noqa: E402

    cpu_kernels.install(sys.modules["metrics_b200._native"])
