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
_mod("torchmetrics.clustering", AdjustedRandScore=_Absent, CalinskiHarabaszScore=_Absent)
_mod("torchmetrics.image", StructuralSimilarityIndexMeasure=_Absent)
sys.modules["torchmetrics.regression"].PearsonCorrCoef = _Absent


def _absent_fn(*a, **k):
    raise NotImplementedError("not part of metrics_b200's scope")


# tests/unittests/regression/test_mean_error.py imports these at module level (NRMSE is outside the scope; `permetrics` is a
# test-only dependency missing from this image and only used as NRMSE's oracle)
_mod("torchmetrics.regression.nrmse", NormalizedRootMeanSquaredError=_Absent)
sys.modules["torchmetrics.functional"].normalized_root_mean_squared_error = _absent_fn
_mod("permetrics")
_mod("permetrics.regression", RegressionMetric=_absent_fn)

# optional: torch-CPU stand-ins for the kernel wrappers, so that the reference's CPU-tensor unit tests reach our host layer
if os.environ.get("MB200_REF_CPU_KERNELS") == "1":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_kernels  # noqa: E402

    cpu_kernels.install(sys.modules["metrics_b200._native"])
