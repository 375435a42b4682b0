"""cProfile of 1000 cfg3 updates (BinaryAUROC + BinaryAveragePrecision in one compute group), host-side breakdown."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metrics_b200 import MetricCollection  # noqa: E402
from metrics_b200.classification import BinaryAUROC, BinaryAveragePrecision, MulticlassAccuracy  # noqa: E402
from tests.helpers import cfg3_inputs  # noqa: E402

dev = torch.device("cuda", 0)
preds, target = cfg3_inputs()
dp, dt = preds.to(dev), target.to(dev)
mc = MetricCollection([BinaryAUROC(validate_args=False), BinaryAveragePrecision(validate_args=False)]).to(dev)


def updates():
    mc.reset()
    for i in range(1000):
        mc.update(dp[i], dt[i])


g = torch.Generator().manual_seed(0)
p1 = torch.randn(100, 1024, 5, generator=g).to(dev)
t1 = torch.randint(0, 5, (100, 1024), generator=g).to(dev)
acc = MulticlassAccuracy(num_classes=5, validate_args=False).to(dev)


def updates_cfg1():
    acc.reset()
    for i in range(100):
        acc.update(p1[i], t1[i])


for name, fn, n in (("cfg3", updates, 1000), ("cfg1", updates_cfg1, 100)):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print(f"{name} wall per update: {(time.perf_counter() - t0) / n * 1e6:.2f} us")
    pr = cProfile.Profile()
    pr.enable()
    fn()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
    print(s.getvalue()[:6000])
