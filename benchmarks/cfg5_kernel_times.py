"""cfg5 on one GPU: 4 updates + compute of MetricCollection([MulticlassF1Score, MulticlassAUROC], C=1000) — run under
`ncu --metrics gpu__time_duration.sum --profile-from-start off` for a per-kernel list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metrics_b200 import MetricCollection  # noqa: E402
from metrics_b200.classification import MulticlassAUROC, MulticlassF1Score  # noqa: E402
from tests.helpers import cfg5_rank_batches  # noqa: E402

dev = torch.device("cuda", 0)
batches = [(lg.to(dev), tg.to(dev)) for lg, tg in cfg5_rank_batches(0, 4)]
mc = MetricCollection([MulticlassF1Score(num_classes=1000, validate_args=False),
                       MulticlassAUROC(num_classes=1000, validate_args=False)]).to(dev)


def run():
    mc.reset()
    for lg, tg in batches:
        mc.update(lg, tg)
    return mc.compute()


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
