"""cProfile of the cfg4 mAP update phase and compute() on the GPU box (host-side cost breakdown)."""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metrics_b200.detection import MeanAveragePrecision  # noqa: E402
from tests.helpers import synth_detection  # noqa: E402

dev = torch.device("cuda", 0)
preds, target = synth_detection(seed=0, n_img=5000, n_gt=20, n_det=100, n_cls=80, crowd_frac=0.02)
to = lambda items: [{k: v.to(dev) for k, v in d.items()} for d in items]  # noqa: E731
preds, target = to(preds), to(target)
m = MeanAveragePrecision().to(dev)
m.warn_on_many_detections = False


def updates():
    m.reset()
    for i in range(0, 5000, 100):
        m.update(preds[i:i + 100], target[i:i + 100])


updates()
m.compute()
torch.cuda.synchronize()
for name, fn in (("update", updates), ("compute", lambda: (setattr(m, "_computed", None), m.compute(), torch.cuda.synchronize()))):
    pr = cProfile.Profile()
    pr.enable()
    fn()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
    print("=====", name)
    print(s.getvalue()[:5000])
