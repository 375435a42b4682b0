"""Fresh-process first-call latencies (lazy module loading, attribute setup) of a few entry points."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
import torch  # noqa: E402

torch.zeros(1, device="cuda").sum().item()
print("torch + context: %.2f s" % (time.perf_counter() - t0))
from metrics_b200 import _native  # noqa: E402
from metrics_b200.detection import MeanAveragePrecision  # noqa: E402
from tests.helpers import synth_detection  # noqa: E402

dev = torch.device("cuda", 0)


def timed(name, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print("%-44s %.1f ms" % (name, (time.perf_counter() - t) * 1e3))


p, t = torch.rand(100000, device=dev), torch.randint(0, 2, (100000,), device=dev)
timed("lib() load", lambda: _native.lib())
timed("curve_evaluate #1", lambda: _native.curve_evaluate(p, t, 1, unit_range=True))
timed("curve_evaluate #2", lambda: _native.curve_evaluate(p, t, 1, unit_range=True))
m = torch.rand(10, 64, 64, device=dev) > 0.5
timed("mask_pack_bits #1", lambda: _native.mask_pack_bits(m))
timed("mask_pack_bits #2", lambda: _native.mask_pack_bits(m))
preds, target = synth_detection(seed=0, n_img=50, n_gt=5, n_det=20, n_cls=5)
to = lambda items: [{k: v.to(dev) for k, v in d.items()} for d in items]  # noqa: E731
mp = MeanAveragePrecision().to(dev)
timed("mAP update #1", lambda: mp.update(to(preds), to(target)))
timed("mAP compute #1 (bbox)", lambda: mp.compute())
mp._computed = None
timed("mAP compute #2 (bbox)", lambda: mp.compute())
ms = MeanAveragePrecision(iou_type="segm").to(dev)
sp = [dict(masks=torch.rand(4, 64, 64, device=dev) > 0.5, scores=torch.rand(4, device=dev), labels=torch.zeros(4, dtype=torch.long, device=dev))]
st = [dict(masks=torch.rand(3, 64, 64, device=dev) > 0.5, labels=torch.zeros(3, dtype=torch.long, device=dev))]
timed("segm update #1", lambda: ms.update(sp, st))
import cProfile  # noqa: E402
import pstats  # noqa: E402

pr = cProfile.Profile()
pr.enable()
timed("segm compute #1", lambda: ms.compute())
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(6)
ms._computed = None
timed("segm compute #2", lambda: ms.compute())
