import time, torch, sys, os
sys.path.insert(0, os.getcwd())
from metrics_b200.detection import MeanAveragePrecision
from tests.helpers import synth_detection
dev = torch.device("cuda", 0)
preds, target = synth_detection(seed=0, n_img=5000, n_gt=20, n_det=100, n_cls=80, crowd_frac=0.02)
to = lambda items: [{k: v.to(dev) for k, v in d.items()} for d in items]
preds, target = to(preds), to(target)
m = MeanAveragePrecision().to(dev); m.warn_on_many_detections = False
import cProfile, pstats
for rep in range(4):
    m.reset(); torch.cuda.synchronize()
    pr = cProfile.Profile() if rep in (0, 3) else None
    t0 = time.perf_counter()
    if pr: pr.enable()
    for i in range(0, 5000, 100):
        m.update(preds[i:i + 100], target[i:i + 100])
    if pr: pr.disable()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rep", rep, "host %.1f ms, +sync %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    if pr: pstats.Stats(pr).sort_stats("tottime").print_stats(8)
