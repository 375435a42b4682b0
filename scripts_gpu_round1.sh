#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python - > gpurun_out/cfg3_timing.log 2>&1 <<'PY'
import time, torch, sys
sys.path.insert(0, '.')
from tests.helpers import cfg3_inputs
from metrics_b200 import MetricCollection, _native
from metrics_b200.classification import BinaryAUROC, BinaryAveragePrecision
preds, target = cfg3_inputs()
preds, target = preds.cuda(), target.cuda()
mc = MetricCollection([BinaryAUROC(validate_args=False), BinaryAveragePrecision(validate_args=False)]).cuda()
for rep in range(3):
    mc.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(1000): mc.update(preds[i], target[i])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    res = mc.compute()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep{rep}: update phase {1e3*(t1-t0):.2f} ms ({1e3*(t1-t0):.1f} us/update), compute {1e3*(t2-t1):.3f} ms", {k: float(v) for k, v in res.items()})
# kernel-only: one curve_evaluate on 1e7 samples
p, t = preds.reshape(-1), target.reshape(-1)
for _ in range(3): _native.curve_evaluate(p, t)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): _native.curve_evaluate(p, t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"curve_evaluate(1e7): {ms:.3f} ms -> {150e6/ms/1e6:.1f} GB/s on the 150 MB algorithmic figure, {450e6/ms/1e6:.1f} GB/s on the 4-pass model")
PY
cat gpurun_out/cfg3_timing.log
