#!/bin/bash
# 8 GPUs: the sharded tests (curves, mAP incl. masks) and the bench line at N=8
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_sharded_curves_gpu.py -q -x -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2y_bench_8gpu.json 2> $O/r2y_bench_8gpu.err; tail -c 300 $O/r2y_bench_8gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench_8gpu.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], d["config"].get("ms_per_step_per_rank"))
print("sync", json.dumps(d["config"]["sync"])[:300]); print("cfg5", json.dumps(d["config"]["cfg5"])[:900])
PY
