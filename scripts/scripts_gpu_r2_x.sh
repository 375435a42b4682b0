#!/bin/bash
# 2 GPUs: sharded paths (label-in-key evaluate_keys, segm), curve tests, the bench line at N=2
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_curves_gpu.py tests/test_sharded_curves_gpu.py tests/test_zz_kld_gpu.py -q -x -m gpu 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2x_bench_2gpu.json 2> $O/r2x_bench_2gpu.err; tail -c 600 $O/r2x_bench_2gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench_2gpu.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], d["config"].get("ms_per_step_per_rank"))
print("sync", json.dumps(d["config"]["sync"])[:300]); print("cfg5", json.dumps(d["config"]["cfg5"])[:700])
PY
