#!/bin/bash
# Round 2, GPU call J (8 GPUs): scaling diagnostic + bench N=8 with the fixed oracle check
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 benchmarks/diag_scale_r2.py > $O/r2j_diag_scale.log 2>&1; tail -40 $O/r2j_diag_scale.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2j_bench8.json 2> $O/r2j_bench8.err; tail -5 $O/r2j_bench8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}, d['roofline']['frac'], d['config']['ms_per_step_per_rank'])
print(json.dumps(d['config']['sync'])[:500]); print(json.dumps(d['config']['cfg5'])[:1200]); print(d['e2e'])
PY
