#!/bin/bash
# Round 2, GPU call I (8 GPUs): the sharded-curves worker and bench.py at N=8 (peer-memory exchange at world 8)
set -x
O=gpurun_out
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 tests/_mgpu_sharded_worker.py > $O/r2i_sharded8.log 2>&1; tail -3 $O/r2i_sharded8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2i_bench8.json 2> $O/r2i_bench8.err; tail -5 $O/r2i_bench8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}, d['roofline']['frac'])
print(json.dumps(d['config']['sync'])[:500]); print(json.dumps(d['config']['cfg5'])[:900]); print(d['e2e'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > $O/r2i_bench8b.json 2> $O/r2i_bench8b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench8b.json'))
print('second run', {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])
PY
