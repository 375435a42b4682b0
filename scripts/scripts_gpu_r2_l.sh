#!/bin/bash
# Round 2, GPU call L (8 GPUs): bench N=8 after the spin-up move + the sync diagnostic at world 8
set -x
O=gpurun_out
mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2l_bench8.json 2> $O/r2l_bench8.err; tail -3 $O/r2l_bench8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}, d['roofline']['frac'], d['config']['ms_per_step_per_rank'])
print(json.dumps(d['config']['sync'])[:300]); print({k:v for k,v in d['config']['cfg5'].items() if k in ('update_ms_4_batches','compute_ms','compute_ms_gather_everything','parity')}); print(d['e2e'])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 benchmarks/diag_sync_r2.py > $O/r2l_diag_sync8.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_diag_sync_8gpu.json'))
for k in ('nccl_all_reduce_8MB_i64','nccl_all_to_all_65MB','nccl_all_gather_65MB_per_rank','confmat_compute','cfg5'): print(k, json.dumps(d.get(k))[:700])
print(json.dumps(d.get('symm'))[:900])
PY
