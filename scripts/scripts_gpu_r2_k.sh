#!/bin/bash
# Round 2, GPU call K (2 GPUs): bench after the spin-up move at N=1 and N=2, K4 timing, reference arm
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2k_ref.json 2> $O/r2k_ref.err; cut -c1-400 $O/r2k_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2k_bench1.json 2> $O/r2k_bench1.err; tail -3 $O/r2k_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2k_bench2.json 2> $O/r2k_bench2.err; tail -3 $O/r2k_bench2.err
python - <<'PY'
import json
for n in (1,2):
    d=json.load(open(f'gpurun_out/r2k_bench{n}.json'))
    print(n, {k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'], d['config']['ms_per_step_per_rank'], d['config']['sync']['compute_ms_min'], d['e2e']['value'])
    print('   cfg5', {k:v for k,v in d['config']['cfg5'].items() if k in ('update_ms_4_batches','compute_ms','compute_ms_gather_everything','parity')})
PY
timeout 300 python benchmarks/prof_one.py k4
timeout 300 python bench.py --config cfg3 --steps 3 --warmup 1 > $O/r2k_cfg3.json 2> $O/r2k_cfg3.err; cut -c1-600 $O/r2k_cfg3.json
