#!/bin/bash
# Round 2, GPU call C (2 GPUs): bench v2 at N=1 (both arms) and N=2
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2c_ref.json 2> $O/r2c_ref.err; cat $O/r2c_ref.json | cut -c1-900
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2c_bench1.json 2> $O/r2c_bench1.err; tail -5 $O/r2c_bench1.err; cat $O/r2c_bench1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2c_bench2.json 2> $O/r2c_bench2.err; tail -5 $O/r2c_bench2.err; cat $O/r2c_bench2.json
