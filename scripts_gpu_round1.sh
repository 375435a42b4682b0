#!/bin/bash
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 10 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 benchmarks/run_configs.py --only cfg5 --out gpurun_out/configs_cfg5_2gpu.json > gpurun_out/cfg5_2gpu.log 2>&1; tail -30 gpurun_out/cfg5_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/bench_ref_2.json 2>&1; cat gpurun_out/bench_ref_2.json | head -c 400
