#!/bin/bash
# Round 2, GPU call D (2 GPUs): peer-memory exchange — correctness worker, diag, bench
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_sharded_curves_gpu.py -x -q > $O/r2d_sharded.log 2>&1; tail -5 $O/r2d_sharded.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 benchmarks/diag_sync_r2.py > $O/r2d_diag_sync.log 2>&1; tail -120 $O/r2d_diag_sync.log
