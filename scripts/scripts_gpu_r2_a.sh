#!/bin/bash
# Round 2, GPU call A (1 GPU): prototype + xfail details + diagnostics + cfg1 + bench
set -x
mkdir -p gpurun_out
O=gpurun_out
make -C metrics_b200/csrc tools > $O/r2_tools_build.log 2>&1
timeout 120 metrics_b200/csrc/build/k2_single_pass_proto > $O/r2_k2_proto.txt 2>&1; cat $O/r2_k2_proto.txt
timeout 400 python -m pytest tests/test_zzz_fuzz2_gpu.py -q -rx > $O/r2_fuzz2.log 2>&1; tail -12 $O/r2_fuzz2.log
timeout 400 python benchmarks/diag_r2.py --out $O/r2_diag.json > $O/r2_diag.log 2>&1; tail -60 $O/r2_diag.log
timeout 300 python benchmarks/run_configs.py --only cfg1 --out $O/r2_cfg1.json > $O/r2_cfg1.log 2>&1; tail -12 $O/r2_cfg1.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_bench20.json 2> $O/r2_bench20.err; cat $O/r2_bench20.json
