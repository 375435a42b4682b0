#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for P in bulk16 bulk32; do
MB200_ROWS_PATH=$P timeout 900 python -m pytest tests/test_confmat_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu_$P.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$P.log
tail -3 gpurun_out/pytest_gpu_$P.log
done
timeout 600 python bench.py --steps 2000 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for P in bulk16 bulk32; do
MB200_ROWS_PATH=$P timeout 600 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/bench_$P.json 2> gpurun_out/bench_$P.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$P.json')); print('$P', d['ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_$P.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_vec_kernel -s 30 -c 2 -o gpurun_out/prof_confmat_vec python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_vec.log 2>&1
MB200_ROWS_PATH=bulk32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_bulk_kernel -s 30 -c 2 -o gpurun_out/prof_confmat_bulk32 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bulk.log 2>&1
