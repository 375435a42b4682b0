#!/bin/bash
# Round 2, GPU call G (1 GPU): K4 fast path parity + timing, ncu of it, rooflines again, torch ops tests, bench
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_binned_gpu.py tests/test_torch_ops_gpu.py tests/test_map_gpu.py tests/test_curves64_gpu.py -q -x > $O/r2g_tests.log 2>&1; tail -6 $O/r2g_tests.log
timeout 300 python benchmarks/prof_one.py k4 > $O/r2g_k4.log 2>&1; cat $O/r2g_k4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:binned_binary_fast -c 2 -f -o $O/r02_prof_k4fast python benchmarks/prof_one.py k4 > $O/r2g_ncu_k4.log 2>&1; tail -2 $O/r2g_ncu_k4.log
timeout 600 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines.json > $O/r2g_rooflines.log 2>&1; tail -3 $O/r2g_rooflines.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2g_bench1.json 2> $O/r2g_bench1.err; tail -3 $O/r2g_bench1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'])
print(json.dumps(d['config']['cfg5'])[:600])
PY
