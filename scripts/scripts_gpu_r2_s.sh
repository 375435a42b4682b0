#!/bin/bash
# K3 step 2: histograms fused into the pack kernel + chained single-pass scan — parity, then timing and the launch list
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_curves_gpu.py tests/test_curves64_gpu.py tests/test_multilabel_gpu.py tests/test_fuzz_gpu.py tests/test_zzz_fuzz2_gpu.py tests/test_consumers_gpu.py tests/test_fusion_gpu.py tests/test_sharded_curves_gpu.py tests/test_torch_ops_gpu.py -q -x -m gpu 2>&1 | tail -8
timeout 600 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines_s.json > $O/r2s_rooflines.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_kernel_rooflines_s.json'))
for k,v in d['kernels'].items():
    if 'K3' in k: print(f"{k:100s} {v['ms']*1e3:9.1f} us  {v['achieved_gbs']:8.0f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file $O/r2s_curve_launches.csv python benchmarks/curve_kernel_times.py > $O/r2s_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file $O/r2s_curve_launches_mc.csv python benchmarks/curve_kernel_times.py 16384 1000 >> $O/r2s_ncu.log 2>&1
