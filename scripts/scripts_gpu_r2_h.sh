#!/bin/bash
# Round 2, GPU call H (1 GPU): K2 single pass, K4 table search, K6 speculative pass — parity, timing, ncu
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_binned_gpu.py tests/test_binary_single_pass_gpu.py tests/test_normalize_aten_gpu.py tests/test_binary_gpu.py tests/test_curves_gpu.py -q -x > $O/r2h_tests.log 2>&1; tail -6 $O/r2h_tests.log
timeout 600 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines.json > $O/r2h_rooflines.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_kernel_rooflines.json'))
for k,v in d['kernels'].items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us  {v['achieved_gbs']:8.0f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
for k in k4:binned_binary_fast k2:bin_count_flat_both k6:sigmoid_spec; do
  name=${k%%:*}; kern=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -c 2 -f -o $O/r02_prof_${name}_v2 python benchmarks/prof_one.py $name > $O/r2h_ncu_$name.log 2>&1; tail -2 $O/r2h_ncu_$name.log
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2h_all.log 2>&1; tail -4 $O/r2h_all.log
