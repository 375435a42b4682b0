#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines.json > $O/r2q_rooflines.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_kernel_rooflines.json'))
for k,v in d['kernels'].items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us  {v['achieved_gbs']:8.0f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
timeout 300 python -m pytest tests/test_normalize_aten_gpu.py tests/test_regression_gpu.py -q -x 2>&1 | tail -3
