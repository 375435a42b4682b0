#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/ -q -x -m gpu 2>&1 | tail -4
python benchmarks/prof_one.py k13
timeout 900 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines_final.json > $O/r2w_rooflines.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_kernel_rooflines_final.json'))
for k,v in d['kernels'].items():
    print(f"{k:100s} {v['ms']*1e3:9.1f} us  {v['achieved_gbs']:8.0f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
