#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_map_segm_gpu.py tests/test_zz_kld_gpu.py -q -x -m gpu 2>&1 | tail -15
timeout 900 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines_u.json > $O/r2u_rooflines.log 2>&1; tail -3 $O/r2u_rooflines.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_kernel_rooflines_u.json'))
for k,v in d['kernels'].items():
    if 'K12' in k or 'K13' in k: print(f"{k:100s} {v['ms']*1e3:9.1f} us  {v['achieved_gbs']:8.0f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
