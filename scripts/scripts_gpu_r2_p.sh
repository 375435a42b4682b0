#!/bin/bash
# Round 2, GPU call P (1 GPU): softmax speculative pass, validation scratch, cfg4/cfg5 lines, rooflines
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_normalize_aten_gpu.py tests/test_confmat_gpu.py tests/test_curves_gpu.py tests/test_fusion_gpu.py tests/test_multilabel_gpu.py tests/test_fuzz_gpu.py -q -x > $O/r2p_tests.log 2>&1; tail -5 $O/r2p_tests.log
timeout 600 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines.json > $O/r2p_rooflines.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_kernel_rooflines.json'))
for k,v in d['kernels'].items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us  {v['achieved_gbs']:8.0f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
timeout 300 python bench.py --config cfg4 --steps 1 --warmup 0 > $O/r2p_cfg4.json 2> $O/r2p_cfg4.err; cut -c1-700 $O/r2p_cfg4.json; tail -2 $O/r2p_cfg4.err
timeout 300 python bench.py --config cfg5 --steps 1 --warmup 0 > $O/r2p_cfg5.json 2> $O/r2p_cfg5.err; cut -c1-900 $O/r2p_cfg5.json; tail -2 $O/r2p_cfg5.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2p_all.log 2>&1; tail -4 $O/r2p_all.log
