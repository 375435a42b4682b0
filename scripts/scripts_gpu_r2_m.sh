#!/bin/bash
# Round 2, GPU call M (1 GPU): full suite with arena / multilabel single pass, memcheck on the new kernels, bench + cfg3 line
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2m_all.log 2>&1; tail -4 $O/r2m_all.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fusion_gpu.py tests/test_binary_single_pass_gpu.py tests/test_curves64_gpu.py tests/test_normalize_aten_gpu.py tests/test_binned_gpu.py -q -x -k "not 1048576 and not 4194307 and not dense" > $O/r02_memcheck.log 2>&1; tail -6 $O/r02_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_fusion_gpu.py tests/test_binary_single_pass_gpu.py tests/test_binned_gpu.py -q -x -k "float32 and not dense and not 1024" > $O/r02_racecheck.log 2>&1; tail -6 $O/r02_racecheck.log
timeout 600 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines.json > $O/r2m_rooflines.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_kernel_rooflines.json'))
for k,v in d['kernels'].items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us  {v['achieved_gbs']:8.0f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2m_bench1.json 2> $O/r2m_bench1.err; tail -3 $O/r2m_bench1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'])
print(json.dumps(d['config']['cfg3'])[:900])
PY
