#!/bin/bash
# Round 2, GPU call E (1 GPU): new parity tests first, then the whole GPU suite, then bench at the driver's settings
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_curves64_gpu.py tests/test_normalize_aten_gpu.py tests/test_fusion_gpu.py tests/test_zzz_fuzz2_gpu.py -q -rf > $O/r2e_new.log 2>&1; tail -30 $O/r2e_new.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2e_all.log 2>&1; tail -6 $O/r2e_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2e_bench1.json 2> $O/r2e_bench1.err; tail -3 $O/r2e_bench1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'], d.get('aten_gpu_baseline'), d['cpu_baseline'])
PY
