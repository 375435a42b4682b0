#!/bin/bash
# First GPU call of round 2 (one B200, ~6 min): everything written after round 1's GPU budget ran out, newest risk first,
# each under its own timeout so that one hang cannot eat the call.  Outputs land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash scripts_gpu_round2_first.sh'
set -x
mkdir -p gpurun_out
O=gpurun_out
# 0. native binaries (seconds): the C-ABI without Python, and the round-2 K2 prototype with its built-in check + timing
make -C metrics_b200/csrc tools > $O/r2_tools_build.log 2>&1
timeout 60 metrics_b200/csrc/build/abi_smoke > $O/r2_abi_smoke.txt 2>&1; tail -3 $O/r2_abi_smoke.txt
timeout 120 metrics_b200/csrc/build/k2_single_pass_proto > $O/r2_k2_proto.txt 2>&1; cat $O/r2_k2_proto.txt
# 1. the launch path with declared ctypes signatures: smoke + the oldest parity suite
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
timeout 300 python -m pytest tests/test_confmat_gpu.py tests/test_regression_gpu.py tests/test_map_gpu.py -q -x > $O/r2_core.log 2>&1; tail -3 $O/r2_core.log
# 2. the new GPU files (Tweedie op, group fairness, second fuzz draw) WITHOUT -x: collect every failure in one go
timeout 400 python -m pytest tests/test_zz_tweedie_gpu.py tests/test_zz_group_fairness_gpu.py tests/test_zzz_fuzz2_gpu.py -q -rxf > $O/r2_new.log 2>&1; tail -25 $O/r2_new.log
# 3. everything else, as the driver runs it
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_all.log 2>&1; tail -4 $O/r2_all.log
# 4. cfg1 again: did the plain-int launch path move the 15.1 us / update?
timeout 300 python benchmarks/run_configs.py --only cfg1 --out $O/r2_cfg1.json > $O/r2_cfg1.log 2>&1; tail -5 $O/r2_cfg1.log
timeout 300 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; cat $O/r2_bench.json
