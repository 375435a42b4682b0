#!/bin/bash
# Round-1 evidence run on one B200: full GPU parity suite, smoke, bench (both arms), per-config timings, ncu launch lists
# and one full capture per dominant kernel.  Outputs land in gpurun_out/ (copied into profiles/ by hand afterwards).
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
timeout 900 python benchmarks/run_configs.py --out $O/configs_1gpu.json > $O/configs_1gpu.log 2>&1; tail -60 $O/configs_1gpu.log
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 > $O/bench_under_ncu.log 2>&1
# full capture of the dominant kernel (K1, vec rows) — one launch after warm-up
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_vec_kernel -s 5 -c 1 -o $O/prof_confmat -f python bench.py --steps 8 --warmup 3 > $O/ncu_full.log 2>&1
# curve pipeline: per-kernel list at 1e7 binary samples and at 16384 x 1000 multiclass; full capture of the one-sweep pass
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/curve_launches_1e7.csv python benchmarks/curve_kernel_times.py 10000000 1 > $O/curve_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/curve_launches_mc.csv python benchmarks/curve_kernel_times.py 16384 1000 >> $O/curve_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:radix_onesweep_kernel -c 1 -o $O/prof_onesweep -f python benchmarks/curve_kernel_times.py 10000000 1 > $O/ncu_onesweep.log 2>&1
timeout 300 python benchmarks/kernel_rooflines.py $O/kernel_rooflines.json > $O/kernel_rooflines.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > $O/smi.txt
ls -la $O | tail -30
