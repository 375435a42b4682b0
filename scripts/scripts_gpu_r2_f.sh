#!/bin/bash
# Round 2, GPU call F (1 GPU): full GPU suite, kernel rooflines, ncu captures (k4, fused, k1b), bench launch list
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2f_all.log 2>&1; tail -6 $O/r2f_all.log
timeout 600 python benchmarks/kernel_rooflines.py $O/r02_kernel_rooflines.json > $O/r2f_rooflines.log 2>&1; tail -3 $O/r2f_rooflines.log
for k in k4:binned_bucket fused:stats_softmax k1b:rows_vec; do
  name=${k%%:*}; kern=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -c 2 -f -o $O/r02_prof_$name python benchmarks/prof_one.py $name > $O/r2f_ncu_$name.log 2>&1; tail -2 $O/r2f_ncu_$name.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/r2f_bench_under_ncu.log 2>&1
