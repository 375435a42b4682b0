#!/bin/bash
# round-2 closing run on one GPU: the full GPU suite, the bench line (both arms), ncu captures of this session's kernels,
# compute-sanitizer over their tests
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/ -q -x -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2v_bench.json 2> $O/r2v_bench.err; tail -c 2500 $O/r2v_bench.json
for pair in "k3:scan_chained" "k3:pack_binary" "k3:radix_onesweep" "k12:mask_pack_bits" "k13:kl_rows"; do
  name=${pair%%:*}; kern=${pair##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -c 2 -f -o $O/r02_prof_${kern} python benchmarks/prof_one.py $name > $O/r2v_ncu_$kern.log 2>&1; tail -2 $O/r2v_ncu_$kern.log
done
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_curves_gpu.py tests/test_map_segm_gpu.py tests/test_zz_kld_gpu.py tests/test_multilabel_gpu.py -q -x -m gpu > $O/r2v_memcheck.log 2>&1; tail -4 $O/r2v_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_curves_gpu.py tests/test_map_segm_gpu.py -q -x -m gpu -k "chained or label_in_key or random_sizes or pack_and_pair or docstring" > $O/r2v_racecheck.log 2>&1; tail -4 $O/r2v_racecheck.log
