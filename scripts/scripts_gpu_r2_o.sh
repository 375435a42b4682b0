#!/bin/bash
# Round 2, GPU call O (2 GPUs): sharded mAP worker, map tests, full suite
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_map_gpu.py tests/test_map_ious.py tests/test_sharded_curves_gpu.py tests/test_confmat_gpu.py -q -x > $O/r2o_tests.log 2>&1; tail -15 $O/r2o_tests.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2o_all.log 2>&1; tail -4 $O/r2o_all.log
