#!/bin/bash
# Round 2, GPU call N (1 GPU): racecheck on the new shared-memory kernels
set -x
O=gpurun_out
mkdir -p $O
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_fusion_gpu.py tests/test_binary_single_pass_gpu.py "tests/test_binned_gpu.py::test_binary_fast_path_equals_generic_kernel_and_oracle" tests/test_confmat_gpu.py -q -x -k "(float32 or bfloat16 or fast_path or deferred or cfg1) and not 1024 and not dense and not 0.999" > $O/r02_racecheck.log 2>&1; tail -8 $O/r02_racecheck.log
