#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_binary_gpu.py -m gpu -x -q > gpurun_out/pytest_bin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bin.log
tail -40 gpurun_out/pytest_bin.log
