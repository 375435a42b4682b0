#!/bin/bash
# segm mAP (K12 + mask mode of the matcher): parity; then the existing mAP / sharded tests and the K3 timing with 4 CTAs/SM
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_map_segm_gpu.py -q -x -m gpu 2>&1 | tail -25
timeout 900 python -m pytest tests/test_sharded_curves_gpu.py -q -x -m gpu 2>&1 | tail -5
