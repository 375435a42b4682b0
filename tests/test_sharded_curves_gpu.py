"""GPU, world_size 2 over NCCL: the class-sharded AUROC/AP exchange (metrics_b200/parallel_curves.py) is bit-identical to
the gather-everything sync (reference: metric.py:511-563 `_sync_dist` + `gather_all_tensors`) and to one GPU evaluating the
concatenated data.  Skipped on a single-GPU box (the driver's scaling tier runs it)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_curves_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "_mgpu_sharded_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
