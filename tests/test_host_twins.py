"""CPU: the GPU parity suites replayed on CPU tensors with the kernel wrappers swapped for their torch stand-ins
(tests/host_twin_plugin.py).  Everything above the C-ABI — validation, the format seams, states, reducers, compute groups,
class / functional / task-wrapper plumbing — is thereby held to the reference's goldens on every CPU run; the kernels are
held to the same goldens by the same files on the GPU box.  One child pytest process, so that the stand-ins never exist
in this process."""
import os
import re
import subprocess
import sys

from tests.conftest import ROOT

SUITES = ["test_confmat_gpu", "test_topk_samplewise_gpu", "test_binary_gpu", "test_consumers_gpu", "test_curves_gpu",
          "test_multilabel_gpu", "test_binned_gpu", "test_atfixed_gpu", "test_regression_gpu", "test_logauc", "test_curves64_gpu",
          "test_fusion_gpu"]


def _replay(files, *extra) -> int:
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "tests.host_twin_plugin", "-p", "no:cacheprovider",
           *extra, *[os.path.join("tests", f"{name}.py") for name in files]]
    run = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = run.stdout.strip().splitlines()[-1] if run.stdout.strip() else run.stderr[-400:]
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-1000:]
    assert "failed" not in tail and "skipped" not in tail, tail
    return int(re.search(r"(\d+) passed", tail).group(1))


def test_gpu_parity_suites_pass_on_kernel_standins():
    assert _replay(SUITES) >= 290


def test_map_marshalling_around_the_kernel_call():
    """`MeanAveragePrecision.compute` with the numpy oracle standing in for the COCO kernels.  The expectations of
    test_map_gpu.py come from the same oracle, so this pins only the marshalling around the call — state concatenation,
    per-image counts, iscrowd / area defaults, box formats, micro / class_metrics re-evaluation, the summary table, the
    extended summary — not the evaluation.  (The oracle-sized and full-size cases are left to the GPU run.)"""
    assert _replay(["test_map_gpu"], "-k", "not cfg4 and not synthetic") >= 5


def test_segm_map_host_layer_on_kernel_standins():
    """`MeanAveragePrecision` with instance masks replayed on CPU: mask state entries, the table marshalling of `_mask_tables`,
    the match + accumulate call sequence, both IoU types with their prefixes, micro / class metrics, the extended-summary IoUs
    and the run-length json round trip — with numpy stand-ins for K12 and the two mAP phases (built on the oracle's matching
    loop and precision sampling, tests/reference_runtime/cpu_kernels.py)."""
    assert _replay(["test_map_segm_gpu"]) >= 12
