"""CPU: Tweedie deviance.  (a) the host layer against goldens from the reference, kernel replaced by its stand-in;
(b) the kernel's own term function (csrc/regression_terms.cuh), compiled for the host by nvcc, element by element against
the reference — so the formulas the GPU runs are checked here even without a GPU."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT
from tests.tweedie_cases import domain_errors, replay


def test_replay_reference_goldens(golden_tweedie, cpu_kernel_standins):
    assert replay(golden_tweedie, "cpu") == 24


def test_domain_errors_and_corners(cpu_kernel_standins):
    domain_errors("cpu")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("reg_terms") / "reg_terms_host")
    src = os.path.join(ROOT, "metrics_b200", "csrc", "tools", "reg_terms_host.cu")
    build = subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-2000:]

    def run(op, param, eps, precision, preds, targets):
        lines = "".join(f"{p!r} {t!r}\n" for p, t in zip(preds.tolist(), targets.tolist()))
        out = subprocess.run([exe, str(op), repr(float(param)), repr(float(eps)), precision], input=lines, capture_output=True,
                             text=True, check=True).stdout
        return np.array([[float(v) for v in row.split()] for row in out.strip().splitlines()])

    return run


@pytest.mark.parametrize("precision,rtol,atol", [("f64", 1e-13, 1e-14), ("f32", 3e-6, 2e-6)])  # atol: p == t gives 0 +- rounding
def test_kernel_term_function_matches_the_reference_per_element(golden_tweedie, harness, precision, rtol, atol):
    g = golden_tweedie
    preds = g["elem/preds"]
    for power in g["powers"].tolist():
        if power == 0:
            continue  # served by the squared-error op
        targets = g[f"elem/p{power}/targets"]
        terms = harness(10, power, 0.0, precision, preds, targets)
        np.testing.assert_allclose(terms[:, 0], g[f"elem/p{power}/deviance"], rtol=rtol, atol=atol, err_msg=f"power {power}")
        np.testing.assert_array_equal(terms[:, 1], (preds <= 0).astype(float))
        np.testing.assert_array_equal(terms[:, 2], (targets < 0).astype(float))
        np.testing.assert_array_equal(terms[:, 3], (targets == 0).astype(float))
    census = harness(10, 2.0, 0.0, precision, np.array([0.0, -1.0, 2.0]), np.array([-3.0, 0.0, 1.0]))
    assert census[:, 1:].tolist() == [[1, 1, 0], [1, 0, 1], [0, 0, 0]]


def test_the_other_ops_are_untouched_by_the_shared_header(harness):
    p, t = np.array([1.5, -2.0, 0.25]), np.array([1.0, 3.0, 0.25])
    d = p - t
    np.testing.assert_allclose(harness(0, 0, 0, "f64", p, t)[:, 0], d * d)
    np.testing.assert_allclose(harness(1, 0, 0, "f64", p, t)[:, 0], np.abs(d))
    np.testing.assert_allclose(harness(2, 0, 1.17e-6, "f64", p, t)[:, 0], np.abs(d) / np.maximum(np.abs(t), 1.17e-6))
    np.testing.assert_allclose(harness(7, 3.0, 0, "f64", p, t)[:, 0], np.abs(d) ** 3)
    np.testing.assert_allclose(harness(4, 0, 0, "f64", p, t), np.stack([np.abs(d), np.abs(t)], 1))
    np.testing.assert_allclose(harness(8, 0, 0, "f64", p, t), np.stack([t * t, t, (t - p) ** 2], 1))
    np.testing.assert_allclose(harness(9, 0, 0, "f64", p, t), np.stack([t - p, (t - p) ** 2, t, t * t], 1))
