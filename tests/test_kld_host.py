"""CPU: KLDivergence host logic (functional, class, reductions, argument checks) with the kernel's torch stand-in, against the
reference goldens; the fp64 oracle against the same goldens."""
import numpy as np
import pytest

from oracle import regression as oreg
from tests.kld_cases import argument_errors, replay


@pytest.fixture()
def standin(monkeypatch):
    from metrics_b200 import _native
    from tests.reference_runtime import cpu_kernels

    monkeypatch.setattr(_native, "kl_divergence_rows", cpu_kernels.kl_divergence_rows)


def test_replay_reference_goldens_on_the_standin(golden_kld, standin):
    assert replay(golden_kld, "cpu") == 24


def test_argument_errors(standin):
    argument_errors("cpu")


def test_oracle_matches_the_reference_goldens(golden_kld):
    for k in range(int(golden_kld["n_cases"])):
        key = f"case{k}"
        log_prob, dt = (int(x) for x in golden_kld[f"{key}/meta"])
        p, q = golden_kld[f"{key}/p"], golden_kld[f"{key}/q"]
        tol = dict(rtol=2e-6, atol=2e-7) if dt == 0 else dict(rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(oreg.kl_divergence_rows(p, q, bool(log_prob)), golden_kld[f"{key}/measures"], **tol)
        for red in ("mean", "sum", "none"):
            np.testing.assert_allclose(oreg.kl_divergence(p, q, bool(log_prob), red), golden_kld[f"{key}/{red}"], **tol)
