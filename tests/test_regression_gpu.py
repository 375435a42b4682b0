"""GPU: K9 regression running sums vs reference goldens (reference accumulates in fp32: tolerance 3e-6 relative;
R2 / explained variance / RSE subtract nearly equal fp32 sums in the reference: 1e-5)."""
import numpy as np
import pytest
import torch

from oracle import regression as orr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = dict(rtol=3e-6, atol=1e-7)
LOOSE = dict(rtol=1e-5, atol=1e-6)


def test_functionals_vs_golden(golden_reg):
    import metrics_b200.functional.regression as F

    g = golden_reg
    p1, t1, p2, t2 = (torch.from_numpy(g[f"reg/{k}"]).to(DEV) for k in ("p1", "t1", "p2", "t2"))
    c = lambda x: x.cpu().numpy()  # noqa: E731
    np.testing.assert_allclose(c(F.mean_squared_error(p1, t1)), g["reg/mse"], **TOL)
    np.testing.assert_allclose(c(F.mean_squared_error(p1, t1, squared=False)), g["reg/rmse"], **TOL)
    np.testing.assert_allclose(c(F.mean_squared_error(p2, t2, num_outputs=5)), g["reg/mse_multi"], **TOL)
    np.testing.assert_allclose(c(F.mean_absolute_error(p1, t1)), g["reg/mae"], **TOL)
    np.testing.assert_allclose(c(F.mean_absolute_percentage_error(p1, t1)), g["reg/mape"], **TOL)
    np.testing.assert_allclose(c(F.symmetric_mean_absolute_percentage_error(p1, t1)), g["reg/smape"], **TOL)
    np.testing.assert_allclose(c(F.weighted_mean_absolute_percentage_error(p1, t1)), g["reg/wmape"], **TOL)
    np.testing.assert_allclose(c(F.mean_squared_log_error(p1, t1)), g["reg/msle"], **TOL)
    np.testing.assert_allclose(c(F.log_cosh_error(p1, t1)), g["reg/logcosh"], **TOL)
    np.testing.assert_allclose(c(F.log_cosh_error(p2, t2)), g["reg/logcosh_multi"], **TOL)
    np.testing.assert_allclose(c(F.minkowski_distance(p1, t1, 3)), g["reg/minkowski3"], **TOL)
    np.testing.assert_allclose(c(F.minkowski_distance(p2, t2, 1.5)), g["reg/minkowski1.5"], **TOL)
    for mo in ("raw_values", "uniform_average", "variance_weighted"):
        np.testing.assert_allclose(c(F.r2_score(p2, t2, multioutput=mo)), g[f"reg/r2_{mo}"], **LOOSE)
        np.testing.assert_allclose(c(F.explained_variance(p2, t2, multioutput=mo)), g[f"reg/ev_{mo}"], **LOOSE)
    np.testing.assert_allclose(c(F.r2_score(p1, t1)), g["reg/r2_1d"], **LOOSE)
    np.testing.assert_allclose(c(F.r2_score(p2, t2, adjusted=3)), g["reg/r2_adj"], **LOOSE)
    np.testing.assert_allclose(c(F.relative_squared_error(p2, t2)), g["reg/rse"], **LOOSE)
    np.testing.assert_allclose(c(F.relative_squared_error(p2, t2, squared=False)), g["reg/rrse"], **LOOSE)
    np.testing.assert_allclose(c(F.explained_variance(p1, t1)), g["reg/ev_1d"], **LOOSE)


def test_classes_over_batches_and_collection(golden_reg):
    from metrics_b200 import MetricCollection
    from metrics_b200.regression import ExplainedVariance, MeanAbsoluteError, MeanSquaredError, R2Score

    g = golden_reg
    p1, t1, p2, t2 = (torch.from_numpy(g[f"reg/{k}"]).to(DEV) for k in ("p1", "t1", "p2", "t2"))
    mc = MetricCollection([MeanSquaredError(), MeanAbsoluteError(), R2Score()]).to(DEV)
    for a, b in zip(p1.chunk(4), t1.chunk(4)):
        mc.update(a, b)
    res = mc.compute()
    np.testing.assert_allclose(res["MeanSquaredError"].cpu().numpy(), g["reg/class/MeanSquaredError"], **TOL)
    np.testing.assert_allclose(res["MeanAbsoluteError"].cpu().numpy(), g["reg/class/MeanAbsoluteError"], **TOL)
    np.testing.assert_allclose(res["R2Score"].cpu().numpy(), g["reg/class/R2Score"], **LOOSE)
    ev, mm = ExplainedVariance(multioutput="raw_values").to(DEV), MeanSquaredError(num_outputs=5).to(DEV)
    for a, b in zip(p2.chunk(4), t2.chunk(4)):
        ev.update(a, b)
        mm.update(a, b)
    np.testing.assert_allclose(ev.compute().cpu().numpy(), g["reg/class/ExplainedVariance"], **LOOSE)
    np.testing.assert_allclose(mm.compute().cpu().numpy(), g["reg/class/MeanSquaredErrorMulti"], **TOL)


def test_large_and_wide_inputs_vs_oracle_and_determinism():
    import metrics_b200.functional.regression as F

    g = torch.Generator().manual_seed(8)
    p, t = torch.randn(3_000_000, generator=g), torch.randn(3_000_000, generator=g)
    a = F.mean_squared_error(p.to(DEV), t.to(DEV))
    b = F.mean_squared_error(p.to(DEV), t.to(DEV))
    assert torch.equal(a, b)  # fixed reduction order
    np.testing.assert_allclose(a.cpu().numpy(), orr.mean_squared_error(p.numpy(), t.numpy()), rtol=1e-6)
    p, t = torch.randn(700, 1000, generator=g), torch.randn(700, 1000, generator=g)
    np.testing.assert_allclose(F.mean_squared_error(p.to(DEV), t.to(DEV), num_outputs=1000).cpu().numpy(),
                               orr.mean_squared_error(p.numpy(), t.numpy(), num_outputs=1000), rtol=1e-6)
    pd, td = p.double(), t.double()
    np.testing.assert_allclose(F.mean_absolute_error(pd.to(DEV), td.to(DEV)).cpu().numpy(), orr.mean_absolute_error(pd.numpy(), td.numpy()), rtol=1e-12)
