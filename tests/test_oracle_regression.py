"""CPU: numpy/fp64 regression oracle vs the reference goldens (the reference accumulates in fp32: 2e-6 relative)."""
import numpy as np
import pytest

from oracle import regression as orr

TOL = dict(rtol=3e-6, atol=1e-7)


def test_regression_functionals(golden_reg):
    g = golden_reg
    p1, t1, p2, t2 = g["reg/p1"], g["reg/t1"], g["reg/p2"], g["reg/t2"]
    np.testing.assert_allclose(orr.mean_squared_error(p1, t1), g["reg/mse"], **TOL)
    np.testing.assert_allclose(orr.mean_squared_error(p1, t1, squared=False), g["reg/rmse"], **TOL)
    np.testing.assert_allclose(orr.mean_squared_error(p2, t2, num_outputs=5), g["reg/mse_multi"], **TOL)
    np.testing.assert_allclose(orr.mean_absolute_error(p1, t1), g["reg/mae"], **TOL)
    np.testing.assert_allclose(orr.mean_absolute_percentage_error(p1, t1), g["reg/mape"], **TOL)
    np.testing.assert_allclose(orr.symmetric_mean_absolute_percentage_error(p1, t1), g["reg/smape"], **TOL)
    np.testing.assert_allclose(orr.weighted_mean_absolute_percentage_error(p1, t1), g["reg/wmape"], **TOL)
    np.testing.assert_allclose(orr.mean_squared_log_error(p1, t1), g["reg/msle"], **TOL)
    np.testing.assert_allclose(orr.log_cosh_error(p1, t1), g["reg/logcosh"], **TOL)
    np.testing.assert_allclose(orr.log_cosh_error(p2, t2), g["reg/logcosh_multi"], **TOL)
    np.testing.assert_allclose(orr.minkowski_distance(p1, t1, 3), g["reg/minkowski3"], **TOL)
    np.testing.assert_allclose(orr.minkowski_distance(p2, t2, 1.5), g["reg/minkowski1.5"], **TOL)
    for mo in ("raw_values", "uniform_average", "variance_weighted"):
        np.testing.assert_allclose(orr.r2_score(p2, t2, mo), g[f"reg/r2_{mo}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(orr.explained_variance(p2, t2, mo), g[f"reg/ev_{mo}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orr.relative_squared_error(p2, t2), g["reg/rse"], rtol=1e-5)
    np.testing.assert_allclose(orr.relative_squared_error(p2, t2, squared=False), g["reg/rrse"], rtol=1e-5)


def test_tweedie_deviance(golden_tweedie):
    g = golden_tweedie
    for c in range(int(g["n_cases"])):
        p, t, power = g[f"case{c}/preds"], g[f"case{c}/targets"], float(g[f"case{c}/power"])
        tol = dict(rtol=3e-6) if p.dtype == np.float32 else dict(rtol=1e-12)
        np.testing.assert_allclose(orr.tweedie_deviance_score(p, t, power), g[f"case{c}/value"], err_msg=f"case {c}", **tol)
