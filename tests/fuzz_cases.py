"""Shared by tests/test_fuzz_gpu.py (real kernels) and tests/test_fuzz_host.py (kernel stand-ins on the CPU): replay one case
of tests/golden/fuzz.npz — random (functional, kwargs, inputs) triples with the unmodified reference's outputs."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

_DT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16, "float64": torch.float64,
       "int64": torch.int64}


def n_cases(name: str = "fuzz") -> int:
    path = os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz")
    return int(np.load(path)["n_cases"])


def _flatten(res):
    if isinstance(res, (tuple, list)):
        out = []
        for part in res:
            out.extend(part if isinstance(part, (tuple, list)) else [part])
        return out
    return [res]


def run_case(g, k: int, DEV: str) -> None:
    import metrics_b200.functional.classification as F_cls
    import metrics_b200.functional.regression as F_reg

    spec = json.loads(str(g[f"{k}/spec"]))
    F = F_reg if spec.get("module") == "regression" else F_cls
    fn, kwargs, pdt = spec["fn"], spec["kwargs"], _DT[spec["preds_dtype"]]
    preds = torch.from_numpy(g[f"{k}/preds"]).to(DEV).to(pdt)
    target = torch.from_numpy(g[f"{k}/target"]).to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            got = _flatten(getattr(F, fn)(preds, target, **kwargs))
        except NotImplementedError as err:
            pytest.skip(f"documented gap: {err}")
    assert len(got) == int(g[f"{k}/n_out"]), (fn, kwargs)
    half = pdt in (torch.float16, torch.bfloat16)
    loose = any(s in fn for s in ("kappa", "matthews"))
    for i, t in enumerate(got):
        exp = g[f"{k}/out{i}"]
        t = t.float() if t.dtype in (torch.float16, torch.bfloat16) else t
        arr = t.cpu().numpy()
        assert arr.shape == exp.shape, f"shape mismatch {fn} {kwargs} out{i}: {arr.shape} vs {exp.shape}"
        if np.issubdtype(exp.dtype, np.integer) or exp.dtype == np.bool_:
            np.testing.assert_array_equal(arr, exp, err_msg=f"{fn} {kwargs} out{i}")
        else:
            # half-precision scores: curve *values* (thresholds, sigmoid outputs) carry the input precision
            rtol = 1e-5 if loose else (4e-3 if half and ("roc" in fn or "curve" in fn) else 1e-6)
            atol = 1e-6 if loose else (1e-3 if half and ("roc" in fn or "curve" in fn) else 1e-7)
            if spec.get("module") == "regression":
                # fp64-accumulated sums vs the reference's fp32 `torch.sum`: the reference itself is only good to ~1e-6
                # relative per sum; ratios of sums (r2, explained variance) amplify that
                rtol, atol = (1e-5, 1e-6) if exp.dtype == np.float32 else (1e-10, 1e-12)
            np.testing.assert_allclose(arr, exp, rtol=rtol, atol=atol, equal_nan=True, err_msg=f"{fn} {kwargs} out{i}")
