"""CPU: binned multi-threshold confusion matrices of the oracle vs the reference goldens."""
import numpy as np
import pytest
import torch

from oracle import curves as oc

THR = {"int11": torch.linspace(0, 1, 11).numpy(), "int200": torch.linspace(0, 1, 200).numpy(),
       "list": np.array([0.9, 0.1, 0.5, 0.3], np.float32), "tensor": np.array([0.0, 0.2, 0.7, 1.0], np.float32)}


@pytest.mark.parametrize("name", list(THR))
def test_binary_binned_confmat(golden_binned, name):
    g = golden_binned
    t = g["b/target"]
    np.testing.assert_array_equal(oc.binned_confmat(g["b/preds"], t, THR[name]), g[f"b/{name}/probs/confmat"])
    logits = oc.sigmoid_if_logits(g["b/logits"])
    got = oc.binned_confmat(logits, t[:3000], THR[name])
    # sigmoid outputs can differ by an ulp between libm and ATen: allow a handful of boundary flips
    assert np.abs(got - g[f"b/{name}/logits/confmat"]).max() <= 2


def test_multiclass_binned_confmat(golden_binned):
    g = golden_binned
    p = oc.softmax_if_logits(g["m/logits"])
    for name, thr in (("int7", torch.linspace(0, 1, 7).numpy()), ("list", np.array([0.05, 0.2, 0.6], np.float32))):
        got = oc.binned_confmat(p, g["m/target"], thr, 6)
        assert np.abs(got - g[f"m/{name}/confmat"]).max() <= 2
