"""Dump the public call signatures and class metadata of the REFERENCE (build container only) to tests/golden/api_surface.json:
for every class / function of the in-scope sub-packages the parameter names + defaults of `__init__` / `__new__`, `update`,
`compute`, and `higher_is_better` / `is_differentiable` / `full_state_update` / plot bounds.  tests/test_api_surface.py holds
this package to it.  Usage: python tests/golden/make_signatures.py"""
import importlib
import inspect
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
SUBPACKAGES = ("classification", "regression", "detection", "wrappers", "functional.classification", "functional.regression",
               "utilities.data", "utilities.compute", "utilities.distributed", "utilities.checks")
ATTRS = ("higher_is_better", "is_differentiable", "full_state_update", "plot_lower_bound", "plot_upper_bound", "plot_legend_name")


def signature(fn):
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return None
    return [[n, "<required>" if p.default is inspect.Parameter.empty else repr(p.default)] for n, p in sig.parameters.items()
            if n != "self"]


def surface(pkg: str) -> dict:
    out = {}
    for sub in SUBPACKAGES:
        try:
            module = importlib.import_module(f"{pkg}.{sub}")
        except ImportError:
            continue
        for name in dir(module):
            obj = getattr(module, name)
            if name.startswith("__") or not getattr(obj, "__module__", "").startswith(pkg):
                continue
            key = f"{sub}.{name}"
            if isinstance(obj, type):
                out[key] = {"init": signature(obj.__new__ if "__new__" in obj.__dict__ else obj.__init__),
                            "update": signature(obj.update) if hasattr(obj, "update") else None,
                            "compute": signature(obj.compute) if hasattr(obj, "compute") else None,
                            "attrs": {a: repr(getattr(obj, a, None)) for a in ATTRS}}
            elif inspect.isfunction(obj):
                out[key] = {"call": signature(obj)}
    return out


def state_registry(pkg: str) -> dict:
    """{"Class[variant]": {state: {default shape/dtype or "list", reduce fn name, persistent}}} for every in-scope metric class
    that can be built from a small table of constructor arguments — what `state_dict()` keys and cross-rank reductions are."""
    import torch

    def describe(metric) -> dict:
        return {k: {"default": "list" if isinstance(v, list) else [list(v.shape), str(v.dtype)],
                    "reduce": getattr(metric._reductions[k], "__name__", None) if metric._reductions[k] is not None else None,
                    "persistent": metric._persistent[k]} for k, v in metric._defaults.items()}

    out = {}
    cls_pkg, reg_pkg = importlib.import_module(f"{pkg}.classification"), importlib.import_module(f"{pkg}.regression")
    floors = ("AtFixed", "SensitivityAt", "SpecificityAt")
    for name in sorted(dir(cls_pkg)):
        cls = getattr(cls_pkg, name)
        prefix = next((p for p in ("Binary", "Multiclass", "Multilabel") if name.startswith(p)), None)
        if not isinstance(cls, type) or prefix is None:
            continue
        for variant, extra in (("default", {}), ("thresholds", {"thresholds": 7}), ("samplewise", {"multidim_average": "samplewise"}),
                               ("micro", {"average": "micro"})):
            args = () if prefix == "Binary" else (3,)
            if "FBeta" in name:
                args = (2.0, *args)
            if any(f in name for f in floors):
                args = (*args, 0.5)
            if name in ("BinaryFairness", "BinaryGroupStatRates"):
                args = (2,)
            try:
                out[f"{name}[{variant}]"] = describe(cls(*args, **extra))
            except Exception:  # this variant's keyword does not exist for this class
                continue
    for name in sorted(dir(reg_pkg)):
        cls = getattr(reg_pkg, name)
        if isinstance(cls, type) and issubclass(cls, torch.nn.Module):
            for variant, extra in (("default", {}), ("multi", {"num_outputs": 3})):
                try:
                    special = {"MinkowskiDistance": {"p": 2.0}, "CriticalSuccessIndex": {"threshold": 0.5}}.get(name, {})
                    out[f"{name}[{variant}]"] = describe(cls(**special, **extra))
                except Exception:
                    continue
            if name == "CriticalSuccessIndex":
                out[f"{name}[sequence]"] = describe(cls(0.5, keep_sequence_dim=0))
    return out


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    sys.path.insert(0, os.path.join(HERE, "_standins"))
    sys.path.insert(0, "/root/reference/src")
    data = surface("torchmetrics")
    with open(os.path.join(HERE, "api_surface.json"), "w") as fh:
        json.dump(data, fh, indent=0, sort_keys=True)
    print("wrote api_surface.json:", len(data), "entries")
    states = state_registry("torchmetrics")
    with open(os.path.join(HERE, "state_registry.json"), "w") as fh:
        json.dump(states, fh, indent=0, sort_keys=True)
    print("wrote state_registry.json:", len(states), "metric configurations")
