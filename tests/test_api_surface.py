"""CPU: the public surface — parameter names and defaults of every shared class constructor / `update` / `compute` /
functional / utility, and the class metadata (`higher_is_better`, `full_state_update`, plot bounds, legend names) — equals the
reference's (tests/golden/api_surface.json, dumped from the unmodified reference by tests/golden/make_signatures.py)."""
import importlib.util
import json
import os

from tests.conftest import GOLDEN_DIR

# the one deliberate metadata deviation: kernel launches carry no autograd graph
NOT_DIFFERENTIABLE_HERE = "regression."


def _mine():
    spec = importlib.util.spec_from_file_location("make_signatures", os.path.join(GOLDEN_DIR, "make_signatures.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module.surface("metrics_b200")


def test_shared_surface_matches_the_reference():
    ref = json.load(open(os.path.join(GOLDEN_DIR, "api_surface.json")))
    mine = _mine()
    shared = sorted(set(ref) & set(mine))
    assert len(shared) >= 240, len(shared)
    problems = []
    for key in shared:
        r, m = ref[key], mine[key]
        for part in ("init", "update", "compute", "call"):
            if part in r and r[part] is not None and m.get(part) is not None and r[part] != m[part]:
                problems.append(f"{key}.{part}: reference {r[part]} != {m[part]}")
        for attr, value in r.get("attrs", {}).items():
            got = m["attrs"][attr]
            if attr == "is_differentiable" and key.startswith(NOT_DIFFERENTIABLE_HERE):
                assert got == "False", key
                continue
            if got != value:
                problems.append(f"{key}.{attr}: reference {value} != {got}")
    assert not problems, "\n".join(problems)


def test_every_in_scope_name_of_the_reference_exists_here():
    ref = json.load(open(os.path.join(GOLDEN_DIR, "api_surface.json")))
    mine = _mine()
    # what the reference has in these sub-packages and this package deliberately does not:
    out_of_scope = (
        # other arithmetic than the three state kinds of the path
        "Calibration", "calibration", "Hinge", "hinge", "Ranking", "ranking", "Coverage", "coverage", "IntersectionOverUnion",
        "Panoptic", "panoptic", "Pearson", "pearson", "Spearman", "spearman", "Kendall", "kendall", "Cosine", "cosine",
        "Concordance", "concordance", "KLDivergence", "kl_divergence", "NormalizedRoot",
        "normalized_root",
        # deprecated upstream (removed in 1.7) and built on the legacy input-format machinery
        "Dice", "dice", "_input_format_classification", "_check_classification_inputs", "_check_num_classes",
        "_check_shape_and_type_consistency", "_basic_input_validation", "_check_top_k", "_input_squeeze", "_check_for_empty_tensors",
        # other wrappers, retrieval helpers, names a module merely imports
        "BootStrapper", "FeatureShare", "Transformer", "MetricTracker", "MinMaxMetric", "MultioutputWrapper", "MultitaskWrapper",
        "wrappers.Running", "_check_retrieval", "_try_proceed_with_timeout", "_simple_gather_all_tensors",
        "_top_k_with_half_precision_support", "utilities.checks.DataType", "utilities.checks.Metric", "utilities.checks.select_topk",
        "utilities.checks.to_onehot", "utilities.data.TorchMetricsUserWarning", "utilities.data.rank_zero_warn",
    )
    missing = sorted(k for k in ref if k not in mine and not any(tok in k for tok in out_of_scope))
    assert not missing, missing


def test_state_registries_match_the_reference():
    """State names, default shapes / dtypes, `dist_reduce_fx` and persistence of every shared metric configuration: a
    `state_dict()` written by the reference must load here and the cross-rank reductions must be the same."""
    import warnings

    spec = importlib.util.spec_from_file_location("make_signatures", os.path.join(GOLDEN_DIR, "make_signatures.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mine = module.state_registry("metrics_b200")
    ref = json.load(open(os.path.join(GOLDEN_DIR, "state_registry.json")))
    shared = sorted(set(ref) & set(mine))
    assert len(shared) >= 165, len(shared)
    different = {k: (ref[k], mine[k]) for k in shared if ref[k] != mine[k]}
    assert not different, different
