"""Replay the docstring examples of the REFERENCE's source files against metrics_b200 (build container only).

    MB200_REF_CPU_KERNELS=1 PYTHONPATH=tests/reference_runtime python tests/reference_runtime/run_doctests.py [subdir ...]

The reference's files are only PARSED (ast) for their docstrings — none of its code is imported or executed; the examples
import `torchmetrics...`, which sitecustomize.py aliases to `metrics_b200`, and run on CPU tensors through the kernel
stand-ins (cpu_kernels.py), so this checks the documented behaviour of the public classes / functionals: constructor
signatures, printed result values to 4 decimals, dict keys, shapes.  Like the reference's own doctest run (src/conftest.py)
every docstring starts from `torch.manual_seed(42)`.  Files whose metric is outside the scope are skipped by name."""
import ast
import doctest
import os
import random
import sys
import warnings

import numpy as np
import torch

SRC = "/root/reference/src/torchmetrics"
IN_SCOPE = {
    "classification": ["accuracy", "auroc", "average_precision", "cohen_kappa", "confusion_matrix", "exact_match", "f_beta",
                       "group_fairness", "hamming", "jaccard", "logauc", "matthews_corrcoef", "negative_predictive_value",
                       "precision_fixed_recall", "precision_recall", "precision_recall_curve", "recall_fixed_precision", "roc",
                       "sensitivity_specificity", "specificity", "specificity_sensitivity", "stat_scores"],
    "regression": ["explained_variance", "log_cosh", "log_mse", "mae", "mape", "minkowski", "mse", "r2", "rse",
                   "symmetric_mape", "tweedie_deviance", "wmape", "csi", "kl_divergence"],
    "wrappers": ["classwise"],
    "utilities": ["data", "compute", "distributed", "enums"],
    "detection": ["mean_ap"],
}
IN_SCOPE["functional/classification"] = IN_SCOPE["classification"]
IN_SCOPE["functional/regression"] = IN_SCOPE["regression"]
FLAGS = doctest.ELLIPSIS | doctest.NORMALIZE_WHITESPACE
SKIP_IF_MENTIONS = ("plot(", "plt.", "matplotlib", "+SKIP")  # plotting is out of scope; +SKIP examples are skipped upstream too


def docstrings(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.Module)):
            doc = ast.get_docstring(node, clean=True)
            if doc and ">>>" in doc:
                yield getattr(node, "name", "<module>"), doc


def module_globals(path):
    """What `--doctest-modules` gives every example for free: the names the file imports at top level (`torch`, `Tensor`,
    ...).  Each top-level import statement of the reference file is executed on its own (with `torchmetrics` aliased to this
    package); the ones naming reference internals this package does not have are simply absent."""
    globs = {}
    source = open(path).read()
    for node in ast.parse(source).body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            names = node.names
            for alias in names:  # one name at a time, so that a missing internal does not take its siblings with it
                single = ast.ImportFrom(module=node.module, names=[alias], level=node.level) if isinstance(node, ast.ImportFrom) \
                    else ast.Import(names=[alias])
                try:
                    exec(compile(ast.fix_missing_locations(ast.Module(body=[single], type_ignores=[])), path, "exec"), globs)
                except Exception:
                    pass
    # ... and the names the file itself defines: here, whatever this package's module of the same dotted path defines
    import importlib

    dotted = "torchmetrics." + os.path.relpath(path, SRC)[:-3].replace(os.sep, ".")
    try:
        globs.update({k: v for k, v in vars(importlib.import_module(dotted)).items() if not k.startswith("__")})
    except Exception as err:
        print(f"   (no module {dotted}: {err})")
    return globs


def run_file(path, verbose):
    parser, tried, failed, bad = doctest.DocTestParser(), 0, 0, []
    base = module_globals(path)
    for name, doc in docstrings(path):
        examples = [e for e in parser.get_examples(doc)]
        # drop plotting blocks: an example that mentions plotting, and everything after it in that docstring
        for i, ex in enumerate(examples):
            if any(tok in ex.source for tok in SKIP_IF_MENTIONS):
                examples = examples[:i]
                break
        if not examples:
            continue
        test = doctest.DocTest(examples, dict(base), f"{os.path.relpath(path, SRC)}::{name}", path, 0, doc)
        random.seed(42), np.random.seed(42), torch.manual_seed(42)
        out = []
        runner = doctest.DocTestRunner(verbose=False, optionflags=FLAGS)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            result = runner.run(test, out=out.append, clear_globs=True)
        tried += result.attempted
        failed += result.failed
        if result.failed:
            bad.append((test.name, "".join(out)))
    if verbose:
        for name, text in bad:
            print("-" * 100, "\n", name, "\n", text[-1800:])
    return tried, failed, [b[0] for b in bad]


def main(argv):
    verbose = "-v" in argv
    subdirs = [a for a in argv if not a.startswith("-")] or list(IN_SCOPE)
    total = [0, 0]
    for sub in subdirs:
        for stem in IN_SCOPE[sub]:
            path = os.path.join(SRC, sub, stem + ".py")
            if not os.path.exists(path):
                continue
            tried, failed, names = run_file(path, verbose)
            total[0] += tried
            total[1] += failed
            print(f"{sub}/{stem}.py: {tried - failed}/{tried} examples ok" + (f"   FAILED in: {', '.join(n.split('::')[1] for n in names)}" if names else ""))
    print(f"TOTAL: {total[0] - total[1]}/{total[0]} examples ok")


if __name__ == "__main__":
    main(sys.argv[1:])
