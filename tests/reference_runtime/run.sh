#!/bin/bash
# Run the REFERENCE's own runtime tests (tests/unittests/bases/*) against metrics_b200's Metric / MetricCollection runtime:
# `torchmetrics` is aliased to `metrics_b200` in every interpreter (sitecustomize.py in this directory), the reference's
# missing test-only dependency `cachier` is stubbed.  Needs /root/reference, i.e. only works in the build container; nothing
# from the reference is copied.  Tests that need a CPU implementation of a metric (there is none by design) or classes
# outside the scope (image / clustering / PearsonCorrCoef) fail with NativeLibraryError / NotImplementedError.
HERE="$(cd "$(dirname "$0")" && pwd)"
printf "[pytest]\naddopts =\n" > /tmp/mb200_ref_pytest.ini
cd /tmp
for f in test_metric test_composition test_hashing test_ddp test_collections; do
    echo "=== bases/$f.py"
    USE_PYTEST_POOL=1 PYTHONPATH="$HERE:/root/reference/tests" python -m pytest -c /tmp/mb200_ref_pytest.ini --rootdir /tmp \
        /root/reference/tests/unittests/bases/$f.py -q --no-header -p no:cacheprovider 2>&1 | grep "passed\|failed" | tail -2
done
