#!/bin/bash
# Run the REFERENCE's own unit tests against metrics_b200 (build container only: needs /root/reference; nothing from the
# reference is copied).  `torchmetrics` is aliased to `metrics_b200` in every interpreter (sitecustomize.py in this
# directory); test-only dependencies missing from the image (`cachier`, `permetrics`) are stubbed.
#
#   run.sh                                   runtime tests (tests/unittests/bases/*): Metric, MetricCollection, sync, hashing
#   run.sh <subdir> <test_module>...         e.g. `run.sh classification test_accuracy test_f_beta`, with the kernel wrappers of
#                                            `metrics_b200._native` replaced by torch-CPU stand-ins (cpu_kernels.py) so that the
#                                            reference's CPU-tensor tests reach our host layer: validation, formatting seams,
#                                            state handling, reducers, task wrappers, error messages.
#
#   MB200_REF_CPU_KERNELS=1 PYTHONPATH=tests/reference_runtime python tests/reference_runtime/run_doctests.py
#                                            the reference's docstring examples replayed against this package
#
# The stand-ins are TEST INFRASTRUCTURE (the product has no CPU path and raises on CPU tensors); they are only ever
# installed by sitecustomize.py when MB200_REF_CPU_KERNELS=1.
HERE="$(cd "$(dirname "$0")" && pwd)"
printf "[pytest]\naddopts =\n" > /tmp/mb200_ref_pytest.ini
cd /tmp
run() {  # <path under tests/unittests> [extra pytest args]
    local f=$1; shift
    PYTHONPATH="$HERE:/root/reference/tests" python -m pytest -c /tmp/mb200_ref_pytest.ini --rootdir /tmp \
        "/root/reference/tests/unittests/$f.py" -q --no-header -p no:cacheprovider "$@" 2>&1 | grep " passed\| failed\| error" | tail -1
}
if [ $# -eq 0 ]; then
    for f in test_metric test_composition test_hashing test_ddp test_collections; do
        echo "=== bases/$f: $(USE_PYTEST_POOL=1 run bases/$f)"
    done
else
    sub=$1; shift
    for f in "$@"; do
        echo "=== $sub/$f: $(USE_PYTEST_POOL=1 MB200_REF_CPU_KERNELS=1 run $sub/$f -k 'not plot and not nrmse and not _Absent')"
    done
fi
