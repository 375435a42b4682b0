"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain numpy restatement of the reference algorithms on the hot path (TorchMetrics 1.7.0dev under
/root/reference/src/torchmetrics), each function citing the reference file:line it follows.  It exists so that the
CUDA kernels can be checked on a GPU box where /root/reference is not present.

Rules:
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it;
  * nothing under `metrics_b200/` imports it — the product has no CPU path;
  * it is pinned against golden vectors generated from the real reference (tests/golden/make_golden.py, run in the
    build container where /root/reference exists; the reference is imported through a documented stand-in for its
    missing `lightning_utilities` dependency, tests/golden/_standins/).

Parity status: classification + curve families: PINNED (goldens produced by the unmodified reference).
Detection mAP: see oracle/coco_map.py header ("parity partially pinned": pycocotools is absent everywhere).
"""
