"""Run ONE kernel family a few times at a large size, for an ncu capture:

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/prof_<name> \
        python benchmarks/prof_one.py <name>        # k4 | k1b | k3 | fused | k2 | k6 | cfg5_eval | k12 | k13
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metrics_b200 import _native  # noqa: E402

dev = torch.device("cuda", 0)
name = sys.argv[1]
g = torch.Generator(device=dev).manual_seed(0)
if name in ("k4", "k2", "k6"):
    n = 1 << 26
    p = torch.rand(n, generator=g, device=dev)
    t = torch.randint(0, 2, (n,), generator=g, device=dev)
    thr = torch.linspace(0, 1, 200, device=dev)
    fn = {"k4": lambda: _native.binned_curve_update(p, t, thr, 1),
          "k2": lambda: _native.binary_stat_counts(p, t, 1, 0.5, None, False),
          "k6": lambda: _native.sigmoid_if_logits(p - 0.5)}[name]
elif name == "k1b":
    from metrics_b200.functional.classification.stat_scores import stat_scores_workspace

    N, C = 65536, 1000
    lg = torch.randn(N, C, generator=g, device=dev).bfloat16()
    tg = torch.randint(0, C, (N,), generator=g, device=dev)
    st = [torch.zeros(C, dtype=torch.int64, device=dev) for _ in range(4)]
    ws = stat_scores_workspace(C, dev)
    fn = lambda: _native.multiclass_stat_scores_update_(*st, ws, lg, tg, C, None, False, None)  # noqa: E731
elif name == "fused":
    from metrics_b200.functional.classification.stat_scores import stat_scores_workspace

    N, C = 65536, 1000
    lg = torch.randn(N, C, generator=g, device=dev)
    tg = torch.randint(0, C, (N,), generator=g, device=dev)
    st = [torch.zeros(C, dtype=torch.int64, device=dev) for _ in range(4)]
    ws = stat_scores_workspace(C, dev)
    fn = lambda: _native.multiclass_stats_softmax_update_(*st, ws, lg, tg, C, False)  # noqa: E731
elif name == "k3":
    p = torch.rand(10_000_000, generator=g, device=dev)
    t = torch.randint(0, 2, (10_000_000,), generator=g, device=dev)
    fn = lambda: _native.curve_evaluate(p, t, 1, unit_range=True)  # noqa: E731
elif name == "cfg5_eval":
    p = torch.softmax(torch.randn(16384, 1000, generator=g, device=dev), 1)
    t = torch.randint(0, 1000, (16384,), generator=g, device=dev)
    fn = lambda: _native.curve_evaluate(p, t, 1000, unit_range=True)  # noqa: E731
elif name == "k12":
    masks = torch.rand(1200, 480 * 640, generator=g, device=dev) > 0.7
    fn = lambda: _native.mask_pack_bits(masks)  # noqa: E731
elif name == "k13":
    p = torch.rand(65536, 1000, generator=g, device=dev) + 1e-3
    q = torch.rand(65536, 1000, generator=g, device=dev) + 1e-3
    fn = lambda: _native.kl_divergence_rows(p, q, False)  # noqa: E731
else:
    raise SystemExit(f"unknown kernel family {name}")
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fn()
e1.record()
torch.cuda.synchronize()
print(name, "ms per call:", e0.elapsed_time(e1) / 5)
