"""Device-resident timing of every kernel family at a large size: achieved algorithmic GB/s vs the measured HBM peak.
One JSON document (profiles/r01_kernel_rooflines.json).  Inputs are larger than L2 or rotated so nothing is re-read from it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metrics_b200 import _native  # noqa: E402
from metrics_b200.functional.classification.stat_scores import stat_scores_workspace  # noqa: E402

dev = torch.device("cuda", 0)
peak = 6574.8
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=20, warm=3, inner=1):
    """Best-of-`reps` CUDA-event time of `inner` back-to-back calls, per call (inner > 1 for kernels shorter than a launch)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


out = {"peak_gbs_measured_copy": peak, "kernels": {}}


def record(name, ms, algo_bytes, note):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    out["kernels"][name] = {"ms": ms, "algorithmic_bytes": algo_bytes, "achieved_gbs": gbs, "frac_of_measured_peak": gbs / peak,
                            "note": note}


g = torch.Generator(device=dev).manual_seed(0)
# K1b: multiclass stat scores, cfg2 shape (65536 x 1000 bf16), eight rotating batches (1 GB >> L2)
N, C = 65536, 1000
lg = [torch.randn(N, C, generator=g, device=dev).bfloat16() for _ in range(8)]
tg = [torch.randint(0, C, (N,), generator=g, device=dev) for _ in range(8)]
st = [torch.zeros(C, dtype=torch.int64, device=dev) for _ in range(4)]
ws = stat_scores_workspace(C, dev)
k = [0]


def k1b():
    i = k[0] = (k[0] + 1) % 8
    _native.multiclass_stat_scores_update_(*st, ws, lg[i], tg[i], C, None, False, None)


record("K1b multiclass_stat_scores_update [65536,1000] bf16", timed(k1b, inner=64), N * C * 2 + N * 8,
       "64 back-to-back launches over 8 distinct batches (1 GB); logits + labels read once; states in L2")
del lg

# K2: binary counts over 2^26 f32 probabilities + int64 targets (805 MB)
n = 1 << 26
p = torch.rand(n, generator=g, device=dev)
t = torch.randint(0, 2, (n,), generator=g, device=dev)
record("K2 binary_stat_counts 2^26 f32 + i64", timed(lambda: _native.binary_stat_counts(p, t, 1, 0.5, None, False)),
       n * 12, "range-flag pass + counting pass: scores are read twice (8 B/elem) + 8 B labels once")
# K2 multilabel: [2^20, 64] f32 probabilities + int64 targets, per-label counters
pm = p[: (1 << 26)].view(1 << 20, 64)
tm = t[: (1 << 26)].view(1 << 20, 64)
record("K2 multilabel counts [2^20, 64] f32 + i64", timed(lambda: _native.binary_stat_counts(pm, tm, 64, 0.5, None, False)),
       n * 12, "generic grouped kernel: group index per element, runs of equal group kept in registers")
# K6: sigmoid_if_logits (flag pass + apply pass): read 4 + read 4 + write 4
record("K6 sigmoid_if_logits 2^26 f32", timed(lambda: _native.sigmoid_if_logits(p)), n * 8,
       "algorithmic = read + write; the global logits vote costs a second read (12 B/elem of traffic)")
# K4: binned curve update, T = 200 thresholds
thr = torch.linspace(0, 1, 200, device=dev)
record("K4 binned_curve_update 2^26 f32, T=200", timed(lambda: _native.binned_curve_update(p, t, thr, 1), reps=5), n * 12,
       "one pass: 8-step binary search per element + shared-memory histogram")
# K9: regression sums (MSE) over 2 x 2^26 f32
q = torch.rand(n, generator=g, device=dev)
record("K9 regression_sums (MSE) 2 x 2^26 f32", timed(lambda: _native.regression_sums(p, q, 0)), n * 8, "two inputs read once")
del q
# K3/K5: exact binary curve evaluation, 10^7 samples
n7 = 10_000_000
p7, t7 = p[:n7].contiguous(), t[:n7].contiguous()
record("K3/K5 curve_evaluate 1e7 f32 (pack + 4-pass sort + scan), (key, label) pairs",
       timed(lambda: _native.curve_evaluate(p7, t7, 1, unit_range=False), reps=10), 150e6,
       "SURVEY 8(d) figure: each 5-byte record read once, written once, scanned once; the working set (50 MB) is L2-resident")
record("K3/K5 curve_evaluate 1e7 f32, label in bit 0 of the key (metric states: non-negative scores)",
       timed(lambda: _native.curve_evaluate(p7, t7, 1, unit_range=True), reps=10), 150e6,
       "same algorithmic figure; the passes move 4-byte keys, the last one splits (key, label)")
# K11: fused stat scores + softmax store, cfg5-shaped batches scaled up: [65536, 1000] f32, 4 rotating batches (1 GB)
lg32 = [torch.randn(N, C, generator=g, device=dev) for _ in range(4)]
tg32 = [torch.randint(0, C, (N,), generator=g, device=dev) for _ in range(4)]
kk = [0]


def k11():
    i = kk[0] = (kk[0] + 1) % 4
    _native.multiclass_stats_softmax_update_(*st, ws, lg32[i], tg32[i], C, False)


record("K11 stats+softmax fused update [65536,1000] f32", timed(k11, inner=16), N * C * 4 * 2 + N * 8,
       "one read + one write of the batch (the unfused members read it three times and write it once)")


def unfused():
    i = kk[0] = (kk[0] + 1) % 4
    _native.multiclass_stat_scores_update_(*st, ws, lg32[i], tg32[i], C, None, False, None)
    _native.softmax_if_logits(lg32[i])


def k6s():
    i = kk[0] = (kk[0] + 1) % 4
    _native.softmax_if_logits(lg32[i])


record("K6 softmax_if_logits [65536,1000] f32", timed(k6s, inner=16), N * C * 4 * 2,
       "speculative single pass: row in registers, one read + one write")
record("K1b + K6 unfused on the same batches [65536,1000] f32", timed(unfused, inner=16), N * C * 4 * 2 + N * 8,
       "same algorithmic bytes; three reads + one write of traffic")
del lg32
# K3 multiclass: 1000 one-vs-rest curves over 16384 samples (cfg5, one rank's share)
pm5 = torch.softmax(torch.randn(16384, 1000, generator=g, device=dev), 1)
tm5 = torch.randint(0, 1000, (16384,), generator=g, device=dev)
record("K3/K5 curve_evaluate [16384,1000] f32, 1000 segments", timed(lambda: _native.curve_evaluate(pm5, tm5, 1000, unit_range=True), reps=10),
       16384 * 1000 * 15, "same per-record figure as the binary case (5-byte record read, written, scanned)")
del pm5
# K12: instance masks, 64 images of 480 x 640 with 100 detections x 20 ground truths each, one class (every pair intersected)
n_img12, d12, g12, hw12 = 64, 100, 20, 480 * 640
dm12 = torch.rand(n_img12 * d12, hw12, generator=g, device=dev) > 0.7
gm12 = torch.rand(n_img12 * g12, hw12, generator=g, device=dev) > 0.7
record("K12 mask_pack_bits 6400 masks of 480x640 (bool -> 1 bit/pixel + areas)", timed(lambda: _native.mask_pack_bits(dm12), reps=5),
       dm12.numel() * (1 + 1 / 8), "one read of the byte masks, one write of the bit rows")
dw12, _ = _native.mask_pack_bits(dm12)
gw12, _ = _native.mask_pack_bits(gm12)
del dm12, gm12
w12 = dw12.shape[1]
i64 = lambda x: torch.tensor(x, dtype=torch.int64, device=dev)  # noqa: E731
args12 = (dw12.reshape(-1), torch.arange(n_img12 * d12, device=dev) * w12, gw12.reshape(-1), torch.arange(n_img12 * g12, device=dev) * w12,
          (torch.arange(n_img12 + 1, device=dev) * d12).int(), (torch.arange(n_img12 + 1, device=dev) * g12).int(),
          torch.full((n_img12,), w12, dtype=torch.int32, device=dev), torch.zeros(n_img12 * d12, dtype=torch.int64, device=dev),
          torch.zeros(n_img12 * g12, dtype=torch.int64, device=dev), True, torch.arange(n_img12, device=dev) * (d12 * g12),
          n_img12 * d12 * g12, d12 * g12)
record("K12 mask_pair_intersections 64 img x (100 x 20) pairs of 480x640 masks",
       timed(lambda: _native.mask_pair_intersections(*args12), reps=5), (dw12.numel() + gw12.numel()) * 4,
       "algorithmic: every bit row read once (the pairs re-read them from L2: 128 000 pairs x 2 x 38.4 KB = 9.8 GB of L2 reads)")
del dw12, gw12, args12
# K13: per-row KL divergence, [65536, 1000] f32 probabilities (and the reference's own op chain on the same tensors)
p13 = torch.rand(65536, 1000, generator=g, device=dev) + 1e-3
q13 = torch.rand(65536, 1000, generator=g, device=dev) + 1e-3
record("K13 kl_divergence_rows [65536,1000] f32", timed(lambda: _native.kl_divergence_rows(p13, q13, False), reps=10),
       2 * p13.numel() * 4, "one read of p and q (the second pass over a row hits L1/L2)")


def kld_aten():
    a = p13 / p13.sum(-1, keepdim=True)
    b = q13 / q13.sum(-1, keepdim=True)
    r = a * torch.log(a / b)
    r[a == 0] = 0.0
    return r.sum(-1)


record("K13 reference op chain in ATen on the same tensors", timed(kld_aten, reps=10), 2 * p13.numel() * 4, "eight passes with [N, d] temporaries")
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(json.dumps(out, indent=1))
