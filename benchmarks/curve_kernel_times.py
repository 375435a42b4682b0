"""One curve_evaluate(1e7) call after warm-up — run under `ncu --metrics gpu__time_duration.sum` for a per-kernel list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metrics_b200 import _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator().manual_seed(0)
if c == 1:
    p = torch.rand(n, generator=g).cuda()
    t = torch.randint(0, 2, (n,), generator=g).cuda()
else:
    p = torch.rand(n, c, generator=g).cuda()
    t = torch.randint(0, c, (n,), generator=g).cuda()
for _ in range(2):
    _native.curve_evaluate(p, t, c, unit_range=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
_native.curve_evaluate(p, t, c, unit_range=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
