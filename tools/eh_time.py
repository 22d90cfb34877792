"""Times encoder_head at the three DTU stage shapes (CUDA events, L2 flushed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = []
for h, H, W in ((16, 592, 800), (32, 296, 400), (48, 148, 200)):
    cost, inv = torch.randn(1, 6, H, W, device=dev), torch.rand(1, 1, H, W, device=dev)
    wc1, bc1 = torch.randn(h, 6, 1, 1, device=dev), torch.randn(h, device=dev)
    wd1, bd1 = torch.randn(h, 1, 7, 7, device=dev), torch.randn(h, device=dev)
    ts = []
    for _ in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.encoder_head(cost, inv, wc1, bc1, wd1, bd1)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[2:])
    res.append("h={} {}x{}: {:.1f} us".format(h, H, W, ts[len(ts) // 2] * 1e3))
print("; ".join(res))
