"""Times delta_head at the three DTU stage shapes (CUDA events, L2 flushed, the host kept ahead of the device)."""
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
    t = torch.randn(1, h, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    w, b = torch.randn(1, h, 3, 3, device=dev) * 0.1, torch.randn(1, device=dev)
    inv = torch.rand(1, 1, H, W, device=dev)
    lo, hi = torch.tensor([1 / 935.0], device=dev), torch.tensor([1 / 425.0], device=dev)
    ts = []
    for _ in range(14):
        flush.zero_(); flush.zero_(); flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.delta_head(t, w, b, inv, lo, hi)
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ts = sorted(ts[2:])
    res.append("h={} {}x{}: {:.1f} us".format(h, H, W, ts[len(ts) // 2] * 1e3))
print("; ".join(res))
