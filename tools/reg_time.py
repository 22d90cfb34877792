"""Times the tensor-core regularization nets at the DTU shapes (CUDA events, L2 flushed): python tools/reg_time.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import hotpath  # noqa: E402
from util import dtu_model  # noqa: E402

dev = "cuda"
hp = hotpath.CudaHotPath("bf16x3", native_projection=True)
model = dtu_model(hp, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2] * 1e3


with torch.no_grad():
    x = torch.randn(1, 1, 48, 148, 200, device=dev)
    t0 = timed(lambda: hp.cost_regularization(model.cost_regularization, x))
    res = ["costreg {:.1f} us".format(t0)]
    for (D, H, W) in ((8, 296, 400), (8, 592, 800)):
        a, b = torch.randn(1, 1, D, H, W, device=dev), torch.randn(1, 1, D, H // 2, W // 2, device=dev)
        res.append("cost_up {}x{} {:.1f} us".format(H, W, timed(lambda: hp.cross_scale(model.CSP_R[0], a, b))))
print("; ".join(res))
