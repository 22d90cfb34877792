"""Times the geometric-consistency filter kernel (a13-a15) at the DTU shape with CUDA events.

    python tools/fusion_bench.py [views=10] [reps=7]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import fusion, ops, synthetic  # noqa: E402


def main():
    v = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    dev = "cuda"
    h, w = 1184, 1600
    torch.manual_seed(0)
    E, K = synthetic.camera_ring(v + 1, w, h)
    cams = synthetic.stage_cameras(E, K, 1)["stage4"].to(dev)
    depths = synthetic.render_plane_scene(E, K, w, h).to(dev)      # smooth surface + 0.15 mm noise: most pixels consistent
    conf = torch.rand(1, h // 2, w // 2, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    by = 4.0 * h * w * (1 + v + 1) + h * w * (1 + 4 + 12)
    for inv in (True, False):
        # the inverses: torch's LU (upstream's .inverse()) or the library's own fp64 kernel; both outside the timed kernel
        inv_c = fusion.inverse_cameras(cams[:, 0], cams[:, 1:]) if inv else ops.fusion_invert_cameras(cams[:, 0], cams[:, 1:])
        ts = []
        for _ in range(reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.fusion_filter(depths[0][None, None], depths[1:][None, :, None], conf, cams[:, 0], cams[:, 1:], inv_c, 2.0, 6.0, 2, 0.3, False, False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        ms = ts[len(ts) // 2]
        print(json.dumps({"kernel": "fusion_kernel", "h": h, "w": w, "v": v, "torch_inverse": inv, "ms": ms, "bytes": by, "GBs": by / ms / 1e6}))


def dtu_variant(v=10, reps=7):
    """the DTU pipeline's NumPy/cv2 filter variant (csrc/dtu_filter.cu), same shape"""
    dev = "cuda"
    h, w = 1184, 1600
    E, K = synthetic.camera_ring(v + 1, w, h)
    depths = synthetic.render_plane_scene(E, K, w, h).to(dev)
    conf = torch.rand(h, w, device=dev)
    K32, E32 = K.numpy().astype("float32"), E.numpy().astype("float32")
    mats = fusion.dtu_camera_pack(K32, E32[0], [K32] * v, list(E32[1:])).to(dev)
    import math
    import numpy as np
    td = [i * 0.5 for i in range(1, 11)]
    tf = [float(np.float32(math.log(max(i, 1.05), 10) * 0.25)) for i in range(1, 11)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    by = 4.0 * h * w * (1 + v + 1) + h * w * (1 + 1 + 4 + 12)
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.dtu_filter(depths[0], depths[1:], conf, mats, td, tf, 1, 11, 0.5, 0.75, False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    ms = ts[len(ts) // 2]
    print(json.dumps({"kernel": "dtu_filter_kernel", "h": h, "w": w, "v": v, "ms": ms, "bytes": by, "GBs": by / ms / 1e6}))


if __name__ == "__main__":
    main()
    dtu_variant()
