"""Times the scene runner's two CUDA graphs (encode one image / cascade from cached features) in isolation."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import hotpath, scene, synthetic  # noqa: E402
from util import dtu_model  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = dtu_model(hotpath.CudaHotPath("bf16x3", native_projection=True), dev)
N, W, H = 8, 1600, 1184
imgs = torch.rand(N, 3, H, W).to(dev)
E, K = synthetic.camera_arc(N, W, H)
cams = {k: v[0].to(dev) for k, v in synthetic.stage_cameras(E, K, 1).items()}
dv = torch.linspace(1 / 935.0, 1 / 425.0, 384, device=dev)
r = scene.GraphedViewRunner(model, imgs, cams, dv, 4)


def t(g, n=10):
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(r.stream):
        a.record()
        for _ in range(n):
            g.replay()
        b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print("EFFIMVS_CONV2D={} encode graph {:.3f} ms, cascade graph {:.3f} ms".format(os.environ.get("EFFIMVS_CONV2D", "auto"), t(r.g_enc), t(r.g_fwd)))
