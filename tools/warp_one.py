"""A few launches of the stage-3 tile warp kernel (channels-last features, local hypotheses) for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import capi, hotpath, ops, synthetic  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
hp = hotpath.CudaHotPath("bf16x3", native_projection=True)
s = synthetic.make_sample("dtu", seed=0, device=dev)
V = s["imgs"].shape[1]
C, D, H, W = 8, 8, 592, 800
feats = [torch.randn(1, C, H, W, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(V)]
proj = hp.relative_projection(s["proj_matrices"]["stage3"])
Es, Ks = synthetic.camera_ring(V, W, H)
cur = synthetic.render_plane_scene(Es[:1], Ks, W, H, noise=0.3)[0].to(dev).reshape(1, 1, H, W)
iv = torch.full((1,), (1 / 425.0 - 1 / 935.0) / 384, device=dev)
wts = torch.rand(1, V - 1, H, W, device=dev)
for _ in range(5):
    ops.warp_corr_agg(feats[0], feats[1:], proj, cur, capi.HYP_LOCAL, iv, wts, D, 1, True)
torch.cuda.synchronize()
print("ok")
