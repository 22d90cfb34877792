"""Runs the hot-path kernels once each at the DTU stage shapes (after a warm-up launch), for
`ncu --set full -k regex:...` captures and for the per-kernel launch list in profiles/.

    python tools/profile_kernels.py [warp|reg|fusion|all] [precision]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import capi, fusion, hotpath, ops, synthetic  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
    dev = "cuda"
    torch.manual_seed(0)
    hp = hotpath.CudaHotPath(prec, native_projection=True)
    s = synthetic.make_sample("dtu", seed=0, device=dev)
    V = s["imgs"].shape[1]
    shapes = [(32, 48, 148, 200), (16, 8, 296, 400), (8, 8, 592, 800)]
    for rep in range(2):
        if what in ("warp", "all"):
            for i, (C, D, H, W) in enumerate(shapes):
                feats = [torch.randn(1, C, H, W, device=dev) for _ in range(V)]
                proj = hp.relative_projection(s["proj_matrices"]["stage{}".format(i + 1)])
                if i == 0:
                    planes = (1.0 / torch.linspace(1 / 935.0, 1 / 425.0, D, device=dev)).reshape(1, D)
                    sims, ent = ops.warp_corr_views(feats[0], feats[1:], proj, planes, capi.HYP_PLANES, D)
                    ops.weighted_agg(sims, torch.rand(1, V - 1, H, W, device=dev))
                    ops.softmax_regress_conf(sims[:, 0], planes, capi.HYP_PLANES)
                else:
                    cur = 680.0 + 40 * torch.rand(1, 1, H, W, device=dev)
                    iv = torch.full((1,), (1 / 425.0 - 1 / 935.0) / 384 * (2 if i == 1 else 1), device=dev)
                    sim, hyp = ops.warp_corr_agg(feats[0], feats[1:], proj, cur, capi.HYP_LOCAL, iv, torch.rand(1, V - 1, H, W, device=dev), D, 1, True)
                    vol = sim[:, 0]
                    ops.dynamic_cost(cur, vol, vol, iv, hyp[:, -1:].contiguous(), hyp[:, :1].contiguous(), 3)
        if what in ("reg", "all"):
            from util import dtu_model
            model = dtu_model(hp, dev)
            x = torch.randn(1, 1, 48, 148, 200, device=dev)
            hp.cost_regularization(model.cost_regularization, x)
            for (D, H, W) in ((8, 296, 400), (8, 592, 800)):
                hp.cross_scale(model.CSP_R[0], torch.randn(1, 1, D, H, W, device=dev), torch.randn(1, 1, D, H // 2, W // 2, device=dev))
        if what in ("fusion", "all"):
            h, w, v = 1184, 1600, 10
            E, K = synthetic.camera_ring(v + 1, w, h)
            cams = synthetic.stage_cameras(E, K, 1)["stage4"].to(dev)
            depths = (680.0 + 20 * torch.rand(v + 1, h, w, device=dev))
            fusion.filter_view(depths[0][None, None], torch.rand(1, h // 2, w // 2, device=dev), depths[1:][None, :, None],
                               cams[:, 0], cams[:, 1:], 2, 6, 2, 0.3, torch_inverse=False)
        torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
