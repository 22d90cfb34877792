"""Where the stage-3 fused warp kernel's time goes: the same launch with parts switched off (EFFIMVS_WARP_DEBUG) and with a
warm / cold L2.  Profiling aid; the debug variants compute wrong results by design.

    python tools/warp_probe.py [stage 2|3] [reps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import capi, hotpath, ops, synthetic  # noqa: E402


def main():
    stage = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    dev = "cuda"
    torch.manual_seed(0)
    hp = hotpath.CudaHotPath("f32", native_projection=True)
    s = synthetic.make_sample("dtu", seed=0, device=dev)
    V = s["imgs"].shape[1]
    Hf, Wf = s["imgs"].shape[-2:]
    C, D, H, W = [(16, 8, Hf // 4, Wf // 4), (8, 8, Hf // 2, Wf // 2)][stage - 2]
    feats = [torch.randn(1, C, H, W, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(V)]
    proj = hp.relative_projection(s["proj_matrices"]["stage{}".format(stage)])
    wts = torch.rand(1, V - 1, H, W, device=dev)
    Es, Ks = synthetic.camera_ring(V, W, H)
    depths = {"noise": 680.0 + 40 * torch.rand(1, 1, H, W, device=dev),
              "smooth": synthetic.render_plane_scene(Es[:1], Ks, W, H, noise=0.0)[0].to(dev).reshape(1, 1, H, W)}
    iv = torch.full((1,), (1 / 425.0 - 1 / 935.0) / 384 * (2 if stage == 2 else 1), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    by = 4.0 * (V * C * H * W + H * W + (V - 1) * H * W + 2 * D * H * W)

    def timed(cur, cold):
        fn = lambda: ops.warp_corr_agg(feats[0], feats[1:], proj, cur, capi.HYP_LOCAL, iv, wts, D, 1, True)   # noqa: E731
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            if cold:
                flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    variants = [("tile kernel", {"EFFIMVS_WARP_SEG": "0"}), ("segment form", {"EFFIMVS_WARP_SEG": "1"}), ("segment form, no prefetch", {"EFFIMVS_WARP_SEG": "1", "EFFIMVS_WARP_DEBUG": "16"}),
                ("segment form, no footprint loads", {"EFFIMVS_WARP_SEG": "1", "EFFIMVS_WARP_DEBUG": "32"}), ("segment form, no lookups", {"EFFIMVS_WARP_SEG": "1", "EFFIMVS_WARP_DEBUG": "64"}),
                ("segment form, neither", {"EFFIMVS_WARP_SEG": "1", "EFFIMVS_WARP_DEBUG": "96"}), ("segment form, fast coordinates", {"EFFIMVS_WARP_SEG": "1", "EFFIMVS_WARP_FAST_COORDS": "1"})]
    for name, env in variants:
        for k in ("EFFIMVS_WARP_SEG", "EFFIMVS_WARP_DEBUG", "EFFIMVS_WARP_FAST_COORDS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        row = {"variant": name, "stage": stage}
        for dk, cur in depths.items():
            for cold in (True, False):
                ms = timed(cur, cold)
                row["{}_{}_ms".format(dk, "cold" if cold else "warm")] = round(ms, 4)
                row["{}_{}_GBs".format(dk, "cold" if cold else "warm")] = round(by / ms / 1e6, 1)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
