"""BASELINE.json configs[1]: fused warp + group-wise correlation + view aggregation microbench.

G = 8, C in {8,16,32}, D in {8,16,32,48}, 5 views, feature maps from 160x128 up to 1600x1184.
For every point: device time of the fused kernel (CUDA events, L2 flushed before each launch),
algorithmic GB/s = 4*(5*C*H*W + D*H*W + 4*H*W + 8*D*H*W) / time, fraction of the
measured HBM copy peak, the same composition in eager PyTorch on the same GPU (the oracle restatement of
upstream's homo_warping_new + correlation + aggregation, TF32 off) and the max relative difference
between the two (the parity bar is 1e-4).

    python tools/microbench.py [--quick] > profiles/r1_microbench.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import capi, hotpath, ops, synthetic  # noqa: E402
from oracle import hotpath as ohp  # noqa: E402


def timed(fn, flush, reps, warm=2):
    for _ in range(warm):          # the caching allocator needs two live output buffers before it stops calling cudaMalloc
        out = fn()
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts[1:]) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma list of WxH:C:D points, e.g. 800x592:32:8,1600x1184:32:16")
    ap.add_argument("--no-ref", action="store_true", help="skip the eager PyTorch composition")
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except (OSError, KeyError, ValueError):
        peak, src = 6650.0, "fallback"
    hp = hotpath.CudaHotPath("f32", native_projection=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sizes = [(160, 128), (200, 148), (320, 256), (400, 296), (800, 592), (1600, 1184)]
    grid = [(C, D) for C in (8, 16, 32) for D in (8, 16, 32, 48)]
    if a.quick:
        sizes, grid = [(200, 148), (800, 592)], [(8, 8), (32, 48)]
    points = [(W, H, C, D) for (W, H) in sizes for (C, D) in grid]
    if a.only:
        points = [(int(wh.split("x")[0]), int(wh.split("x")[1]), int(c), int(d))
                  for wh, c, d in (p.split(":") for p in a.only.split(","))]
    rows = []
    G, V = 8, 5
    with torch.no_grad():
        for (W, H, C, D) in points:
            out_bytes = 4.0 * G * D * H * W
            if out_bytes > 6e9:
                continue
            feats, cams, hyp, wts = synthetic.microbench_inputs(C, D, H, W, views=V, seed=C + D, device=dev)
            proj = hp.relative_projection(cams)          # 4x4 algebra once, outside the timed kernel
            ms, got = timed(lambda: ops.warp_corr_agg(feats[0], feats[1:], proj, hyp, capi.HYP_TENSOR, None, wts, D, G, False)[0],
                            flush, 3)
            row = {"W": W, "H": H, "C": C, "D": D, "G": G, "views": V}
            by = 4.0 * (V * C * H * W + D * H * W + (V - 1) * H * W + G * D * H * W)
            row.update(ms=ms, algorithmic_MB=by / 1e6, GBs=by / ms / 1e6, frac_of_hbm_peak=by / ms / 1e6 / peak)
            # the same call on channels-last maps (what the channels_last FPN of this repo emits): tiled TMA kernel
            fcl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
            ms_cl, got_cl = timed(lambda: ops.warp_corr_agg(fcl[0], fcl[1:], proj, hyp, capi.HYP_TENSOR, None, wts, D, G, False)[0],
                                  flush, 3)
            row.update(nhwc_ms=ms_cl, nhwc_GBs=by / ms_cl / 1e6, nhwc_frac_of_hbm_peak=by / ms_cl / 1e6 / peak,
                       nhwc_vs_nchw_rel_diff=float((got_cl - got).abs().max() / got.abs().max()))
            del fcl, got_cl
            # eager PyTorch reference on the same device; skip the sizes whose materialised warped volumes are huge
            if 4.0 * C * D * H * W < 3e9 and not a.no_ref:
                def ref():
                    sims = [ohp.view_similarity(feats[0], feats[v], cams[:, 0], cams[:, v], hyp, G) for v in range(1, V)]
                    return ohp.weighted_aggregate(sims, [wts[:, i:i + 1] for i in range(V - 1)])
                rms, want = timed(ref, flush, 1)
                row.update(ref_eager_ms=rms, ref_GBs=by / rms / 1e6, speedup_vs_eager=rms / ms,
                           rel_max_diff=float((got - want).abs().max() / want.abs().max()))
                del want
            rows.append(row)
            del feats, hyp, wts, got
    print(json.dumps({"config": "BASELINE.json configs[1]", "hbm_peak_GBs": peak, "peak_source": src, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
