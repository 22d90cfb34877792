"""ORACLE (test infrastructure only -- never imported by the product path): NumPy restatement of upstream's
DTU geometric-consistency filter, test_dtu_dypcd.py:164-233 (reproject_with_depth,
check_geometric_consistency) and the aggregation of filter_depth, :261-309, :320-337.

Third-party arithmetic on this path: ``cv2.remap(depth_src, x_src, y_src, cv2.INTER_LINEAR)`` (OpenCV, not
vendored upstream; opencv-python 4.13 in the build container).  Its published algorithm for a CV_32FC1
image with two CV_32FC1 maps (modules/imgproc/src/imgwarp.cpp, RemapInvoker + remapBilinear) is restated
in ``remap_bilinear``: coordinates are rounded to 1/32 pixel (INTER_BITS = 5, round half to even), the four
weights come from a float table ((1-fy)*(1-fx), (1-fy)*fx, fy*(1-fx), fy*fx with fx, fy multiples of
1/32), the taps are accumulated left to right in float32 and out-of-image taps read the constant border 0.
tests/test_oracle.py pins it against cv2 itself and against the golden fixture made by upstream's own
functions (tests/golden/make_golden_dtu_filter.py).

Precision as upstream's NumPy promotion rules produce it: camera matrices, their inverses and the two
relative poses are float32; everything multiplied with the int64 pixel grid is float64.
"""
import math

import numpy as np

S, E = 1, 11                     # test_dtu_dypcd.py:33-34
DIST_BASE, DIFF_BASE = 1 / 2, 0.25   # :36-37
INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS


def remap_bilinear(src, map_x, map_y):
    """cv2.remap(src, map_x, map_y, INTER_LINEAR) for float32 src / maps, BORDER_CONSTANT 0."""
    h, w = src.shape
    sx = np.rint(map_x.astype(np.float32) * np.float32(INTER_TAB)).astype(np.int64)   # cvRound: half to even
    sy = np.rint(map_y.astype(np.float32) * np.float32(INTER_TAB)).astype(np.int64)
    bad = ~(np.isfinite(map_x) & np.isfinite(map_y))
    x0, y0 = sx >> INTER_BITS, sy >> INTER_BITS
    fx = ((sx & (INTER_TAB - 1)).astype(np.float32) / np.float32(INTER_TAB)).astype(np.float32)
    fy = ((sy & (INTER_TAB - 1)).astype(np.float32) / np.float32(INTER_TAB)).astype(np.float32)
    one = np.float32(1.0)
    w00 = ((one - fy) * (one - fx)).astype(np.float32)
    w01 = ((one - fy) * fx).astype(np.float32)
    w10 = (fy * (one - fx)).astype(np.float32)
    w11 = (fy * fx).astype(np.float32)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w) & ~bad
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0)).astype(np.float32)

    out = tap(y0, x0) * w00
    out = (out + tap(y0, x0 + 1) * w01).astype(np.float32)
    out = (out + tap(y0 + 1, x0) * w10).astype(np.float32)
    out = (out + tap(y0 + 1, x0 + 1) * w11).astype(np.float32)
    return out


def reproject_with_depth(depth_ref, K_ref, E_ref, depth_src, K_src, E_src):
    """test_dtu_dypcd.py:164-204"""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x_ref, y_ref = x_ref.reshape(-1), y_ref.reshape(-1)
    xyz_ref = np.matmul(np.linalg.inv(K_ref), np.vstack((x_ref, y_ref, np.ones_like(x_ref))) * depth_ref.reshape(-1))
    xyz_src = np.matmul(np.matmul(E_src, np.linalg.inv(E_ref)), np.vstack((xyz_ref, np.ones_like(x_ref))))[:3]
    K_xyz_src = np.matmul(K_src, xyz_src)
    xy_src = K_xyz_src[:2] / K_xyz_src[2:3]
    x_src = xy_src[0].reshape(height, width).astype(np.float32)
    y_src = xy_src[1].reshape(height, width).astype(np.float32)
    sampled = remap_bilinear(depth_src, x_src, y_src)
    xyz_src = np.matmul(np.linalg.inv(K_src), np.vstack((xy_src, np.ones_like(x_ref))) * sampled.reshape(-1))
    xyz_rep = np.matmul(np.matmul(E_ref, np.linalg.inv(E_src)), np.vstack((xyz_src, np.ones_like(x_ref))))[:3]
    depth_rep = xyz_rep[2].reshape(height, width).astype(np.float32)
    K_xyz_rep = np.matmul(K_ref, xyz_rep)
    K_xyz_rep[2:3][K_xyz_rep[2:3] == 0] += 0.00001
    xy_rep = K_xyz_rep[:2] / K_xyz_rep[2:3]
    return (depth_rep, xy_rep[0].reshape(height, width).astype(np.float32), xy_rep[1].reshape(height, width).astype(np.float32),
            x_src, y_src)


def check_geometric_consistency(depth_ref, K_ref, E_ref, depth_src, K_src, E_src):
    """test_dtu_dypcd.py:207-233 -> (masks [E-S], mask, depth_reprojected (zeroed outside mask), x_src, y_src)"""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth(depth_ref, K_ref, E_ref, depth_src, K_src, E_src)
    dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
    depth_diff = np.abs(depth_rep - depth_ref)
    masks = [np.logical_and(dist < i * DIST_BASE, depth_diff < math.log(max(i, 1.05), 10) * DIFF_BASE) for i in range(S, E)]
    mask = masks[-1]
    depth_rep[~mask] = 0
    return masks, mask, depth_rep, x_src, y_src


def filter_view(ref_depth, confidence, srcs_depth, K_ref, E_ref, Ks_src, Es_src, conf_thres=0.5):
    """test_dtu_dypcd.py:261-309, 320-337 for one reference view; confidence already at the depth map's size.
    -> dict(final, geo, depth_avg (float64), points (3,h,w) float64, masks (v, E-S, h, w))"""
    v = srcs_depth.shape[0]
    photo = confidence > conf_thres
    geo_sum = 0
    sums = [0] * (E - S)
    reps, all_masks = [], []
    for i in range(v):
        masks, m, rep, _, _ = check_geometric_consistency(ref_depth, K_ref, E_ref, srcs_depth[i], Ks_src[i], Es_src[i])
        geo_sum = geo_sum + m.astype(np.int32)
        for k in range(E - S):
            sums[k] = sums[k] + masks[k].astype(np.int32)
        reps.append(rep)
        all_masks.append(np.stack(masks))
    avg = (sum(reps) + ref_depth) / (geo_sum + 1)
    avg[confidence > 0.75] = ref_depth[confidence > 0.75]
    geo = geo_sum >= E
    for k in range(E - S):
        geo = np.logical_or(geo, sums[k] >= (S + k))
    h, w = ref_depth.shape
    x, y = np.meshgrid(np.arange(0, w), np.arange(0, h))
    x, y = x.reshape(-1), y.reshape(-1)
    xyz_ref = np.matmul(np.linalg.inv(K_ref), np.vstack((x, y, np.ones_like(x))) * avg.reshape(-1))
    xyz_world = np.matmul(np.linalg.inv(E_ref), np.vstack((xyz_ref, np.ones_like(x))))[:3]
    return {"final": np.logical_and(photo, geo), "geo": geo, "depth_avg": avg, "points": xyz_world.reshape(3, h, w),
            "masks": np.stack(all_masks), "reproj_depth": np.stack(reps)}
