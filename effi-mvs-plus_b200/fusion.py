"""Geometric-consistency fusion on the CUDA path, with upstream's function names where a call
site exists (misc/fusion.py:117-181; driver arithmetic of test_tank.py:470-515)."""
from __future__ import annotations

import torch

from . import ops


def inverse_cameras(ref_cam: torch.Tensor, srcs_cam: torch.Tensor) -> torch.Tensor:
    """(n,2,4,4), (n,v,2,4,4) -> (n,1+v,2,4,4) holding inverse(E) and inverse(K) (padded to 4x4),
    taken with torch's LU like upstream's ``.inverse()`` calls (fusion.py:24,32) but without the
    host synchronisation of the error check."""
    cams = torch.cat([ref_cam.unsqueeze(1), srcs_cam], dim=1)
    out = torch.zeros_like(cams)
    out[:, :, 0] = torch.linalg.inv_ex(cams[:, :, 0])[0]
    out[:, :, 1, :3, :3] = torch.linalg.inv_ex(cams[:, :, 1, :3, :3])[0]
    return out.contiguous()


def get_reproj_dynamic(ref_depth, srcs_depth, ref_cam, srcs_cam, torch_inverse: bool = True):
    """Drop-in for misc/fusion.py:117 ``get_reproj_dynamic``.  Returns (reproj_xyd (n,v,3,h,w), None, None):
    the two camera-space tensors upstream also returns are never read by its own caller's
    arithmetic (vis_filter_dynamic only reshapes them, fusion.py:161-162) and are not materialised."""
    inv = inverse_cameras(ref_cam, srcs_cam) if torch_inverse else ops.fusion_invert_cameras(ref_cam, srcs_cam)
    return ops.fusion_reproject(ref_depth, srcs_depth, ref_cam, srcs_cam, inv), None, None


def vis_filter_dynamic(ref_depth, reproj_xyd, ref_idx_world=None, src2ref_idx_cam=None, dist_base=4, rel_diff_base=1300,
                       thres_view=2, relative=False):
    """Drop-in for misc/fusion.py:157 ``vis_filter_dynamic``: (masks (n,v,K,h,w) bool, mask (n,v,1,h,w) bool).
    The two camera-space arguments are accepted and ignored, as upstream's own arithmetic ignores them."""
    masks = ops.fusion_masks(ref_depth, reproj_xyd, float(dist_base), float(rel_diff_base), int(thres_view), bool(relative)).bool()
    return masks, masks[:, :, -1:]


def filter_view(ref_depth, ref_conf, srcs_depth, ref_cam, srcs_cam, dist_base, rel_diff_base, thres_view,
                prob_threshold, relative: bool = False, want_masks: bool = False, torch_inverse: bool = True):
    """One reference view of ``dynamic_filter_depth`` (test_tank.py:470-515) in a single kernel.

    ref_depth (n,1,h,w), ref_conf (n,hc,wc), srcs_depth (n,v,1,h,w), cams as upstream.
    -> dict(final (n,1,h,w) bool, depth_avg (n,1,h,w), points (n,3,h,w)[, masks (n,v,K,h,w) bool])."""
    inv = inverse_cameras(ref_cam, srcs_cam) if torch_inverse else ops.fusion_invert_cameras(ref_cam, srcs_cam)
    final, avg, pts, masks = ops.fusion_filter(ref_depth, srcs_depth, ref_conf, ref_cam, srcs_cam, inv, float(dist_base),
                                               float(rel_diff_base), int(thres_view), float(prob_threshold), bool(relative),
                                               bool(want_masks))
    out = {"final": final.bool(), "depth_avg": avg, "points": pts}
    if want_masks:
        out["masks"] = masks.bool()
    return out


# ---- the DTU pipeline's filter (test_dtu_dypcd.py:164-350) ---------------------------
DTU_FIRST_RUNG, DTU_END_RUNG = 1, 11          # s, e (test_dtu_dypcd.py:33-34)
DTU_DIST_BASE, DTU_DIFF_BASE = 1 / 2, 0.25    # :36-37


def dtu_camera_pack(K_ref, E_ref, Ks_src, Es_src):
    """The float32 matrices of reproject_with_depth / filter_depth, formed on the host with the same NumPy
    calls upstream uses (np.linalg.inv and np.matmul on float32 arrays, test_dtu_dypcd.py:172-176, 192-196,
    330-333): (34 + 50 v,) float32 in the layout effimvs_dtu_filter_f32 documents."""
    import numpy as np
    f32 = lambda a: np.asarray(a.cpu() if torch.is_tensor(a) else a, dtype=np.float32)   # noqa: E731
    K_ref, E_ref = f32(K_ref), f32(E_ref)
    parts = [np.linalg.inv(K_ref).reshape(-1), K_ref.reshape(-1), np.linalg.inv(E_ref).reshape(-1)]
    for K_src, E_src in zip(Ks_src, Es_src):
        K_src, E_src = f32(K_src), f32(E_src)
        parts += [np.matmul(E_src, np.linalg.inv(E_ref)).reshape(-1), K_src.reshape(-1), np.linalg.inv(K_src).reshape(-1),
                  np.matmul(E_ref, np.linalg.inv(E_src)).reshape(-1)]
    return torch.from_numpy(np.concatenate(parts).astype(np.float32))


def dtu_filter_view(ref_depth, confidence, srcs_depth, K_ref, E_ref, Ks_src, Es_src, conf_thres: float = 0.5,
                    want_masks: bool = False, first_rung: int = DTU_FIRST_RUNG, end_rung: int = DTU_END_RUNG,
                    dist_base: float = DTU_DIST_BASE, diff_base: float = DTU_DIFF_BASE):
    """One reference view of upstream's DTU ``filter_depth`` (test_dtu_dypcd.py:236-337) in a single kernel.

    ref_depth (h,w), srcs_depth (v,h,w) CUDA fp32; confidence (hc,wc) CUDA fp32, resized to (h,w) bilinearly with
    half-pixel centres like cv2.resize (:258) when the sizes differ; cameras as host arrays (3x3 / 4x4).
    -> dict(final, geo (h,w) bool, depth_avg (h,w), points (3,h,w)[, masks (v,K,h,w) bool, reproj_depth (v,h,w)])."""
    import math
    import numpy as np
    h, w = ref_depth.shape[-2:]
    if tuple(confidence.shape[-2:]) != (h, w):
        confidence = torch.nn.functional.interpolate(confidence.reshape(1, 1, *confidence.shape[-2:]), size=(h, w), mode="bilinear",
                                                     align_corners=False).reshape(h, w)
    rungs = range(first_rung, end_rung)
    thr_dist = [i * dist_base for i in rungs]
    thr_diff = [float(np.float32(math.log(max(i, 1.05), 10) * diff_base)) for i in rungs]
    mats = dtu_camera_pack(K_ref, E_ref, Ks_src, Es_src).to(ref_depth.device)
    final, geo, avg, pts, masks, rep = ops.dtu_filter(ref_depth.reshape(h, w), srcs_depth.reshape(-1, h, w), confidence.reshape(h, w), mats,
                                                      thr_dist, thr_diff, first_rung, end_rung, float(conf_thres), 0.75, bool(want_masks))
    out = {"final": final.bool(), "geo": geo.bool(), "depth_avg": avg, "points": pts}
    if want_masks:
        out["masks"], out["reproj_depth"] = masks.bool(), rep
    return out
