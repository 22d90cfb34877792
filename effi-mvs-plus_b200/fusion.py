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
    inv = inverse_cameras(ref_cam, srcs_cam) if torch_inverse else None
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
    inv = inverse_cameras(ref_cam, srcs_cam) if torch_inverse else None
    final, avg, pts, masks = ops.fusion_filter(ref_depth, srcs_depth, ref_conf, ref_cam, srcs_cam, inv, float(dist_base),
                                               float(rel_diff_base), int(thres_view), float(prob_threshold), bool(relative),
                                               bool(want_masks))
    out = {"final": final.bool(), "depth_avg": avg, "points": pts}
    if want_masks:
        out["masks"] = masks.bool()
    return out
