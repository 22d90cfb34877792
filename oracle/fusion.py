"""ORACLE (test infrastructure, not product code) -- restatement of the geometric
consistency filter: misc/fusion.py:8-47, 117-181 and the vote / average /
back-projection arithmetic of test_tank.py:470-515 (upstream tree).

Device-agnostic (upstream hard-codes ``.cuda()`` at misc/fusion.py:9-10).  Pinned
against upstream outputs by tests/golden/make_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import hotpath as _hp


def pixel_centres(h: int, w: int, device) -> torch.Tensor:
    """(h,w,3,1) homogeneous pixel centres (x+0.5, y+0.5, 1)  (fusion.py:8-13)."""
    x = (torch.arange(w, dtype=torch.float32, device=device) + 0.5).repeat(h, 1)
    y = (torch.arange(h, dtype=torch.float32, device=device) + 0.5).repeat(w, 1).t()
    return torch.stack([x, y, torch.ones_like(x)], dim=-1).unsqueeze(-1)


def img2cam(pix, depth, cam):
    """K^-1 * pix, renormalised by (z + 1e-9), scaled by depth; homogeneous (fusion.py:23-28)."""
    c = cam[:, 1:2, :3, :3].unsqueeze(1).inverse() @ pix
    c = c / (c[..., -1:, :] + 1e-9) * depth.permute(0, 2, 3, 1).unsqueeze(4)
    return torch.cat([c, torch.ones_like(c[..., -1:, :])], dim=-2)


def cam2world(pc, cam):
    """E^-1 * pc, divided by (w + 1e-9) (fusion.py:31-34)."""
    pw = cam[:, 0:1].unsqueeze(1).inverse() @ pc
    return pw / (pw[..., -1:, :] + 1e-9)


def world2cam(pw, cam):
    """E * pw, divided by (w + 1e-9) (fusion.py:37-40)."""
    pc = cam[:, 0:1].unsqueeze(1) @ pw
    return pc / (pc[..., -1:, :] + 1e-9)


def cam2img(pc, cam):
    """K * (pc.xyz / (w + 1e-9)), divided by (z + 1e-9) (fusion.py:43-47)."""
    c = pc[..., :3, :] / (pc[..., 3:4, :] + 1e-9)
    p = cam[:, 1:2, :3, :3].unsqueeze(1) @ c
    return p / (p[..., -1:, :] + 1e-9)


def reproject(ref_depth, srcs_depth, ref_cam, srcs_cam):
    """get_reproj_dynamic (fusion.py:117-154).

    ref_depth (n,1,h,w), srcs_depth (n,v,1,h,w), ref_cam (n,2,4,4), srcs_cam (n,v,2,4,4)
    -> reproj_xyd (n,v,3,h,w): the ref pixel re-projected through each source view.
    """
    n, v, _, h, w = srcs_depth.shape
    sd = srcs_depth.reshape(n * v, 1, h, w)
    sc = srcs_cam.reshape(n * v, 2, 4, 4)
    rc = ref_cam.unsqueeze(1).repeat(1, v, 1, 1, 1).reshape(n * v, 2, 4, 4)
    rd = ref_depth.unsqueeze(1).repeat(1, v, 1, 1, 1).reshape(n * v, 1, h, w)
    pix = pixel_centres(h, w, ref_depth.device).unsqueeze(0)

    in_src = cam2img(world2cam(cam2world(img2cam(pix, rd, rc), rc), sc), sc)
    uv = in_src[..., :2, 0]
    gx = uv[..., 0] / ((w - 1) / 2) - 1
    gy = uv[..., 1] / ((h - 1) / 2) - 1
    d_src = _hp._grid_sample(sd, torch.stack((gx, gy), dim=-1))
    uv1 = torch.cat([uv, torch.ones_like(uv[..., -1:])], dim=-1).unsqueeze(-1)
    back = world2cam(cam2world(img2cam(uv1, d_src, sc), sc), rc)
    depth_back = back[:, :, :, 2, 0].clone()
    xy_back = cam2img(back, rc)
    xyd = torch.cat([xy_back[..., :2, 0], depth_back.unsqueeze(-1)], dim=-1).permute(0, 3, 1, 2)
    return xyd.reshape(n, v, 3, h, w)


def consistency_masks(ref_depth, reproj_xyd, dist_base, rel_diff_base, thres_view, relative=False):
    """vis_filter_dynamic (fusion.py:157-181) -> masks (n,v,K,h,w) bool, K = v - thres_view + 1."""
    n, v, _, h, w = reproj_xyd.shape
    dev = reproj_xyd.device
    xy = pixel_centres(h, w, dev).permute(3, 2, 0, 1).unsqueeze(1)[:, :, :2]
    e_xy = (reproj_xyd[:, :, :2] - xy).norm(dim=2, keepdim=True)
    e_d = (ref_depth.unsqueeze(1) - reproj_xyd[:, :, 2:]).abs()
    if relative:
        e_d = e_d / ref_depth.unsqueeze(1)
    k = torch.arange(thres_view, v + 1).reshape(1, 1, -1, 1, 1).to(dev)
    return torch.min(e_xy < k / dist_base, e_d < k / rel_diff_base)


def fuse_view(ref_depth, ref_conf, srcs_depth, ref_cam, srcs_cam,
              dist_base, rel_diff_base, thres_view, prob_threshold, relative=False):
    """One reference view of dynamic_filter_depth (test_tank.py:470-515).

    ref_conf (n,Hc,Wc) is nearest-resized to the depth resolution (:473).
    Returns dict(final (n,1,h,w) bool, geo, prob, depth_avg (n,1,h,w), points (n,3,h,w),
                 reproj_xyd, masks).
    """
    n, v, _, h, w = srcs_depth.shape
    conf = F.interpolate(ref_conf.unsqueeze(1), size=[h, w], mode="nearest")
    prob_mask = conf > prob_threshold
    xyd = reproject(ref_depth, srcs_depth, ref_cam, srcs_cam)
    masks = consistency_masks(ref_depth, xyd, dist_base, rel_diff_base, thres_view, relative)
    last = masks[:, :, -1:]
    rd = xyd[:, :, -1].clone()
    rd[~last.squeeze(2)] = 0
    votes = masks.sum(dim=1)                    # (n,K,h,w)
    n_last = last.sum(dim=1)                    # (n,1,h,w)
    avg = (rd.sum(dim=1, keepdim=True) + ref_depth) / (n_last + 1)
    dy_range = v + 1
    geo = n_last >= dy_range
    for i in range(thres_view, dy_range):
        geo = torch.logical_or(geo, votes[:, i - thres_view:i - thres_view + 1] >= i)
    final = torch.min(prob_mask, geo)
    pix = pixel_centres(h, w, ref_depth.device).unsqueeze(0)
    pts = cam2world(img2cam(pix, avg, ref_cam), ref_cam)[..., :3, 0].permute(0, 3, 1, 2)
    return {"final": final, "geo": geo, "prob": prob_mask, "depth_avg": avg, "points": pts,
            "reproj_xyd": xyd, "masks": masks}
