"""Seeded synthetic inputs at DTU / Tanks&Temples shapes.

There are no datasets on the build or GPU box, so tests and bench.py use these:
images in [0,1], pinhole cameras on an arc that all look at the working volume
(depth 425..935 mm, DTU-like), and the sample-dict layout of upstream's loaders
(datasets/general_eval.py:197-228): ``proj_matrices["stageK"]`` is (B,V,2,4,4) with
``[:,:,0]`` the 4x4 extrinsic and ``[:,:,1,:3,:3]`` the intrinsics scaled to stage K.
"""
from __future__ import annotations

import math

import torch

DEPTH_NEAR, DEPTH_FAR, NUM_DEPTH = 425.0, 935.0, 384

SHAPES = {
    "plumbing": dict(width=640, height=512, views=5, ndepths="48,8,8"),
    "dtu": dict(width=1600, height=1184, views=5, ndepths="48,8,8"),
    "tanks": dict(width=1920, height=1056, views=7, ndepths="96,8,8"),
}


def _rot_y(theta: float) -> torch.Tensor:
    c, s = math.cos(theta), math.sin(theta)
    return torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float64)


def _rot_x(theta: float) -> torch.Tensor:
    c, s = math.cos(theta), math.sin(theta)
    return torch.tensor([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]], dtype=torch.float64)


def camera_ring(views: int, width: int, height: int, focus: float = 680.0):
    """Extrinsics (V,4,4) and full-resolution intrinsics (3,3), float64.

    View 0 is the identity; view v sits on an arc (alternating sides, +-0.06*ceil(v/2) rad
    about y, a slight tilt about x) and is aimed at the point (0,0,focus) with a small
    decentring so that the epipolar geometry is not degenerate.
    """
    f = 1.8075 * width
    K = torch.tensor([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    E = torch.eye(4, dtype=torch.float64).repeat(views, 1, 1)
    target = torch.tensor([0.0, 0.0, focus], dtype=torch.float64)
    for v in range(1, views):
        k = (v + 1) // 2
        sign = 1.0 if v % 2 else -1.0
        R = _rot_y(sign * 0.06 * k) @ _rot_x(0.015 * k * (1.0 if v % 4 < 2 else -1.0))
        t = target - R @ target + torch.tensor([4.0 * sign * k, 3.0 * k, 2.0 * k], dtype=torch.float64)
        E[v, :3, :3] = R
        E[v, :3, 3] = t
    return E, K


def camera_arc(views: int, width: int, height: int, step: float = 0.02, focus: float = 680.0):
    """Scene-sized camera set: `views` cameras on an arc about the y axis (`step` rad apart, centred on
    view views//2), all aimed at (0,0,focus) -- neighbouring views overlap like a DTU scan."""
    f = 1.8075 * width
    K = torch.tensor([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    E = torch.eye(4, dtype=torch.float64).repeat(views, 1, 1)
    target = torch.tensor([0.0, 0.0, focus], dtype=torch.float64)
    for v in range(views):
        a = step * (v - views // 2)
        R = _rot_y(a) @ _rot_x(0.01 * ((v % 3) - 1))
        E[v, :3, :3] = R
        E[v, :3, 3] = target - R @ target + torch.tensor([1.5 * (v % 2), 1.0 * (v % 3), 0.5 * (v % 5)], dtype=torch.float64)
    return E, K


def stage_cameras(E: torch.Tensor, K: torch.Tensor, batch: int = 1, dtype=torch.float32):
    """{"stage1".."stage4"}: (B,V,2,4,4), intrinsic rows 0-1 scaled by 1/8, 1/4, 1/2, 1."""
    V = E.shape[0]
    out = {}
    for i, s in enumerate((0.125, 0.25, 0.5, 1.0)):
        cam = torch.zeros(V, 2, 4, 4, dtype=torch.float64)
        cam[:, 0] = E
        Ks = K.clone()
        Ks[:2] *= s
        cam[:, 1, :3, :3] = Ks
        out["stage{}".format(i + 1)] = cam.to(dtype).unsqueeze(0).repeat(batch, 1, 1, 1, 1).contiguous()
    return out


def make_sample(shape: str = "dtu", seed: int = 0, batch: int = 1, device="cpu", width=None, height=None, views=None):
    """imgs (B,V,3,H,W), proj_matrices, depth_values (B,384) inverse depth ascending."""
    cfg = dict(SHAPES[shape])
    if width:
        cfg["width"] = width
    if height:
        cfg["height"] = height
    if views:
        cfg["views"] = views
    g = torch.Generator().manual_seed(seed)
    W, H, V = cfg["width"], cfg["height"], cfg["views"]
    # smooth-ish random images: low-res noise upsampled plus fine noise, in [0,1]
    coarse = torch.rand(batch * V, 3, H // 16, W // 16, generator=g)
    imgs = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    imgs = (0.7 * imgs + 0.3 * torch.rand(batch * V, 3, H, W, generator=g)).reshape(batch, V, 3, H, W)
    E, K = camera_ring(V, W, H)
    proj = stage_cameras(E, K, batch)
    depth_values = torch.linspace(1.0 / DEPTH_FAR, 1.0 / DEPTH_NEAR, NUM_DEPTH, dtype=torch.float32).unsqueeze(0).repeat(batch, 1)
    return {"imgs": imgs.to(device), "proj_matrices": {k: v.to(device) for k, v in proj.items()},
            "depth_values": depth_values.to(device), "ndepths": cfg["ndepths"]}


def microbench_inputs(C: int, D: int, H: int, W: int, views: int = 5, seed: int = 0, device="cpu"):
    """Config-2 microbench tensors: V feature maps randn (1,C,H,W), hypotheses
    1/linspace(1/935,1/425,D) broadcast to (1,D,H,W), per-view weights rand (1,V-1,H,W),
    cameras at the scale where the full-res image is 8x the feature map."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(1, C, H, W, generator=g).to(device) for _ in range(views)]
    hyp = (1.0 / torch.linspace(1.0 / DEPTH_FAR, 1.0 / DEPTH_NEAR, D)).reshape(1, D, 1, 1).repeat(1, 1, H, W).to(device)
    wts = torch.rand(1, views - 1, H, W, generator=g).to(device)
    E, K = camera_ring(views, W, H)
    cams = stage_cameras(E, K, 1)["stage4"].to(device)
    return feats, cams, hyp, wts


def render_plane_scene(E: torch.Tensor, K: torch.Tensor, width: int, height: int, noise: float = 0.15, seed: int = 0):
    """Depth maps (V,H,W) of a smooth synthetic surface z = 680 + 60 sin(x/90) + 40 cos(y/70)
    seen from every camera (fixed-point ray marching), plus N(0, noise) mm; used by the
    fusion tests and bench so that most pixels are geometrically consistent."""
    V = E.shape[0]
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float64) + 0.5,
                            torch.arange(width, dtype=torch.float64) + 0.5, indexing="ij")
    rays = torch.linalg.inv(K) @ torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, dtype=torch.float64)])
    out = torch.zeros(V, height, width, dtype=torch.float32)
    for v in range(V):
        Rinv = E[v, :3, :3].T
        c = -Rinv @ E[v, :3, 3]
        dirs = Rinv @ rays                       # world-space ray per pixel, cam z == 1
        z = torch.full((height * width,), 680.0, dtype=torch.float64)
        for _ in range(30):
            p = c.reshape(3, 1) + dirs * z
            surf = 680.0 + 60.0 * torch.sin(p[0] / 90.0) + 40.0 * torch.cos(p[1] / 70.0)
            z = z + (surf - p[2]) / dirs[2].clamp(min=0.2)
        out[v] = (z.reshape(height, width) + noise * torch.randn(height, width, generator=g, dtype=torch.float64)).float()
    return out
