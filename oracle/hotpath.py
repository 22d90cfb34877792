"""ORACLE (test infrastructure, not product code) -- CPU/torch restatement of the
Effi-MVS+ cost-volume hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(``effi-mvs-plus_b200``) never does; it fails loudly when its CUDA library is
missing.

Every function restates one row of SURVEY.md section 8(a) and cites the reference
lines it follows (paths relative to the upstream tree, bdwsq1996/Effi-MVS-plus).
The arithmetic that upstream delegates to ATen (``F.grid_sample``, ``conv3d``,
``conv_transpose3d``; torch 2.11.0, not vendored upstream) is restated explicitly
in ``grid_sample_zeros_ac``, ``conv3d_taps`` and ``deconv3d_taps`` and cross-checked
against the ATen calls in ``tests/test_oracle.py``; the default path (``ATEN=True``)
issues the same ATen calls as upstream so that the CPU baseline is timed on the
same library kernels upstream would run.

Parity pin: upstream ships no tests or golden vectors, so this
oracle is pinned against outputs of the upstream code itself, run in the build
container by ``tests/golden/make_golden.py`` (committed with the fixtures it
wrote).  ``tests/test_oracle.py`` replays those fixtures without upstream present.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# True: call ATen grid_sample / conv3d like upstream does.  False: use the explicit
# restatements below (slower; used by tests to prove the two agree).
ATEN = True


# ----------------------------------------------------------------------------
# third-party arithmetic restated (ATen grid_sampler_2d, conv3d, conv_transpose3d)
# ----------------------------------------------------------------------------
def grid_sample_zeros_ac(img: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """Bilinear ``grid_sample(padding_mode='zeros', align_corners=True)`` restated.

    ATen's published algorithm (GridSampler.cpp/.cu, torch 2.11): un-normalise
    ``i = ((g + 1) / 2) * (size - 1)``; corners nw=(floor ix, floor iy), ne, sw, se;
    weights nw=(ix_se-ix)(iy_se-iy), ne=(ix-ix_sw)(iy_sw-iy), sw=(ix_ne-ix)(iy-iy_ne),
    se=(ix-ix_nw)(iy-iy_nw); every corner outside the image contributes zero.
    Upstream call sites: models/module.py:340, models/Effi_MVS_plus.py:112,
    misc/fusion.py:139.   img (N,C,H,W), grid (N,Ho,Wo,2) -> (N,C,Ho,Wo)
    """
    N, C, H, W = img.shape
    gx, gy = grid[..., 0], grid[..., 1]
    ix = ((gx + 1) / 2) * (W - 1)
    iy = ((gy + 1) / 2) * (H - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = img.reshape(N, C, H * W)
    out = torch.zeros((N, C) + tuple(gx.shape[1:]), dtype=img.dtype, device=img.device)
    for xc, yc, wt in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        ok = (xc >= 0) & (xc <= W - 1) & (yc >= 0) & (yc <= H - 1)  # NaN/inf -> False
        xi = torch.where(ok, xc, torch.zeros_like(xc)).long()
        yi = torch.where(ok, yc, torch.zeros_like(yc)).long()
        idx = (yi * W + xi).reshape(N, 1, -1).expand(N, C, -1)
        val = torch.gather(flat, 2, idx).reshape(out.shape)
        wt = torch.where(ok, wt, torch.zeros_like(wt))
        out = out + val * wt.unsqueeze(1)
    return out


def _grid_sample(img, grid):
    if ATEN:
        return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return grid_sample_zeros_ac(img, grid)


def conv3d_taps(x: torch.Tensor, w: torch.Tensor, stride=(1, 1, 1)) -> torch.Tensor:
    """3x3x3 cross-correlation, zero padding 1, restated tap by tap.

    out[o, z, y, x] = sum_{i,a,b,c} w[o,i,a,b,c] * xpad[i, z*sd + a, y*sh + b, x*sw + c]
    (cuDNN/oneDNN convolution as used by nn.Conv3d, models/module.py:146).
    """
    sd, sh, sw = stride
    B, Ci, D, H, W = x.shape
    Do, Ho, Wo = (D + 2 - 3) // sd + 1, (H + 2 - 3) // sh + 1, (W + 2 - 3) // sw + 1
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    out = torch.zeros(B, w.shape[0], Do, Ho, Wo, dtype=x.dtype, device=x.device)
    for a in range(3):
        for b in range(3):
            for c in range(3):
                sl = xp[:, :, a:a + sd * (Do - 1) + 1:sd, b:b + sh * (Ho - 1) + 1:sh, c:c + sw * (Wo - 1) + 1:sw]
                out += torch.einsum("bidhw,oi->bodhw", sl, w[:, :, a, b, c])
    return out


def deconv3d_taps(x: torch.Tensor, w: torch.Tensor, stride=(2, 2, 2), output_padding=(1, 1, 1)) -> torch.Tensor:
    """3x3x3 transposed convolution, padding 1, restated as a scatter of taps.

    out[o, z*sd + a - 1, y*sh + b - 1, x*sw + c - 1] += w[i,o,a,b,c] * in[i,z,y,x]
    with output size (n-1)*s - 2 + 3 + output_padding (nn.ConvTranspose3d,
    models/module.py:189; weight layout (Cin, Cout, 3,3,3)).
    """
    sd, sh, sw = stride
    B, Ci, D, H, W = x.shape
    Do = (D - 1) * sd + 1 + output_padding[0]
    Ho = (H - 1) * sh + 1 + output_padding[1]
    Wo = (W - 1) * sw + 1 + output_padding[2]
    # canvas with a one-voxel apron so that index -1 and the far edge are writable
    canvas = torch.zeros(B, w.shape[1], (D - 1) * sd + 3, (H - 1) * sh + 3, (W - 1) * sw + 3,
                         dtype=x.dtype, device=x.device)
    for a in range(3):
        for b in range(3):
            for c in range(3):
                contrib = torch.einsum("bidhw,io->bodhw", x, w[:, :, a, b, c])
                canvas[:, :, a:a + sd * (D - 1) + 1:sd, b:b + sh * (H - 1) + 1:sh, c:c + sw * (W - 1) + 1:sw] += contrib
    return canvas[:, :, 1:1 + Do, 1:1 + Ho, 1:1 + Wo].contiguous()


def _conv3d(x, w, stride):
    if ATEN:
        return F.conv3d(x, w, None, stride=stride, padding=1)
    return conv3d_taps(x, w, stride)


def _deconv3d(x, w, stride, output_padding):
    if ATEN:
        return F.conv_transpose3d(x, w, None, stride=stride, padding=1, output_padding=output_padding)
    return deconv3d_taps(x, w, stride, output_padding)


# ----------------------------------------------------------------------------
# a1  projection compose             models/Effi_MVS_plus.py:34-37 (dup :217-220)
# ----------------------------------------------------------------------------
def compose_projection(cam: torch.Tensor) -> torch.Tensor:
    """cam (B,2,4,4): [:,0]=extrinsic E, [:,1,:3,:3]=intrinsic K -> P (B,4,4), P[:3,:4]=K@E[:3,:4]."""
    P = cam[:, 0].clone()
    P[:, :3, :4] = torch.matmul(cam[:, 1, :3, :3], cam[:, 0, :3, :4])
    return P


def relative_projection(src_P: torch.Tensor, ref_P: torch.Tensor) -> torch.Tensor:
    """proj = P_src @ inverse(P_ref)  (models/module.py:314)."""
    return torch.matmul(src_P, torch.inverse(ref_P))


# ----------------------------------------------------------------------------
# a2  homography warp                models/module.py:303-344
# ----------------------------------------------------------------------------
def warp_coordinates(proj: torch.Tensor, depth: torch.Tensor):
    """Source-image pixel coordinates of every (hypothesis, ref pixel).

    proj (B,4,4), depth (B,D,H,W) -> u, v (B,D,H*W) in source pixels (integer ref
    pixel centres, models/module.py:318-330; z==0 -> z+1e-8 at :328).
    """
    B, D, H, W = depth.shape
    dev = depth.device
    rot, trans = proj[:, :3, :3], proj[:, :3, 3:4]
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                            torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack((xx.reshape(-1), yy.reshape(-1), torch.ones(H * W, device=dev)))  # (3,HW)
    ray = torch.matmul(rot, pix.unsqueeze(0).expand(B, 3, H * W))                        # (B,3,HW)
    p = ray.unsqueeze(2) * depth.reshape(B, 1, D, H * W) + trans.reshape(B, 3, 1, 1)     # (B,3,D,HW)
    z = p[:, 2]
    z = torch.where(z == 0, z + 1e-8, z)
    return p[:, 0] / z, p[:, 1] / z


def homo_warp(src_fea: torch.Tensor, src_P: torch.Tensor, ref_P: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """Restates homo_warping_new (models/module.py:303-344) -> (B,C,D,H,W).

    The normalise (:336-337) / ATen un-normalise round trip is kept so that the
    fp32 coordinates round exactly like upstream's.
    """
    B, C, H, W = src_fea.shape
    D = depth.shape[1]
    u, v = warp_coordinates(relative_projection(src_P, ref_P), depth)
    gx = u / ((W - 1) / 2) - 1
    gy = v / ((H - 1) / 2) - 1
    grid = torch.stack((gx, gy), dim=3).reshape(B, D * H, W, 2)
    return _grid_sample(src_fea, grid).reshape(B, C, D, H, W)


# ----------------------------------------------------------------------------
# a3  group-wise correlation         models/Effi_MVS_plus.py:39-40, :222-224
# ----------------------------------------------------------------------------
def group_correlation(warped: torch.Tensor, ref_fea: torch.Tensor, G: int) -> torch.Tensor:
    """(B,C,D,H,W) x (B,C,H,W) -> (B,G,D,H,W): mean over the C/G channels of a group."""
    B, C, D, H, W = warped.shape
    return (warped.reshape(B, G, C // G, D, H, W) * ref_fea.reshape(B, G, C // G, 1, H, W)).mean(2)


def view_similarity(ref_fea, src_fea, ref_cam, src_cam, depth, G):
    """One source view: a1 + a2 + a3 -> (B,G,D,H,W)."""
    warped = homo_warp(src_fea, compose_projection(src_cam), compose_projection(ref_cam), depth)
    return group_correlation(warped, ref_fea, G)


# ----------------------------------------------------------------------------
# a4  stage-1 view weighting         models/Effi_MVS_plus.py:41-53, 65-67
# ----------------------------------------------------------------------------
def similarity_entropy(sim: torch.Tensor) -> torch.Tensor:
    """sim (B,1,D,H,W) -> entropy (B,1,H,W) of softmax over D, log(p + 1e-7) (:43-44)."""
    p = F.softmax(sim.squeeze(1), dim=1)
    return (-p * torch.log(p + 1e-7)).sum(dim=1, keepdim=True)


def weighted_aggregate(sims, weights):
    """sum_v w_v * sim_v / (sum_v w_v + 1e-6); sims[v] (B,G,D,H,W), weights[v] (B,1,H,W) (:52-53, :67)."""
    num = 0
    den = 0
    for s, w in zip(sims, weights):
        num = num + s * w.unsqueeze(1)
        den = den + w.unsqueeze(1)
    return num / (den + 1e-6)


def stage1_volume(features, cams, depth, pixel_wise_net, G=1):
    """Stage-1 plane-sweep volume with learned view weights (DepthNet.forward :14-71).

    features: list of V (B,C,H,W); cams (B,V,2,4,4); depth (B,D,H,W).
    Returns similarity (B,G,D,H,W) and view_weights (B,V-1,H,W).
    """
    ref_cam = cams[:, 0]
    sims, wts = [], []
    for v in range(1, len(features)):
        s = view_similarity(features[0], features[v], ref_cam, cams[:, v], depth, G)
        wts.append(pixel_wise_net(similarity_entropy(s)))
        sims.append(s)
    return weighted_aggregate(sims, wts), torch.cat(wts, dim=1)


# ----------------------------------------------------------------------------
# a5  local hypotheses + local volume  models/module.py:554-591, Effi_MVS_plus.py:184-251
# ----------------------------------------------------------------------------
def local_inverse_depth_samples(inv_depth: torch.Tensor, ndepth: int, interval) -> torch.Tensor:
    """get_cur_depth_range_samples (models/module.py:554-570) on inverse depth.

    inv_depth (B,H,W); interval scalar or broadcastable to (B,H,W) -> (B,ndepth,H,W).
    """
    half = ndepth // 2 * interval
    lo = (inv_depth - half).clamp(min=1e-4)
    hi = (inv_depth + half).clamp(min=1e-4, max=1e4)
    step = (hi - lo) / (ndepth - 1)
    k = torch.arange(0, ndepth, device=inv_depth.device, dtype=inv_depth.dtype).reshape(1, -1, 1, 1)
    return (lo.unsqueeze(1) + k * step.unsqueeze(1)).clamp(min=1e-5)


def uniform_inverse_depth_samples(depth_values: torch.Tensor, ndepth: int, H: int, W: int) -> torch.Tensor:
    """get_depth_range_samples, 2-D branch (models/module.py:577-585): (B,Dv) -> (B,ndepth,H,W)."""
    lo, hi = depth_values[:, 0], depth_values[:, -1]
    step = (hi - lo) / (ndepth - 1)
    k = torch.arange(0, ndepth, device=depth_values.device, dtype=depth_values.dtype).reshape(1, -1)
    s = lo.unsqueeze(1) + k * step.unsqueeze(1)
    return s.reshape(s.shape[0], ndepth, 1, 1).repeat(1, 1, H, W)


def local_volume(cur_depth, features, cams, interval, view_weights, ndepth, G=1):
    """GetCost_initvolume.forward (Effi_MVS_plus.py:184-251), Inverse=True.

    cur_depth (B,1,H,W); interval (B,1,1,1) or float; view_weights (B,V-1,H,W) or None.
    Returns similarity (B,G*ndepth,H,W) and the depth hypotheses (B,ndepth,H,W).
    """
    inv = 1.0 / cur_depth
    iv = interval.squeeze(1) if torch.is_tensor(interval) else interval
    hyp = 1.0 / local_inverse_depth_samples(inv.squeeze(1), ndepth, iv)
    ref_cam = cams[:, 0]
    sims = [view_similarity(features[0], features[v], ref_cam, cams[:, v], hyp, G) for v in range(1, len(features))]
    if view_weights is not None:
        sim = weighted_aggregate(sims, [view_weights[:, i:i + 1] for i in range(len(sims))])
    else:
        sim = sum(sims) / len(sims)
    B, _, _, H, W = sim.shape
    return sim.reshape(B, G * ndepth, H, W), hyp


# ----------------------------------------------------------------------------
# a6  volume lookup                   models/Effi_MVS_plus.py:102-164
# ----------------------------------------------------------------------------
def volume_lookup(volume: torch.Tensor, depth_sample: torch.Tensor, depth_min, depth_max) -> torch.Tensor:
    """pro_bilinear_sampler on a (B,D,H,W) volume (upstream takes the permuted (B*H*W,1,1,D) copy).

    t = (1/depth - 1/dmax) / ((1/dmin - 1/dmax) + 1e-10) * (D-1)   (:151-164, :123)
    then a 1-D linear interpolation along D with zero contribution from taps
    outside [0, D-1] (grid_sample on a 1 x D image, :107-112).
    depth_sample (B,d,H,W); depth_min/max broadcastable to (B,1,H,W) -> (B,d,H,W)
    """
    B, D, H, W = volume.shape
    d = depth_sample.shape[1]
    disp = (1 / depth_sample - 1 / depth_max) / ((1 / depth_min - 1 / depth_max) + 1e-10)
    t = disp * (D - 1)
    pro = volume.permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, D)
    x = t.permute(0, 2, 3, 1).reshape(B * H * W, 1, d, 1)
    gx = 2 * x / (D - 1) - 1
    grid = torch.cat([gx, torch.zeros_like(gx)], dim=-1)
    out = _grid_sample(pro, grid)                     # (BHW,1,1,d)
    return out.reshape(B, H, W, d).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------
# a7  dynamic cost lookup             models/Effi_MVS_plus.py:257-303
# ----------------------------------------------------------------------------
def dynamic_cost(cur_depth, raw_volume, reg_volume, interval, vol_depth_min, vol_depth_max, ndepth=3):
    """GetCost.forward: ndepth hypotheses around cur_depth, looked up in the raw
    (pro[-1]) and regularized (pro[0]) volumes -> (B, 2*ndepth, H, W)."""
    inv = 1.0 / cur_depth
    iv = interval.squeeze(1) if torch.is_tensor(interval) else interval
    hyp = 1.0 / local_inverse_depth_samples(inv.squeeze(1), ndepth, iv)
    a = volume_lookup(raw_volume, hyp, vol_depth_min, vol_depth_max)
    b = volume_lookup(reg_volume, hyp, vol_depth_min, vol_depth_max)
    return torch.cat([a, b], dim=1)


# ----------------------------------------------------------------------------
# a9 / a10  3-D regularization        models/module.py:435-463, 501-516
# ----------------------------------------------------------------------------
def fold_bn(conv_w: torch.Tensor, bn, transposed=False):
    """Eval-mode BatchNorm folded into the preceding bias-free conv (module.py:146-160, 189-203).

    Returns (w_folded, bias).  Deconv weights are (Cin,Cout,...): scale on dim 1.
    """
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    if transposed:
        return conv_w * scale.reshape(1, -1, 1, 1, 1), shift
    return conv_w * scale.reshape(-1, 1, 1, 1, 1), shift


def _cbr(x, blk, stride=(1, 1, 1)):
    """conv3d + eval BN + ReLU, computed the upstream way (conv, then affine) in fp32."""
    y = _conv3d(x, blk.conv.weight, stride)
    bn = blk.bn
    y = (y - bn.running_mean.reshape(1, -1, 1, 1, 1)) / torch.sqrt(bn.running_var.reshape(1, -1, 1, 1, 1) + bn.eps)
    y = y * bn.weight.reshape(1, -1, 1, 1, 1) + bn.bias.reshape(1, -1, 1, 1, 1)
    return torch.relu(y)


def _dbr(x, blk, stride, output_padding):
    y = _deconv3d(x, blk.conv.weight, stride, output_padding)
    bn = blk.bn
    y = (y - bn.running_mean.reshape(1, -1, 1, 1, 1)) / torch.sqrt(bn.running_var.reshape(1, -1, 1, 1, 1) + bn.eps)
    y = y * bn.weight.reshape(1, -1, 1, 1, 1) + bn.bias.reshape(1, -1, 1, 1, 1)
    return torch.relu(y)


def cost_regularization(net, x: torch.Tensor):
    """CostRegNet_2_sample_FPN3D_Fast.forward (module.py:453-463).  net holds conv0..conv7, prob."""
    c1 = _cbr(_cbr(x, net.conv0), net.conv1)
    c3 = _cbr(_cbr(c1, net.conv2, (2, 2, 2)), net.conv3)
    y = _cbr(_cbr(c3, net.conv4, (2, 2, 2)), net.conv5)
    y = c3 + _dbr(y, net.conv6, (2, 2, 2), (1, 1, 1))
    pro = c1 + _dbr(y, net.conv7, (2, 2, 2), (1, 1, 1))
    return _conv3d(pro, net.prob.weight, (1, 1, 1)), pro


def cross_scale_net(net, x: torch.Tensor, prev: torch.Tensor):
    """cost_up_small.forward (module.py:509-516): x (B,1,D,H,W) full res, prev (B,1,D,H/2,W/2)."""
    a = _cbr(x, net.conv0, (1, 2, 2))
    b = _cbr(prev, net.conv_cost)
    c1 = _cbr(torch.cat([a, b], dim=1), net.conv1)
    return _dbr(c1, net.conv2, (1, 2, 2), (0, 1, 1)), c1


# ----------------------------------------------------------------------------
# a11 / a12  regression + confidence  Effi_MVS_plus.py:78-88, module.py:518-524
# ----------------------------------------------------------------------------
def softmax_regress_confidence(prob_pre: torch.Tensor, hyp: torch.Tensor):
    """prob_pre, hyp (B,D,H,W) -> depth (B,H,W), confidence (B,H,W).

    confidence = sum of p over the 4 bins [idx-1, idx+2] around idx = trunc(E[d]) clamped
    to [0, D-1] (zero outside the volume).
    """
    B, D, H, W = prob_pre.shape
    p = F.softmax(prob_pre, dim=1)
    depth = torch.sum(p * hyp, 1)
    sum4 = 4 * F.avg_pool3d(F.pad(p.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
    idx = torch.sum(p * torch.arange(D, device=p.device, dtype=torch.float).reshape(1, D, 1, 1), 1).long()
    idx = idx.clamp(min=0, max=D - 1)
    conf = torch.gather(sum4, 1, idx.unsqueeze(1)).squeeze(1)
    return depth, conf


# ----------------------------------------------------------------------------
# hot-path table handed to the host model (effi-mvs-plus_b200/net.py) by tests/bench
# ----------------------------------------------------------------------------
class OracleHotPath:
    """Same method table as the product's CudaHotPath, computed with the restatements above."""

    name = "oracle"

    def stage1(self, features, cams, depth_hyp, pixel_wise_net, reg_net, G):
        sim, vw = stage1_volume(features, cams, depth_hyp, pixel_wise_net, G)
        prob_pre, _ = cost_regularization(reg_net, sim)
        prob_pre = prob_pre.squeeze(1)
        depth, conf = softmax_regress_confidence(prob_pre, depth_hyp)
        return {"depth": depth, "photometric_confidence": conf, "view_weights": vw,
                "reg_volume": prob_pre, "volume": sim}

    def local_volume(self, cur_depth, features, cams, interval, view_weights, ndepth, G):
        return local_volume(cur_depth, features, cams, interval, view_weights, ndepth, G)

    def volume_lookup(self, volume, depth_sample, depth_min, depth_max):
        return volume_lookup(volume, depth_sample, depth_min, depth_max)

    def cost_regularization(self, net, x):
        return cost_regularization(net, x)[0]

    def cross_scale(self, net, cur_volume, prev_resampled):
        out, _ = cross_scale_net(net, cur_volume, prev_resampled)
        return out

    def dynamic_cost(self, cur_depth, raw_volume, reg_volume, interval, vmin, vmax, ndepth):
        return dynamic_cost(cur_depth, raw_volume, reg_volume, interval, vmin, vmax, ndepth)
