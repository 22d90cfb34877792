"""Upstream-named call sites backed by the hot-path table, and ``patch(model)``.

Every function / class here has the name, argument order and return structure of the upstream
symbol it replaces (bdwsq1996/Effi-MVS-plus), so that the swap is a monkey-patch:

    import sys; sys.path.insert(0, "<upstream checkout>")
    import models                                   # upstream
    import effimvs_b200
    from effimvs_b200 import dropin
    model = sys.modules["models.Effi_MVS_plus"].Effi_MVS_plus(args).cuda().eval()
    dropin.patch(model)                              # hot path now runs on libeffimvs.so
    out = model(imgs, proj_matrices, depth_values)   # upstream's own forward, untouched

What is replaced:
    DepthNet.forward                       models/Effi_MVS_plus.py:14
    GetCost_initvolume.forward             models/Effi_MVS_plus.py:184
    GetCost.forward                        models/Effi_MVS_plus.py:257
    pro_bilinear_sampler                   models/Effi_MVS_plus.py:118   (module global)
    CostRegNet_2_sample_FPN3D_Fast.forward models/module.py:453
    cost_up_small.forward                  models/module.py:509
    depth_regression                       models/module.py:518          (module global)
    BasicUpdateBlock.forward               models/update.py:114          (CUDA table only; section 8(f) row 3)
    upsample_depth                         models/Effi_MVS_plus.py:167   (module global; CUDA table only)

The table defaults to ``CudaHotPath``; tests inject the oracle's table to check the adapters
against unpatched upstream on the CPU.
"""
from __future__ import annotations

import sys
import types

import torch


def _table(hotpath):
    if hotpath is not None:
        return hotpath
    from .hotpath import CudaHotPath
    return CudaHotPath()


class _LazyPro:
    """Stands in for upstream's permuted ``(B*H*W,1,1,D)`` copy of a volume (Effi_MVS_plus.py:517,
    527, 541-545): remembers the un-permuted (B,D,H,W) tensor so that the lookup kernels read it in
    place.  Supports the only things upstream does with the copy: ``.shape`` and being handed to
    ``pro_bilinear_sampler`` / ``GetCost``."""

    def __init__(self, volume):
        self.volume = volume
        B, D, H, W = volume.shape
        self.shape = torch.Size((B * H * W, 1, 1, D))


def _volume_of(pro, like):
    """(B*H*W,1,1,D) tensor or _LazyPro -> (B,D,H,W), `like` gives B,H,W."""
    if isinstance(pro, _LazyPro):
        return pro.volume
    B, _, H, W = like.shape
    D = pro.shape[-1]
    return pro.reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous()


def make_pro_bilinear_sampler(hotpath=None):
    hp = _table(hotpath)

    def pro_bilinear_sampler(pro, depth_sample, depth_min, depth_max):
        """Drop-in for models/Effi_MVS_plus.py:118."""
        return hp.volume_lookup(_volume_of(pro, depth_sample), depth_sample, depth_min, depth_max)
    return pro_bilinear_sampler


def homo_warping_new(src_fea, src_proj, ref_proj, depth_values):
    """Drop-in for models/module.py:303 (materialises the warped (B,C,D,H,W) volume; CUDA only).
    depth_values (B,D) or (B,D,H,W)."""
    from . import capi, ops
    B = src_fea.shape[0]
    rel = torch.matmul(src_proj, torch.linalg.inv_ex(ref_proj)[0])
    proj = torch.cat([rel[:, :3, :3].reshape(B, 9), rel[:, :3, 3]], dim=-1).contiguous()
    D = depth_values.shape[1]
    mode = capi.HYP_PLANES if depth_values.dim() == 2 else capi.HYP_TENSOR
    return ops.homo_warp(src_fea, proj, depth_values, mode, D)


def get_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, device=None, dtype=None, shape=None, max_depth=192.0, min_depth=0.0):
    """Drop-in for models/module.py:572 (both branches).  The 2-D branch is upstream's own three torch
    expressions; the per-pixel branch runs the library kernel."""
    if cur_depth.dim() == 2:
        lo, hi = cur_depth[:, 0], cur_depth[:, -1]
        step = (hi - lo) / (ndepth - 1)
        k = torch.arange(0, ndepth, device=cur_depth.device, dtype=cur_depth.dtype).reshape(1, -1)
        s = lo.unsqueeze(1) + k * step.unsqueeze(1)
        return s.unsqueeze(-1).unsqueeze(-1).repeat(1, 1, shape[1], shape[2])
    from . import ops
    B = cur_depth.shape[0]
    assert cur_depth.shape == torch.Size(shape), "cur_depth:{}, input shape:{}".format(cur_depth.shape, shape)
    iv = depth_inteval_pixel.reshape(-1).expand(B) if torch.is_tensor(depth_inteval_pixel) else \
        torch.full((B,), float(depth_inteval_pixel), device=cur_depth.device)
    return ops.depth_range_samples(cur_depth, iv.float().contiguous(), ndepth)


def make_depth_regression(hotpath=None):
    def depth_regression(p, depth_values):
        """Drop-in for models/module.py:518 (p is already a probability volume): kept as the one
        torch expression it is -- the fused softmax+regression+confidence kernel is used by the
        DepthNet drop-in, where the logits are available."""
        if depth_values.dim() <= 2:
            depth_values = depth_values.view(*depth_values.shape, 1, 1)
        return torch.sum(p * depth_values, 1)
    return depth_regression


def make_depthnet_forward(hotpath=None):
    hp = _table(hotpath)

    def forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization, pixel_wise_net, G=8):
        """Drop-in for DepthNet.forward (models/Effi_MVS_plus.py:14-89)."""
        assert len(features) == proj_matrices.shape[1], "Different number of images and projection matrices"
        assert depth_values.shape[1] == num_depth, "depth_values.shape[1]:{}  num_depth:{}".format(depth_values.shape[1], num_depth)
        if pixel_wise_net is None:
            raise NotImplementedError("DepthNet without pixel_wise_net is not on upstream's inference path")
        return hp.stage1(list(features), proj_matrices, depth_values, pixel_wise_net, cost_regularization, G)
    return forward


def make_getcost_initvolume_forward(hotpath=None):
    hp = _table(hotpath)

    def forward(self, depth_values, features, proj_matrices, depth_interval, depth_max, depth_min, view_weights,
                CostNum=4, Inverse=True, G=8, iter=1, inter_iter=[1, 1, 1, 1]):
        """Drop-in for GetCost_initvolume.forward (models/Effi_MVS_plus.py:184-251)."""
        if not Inverse:
            raise NotImplementedError("upstream always runs with Inverse=True (Effi_MVS_plus.py:317)")
        interval = depth_interval * inter_iter[iter]
        return hp.local_volume(depth_values, list(features), proj_matrices, interval, view_weights, CostNum, G)
    return forward


def make_getcost_forward(hotpath=None):
    hp = _table(hotpath)

    def forward(self, depth_values, pro, features, proj_matrices, depth_interval, depth_max, depth_min, view_weights,
                CostNum=4, Inverse=True, G=8, depth_max_cur_volume=0, depth_min_cur_volume=0, iter=1,
                inter_iter=[1, 1, 1, 1]):
        """Drop-in for GetCost.forward (models/Effi_MVS_plus.py:257-303): pro[-1] raw, pro[0] regularized."""
        if not Inverse:
            raise NotImplementedError("upstream always runs with Inverse=True (Effi_MVS_plus.py:317)")
        interval = depth_interval * inter_iter[iter]
        raw, reg = _volume_of(pro[-1], depth_values), _volume_of(pro[0], depth_values)
        return hp.dynamic_cost(depth_values, raw, reg, interval, depth_min_cur_volume, depth_max_cur_volume, CostNum)
    return forward


def make_costreg_forward(hotpath=None):
    hp = _table(hotpath)

    def forward(self, x):
        """Drop-in for CostRegNet_2_sample_FPN3D_Fast.forward (models/module.py:453-463).  Returns
        (prob_pre, None): the second output (`pro`) is never read by upstream's caller
        (Effi_MVS_plus.py:77)."""
        return hp.cost_regularization(self, x), None
    return forward


def make_cost_up_forward(hotpath=None):
    hp = _table(hotpath)

    def forward(self, x, IGEV_cost):
        """Drop-in for cost_up_small.forward (models/module.py:509-516).  Returns (conv2, None): the
        second output is discarded by both call sites (Effi_MVS_plus.py:521, 530)."""
        return hp.cross_scale(self, x, IGEV_cost), None
    return forward


def make_update_block_forward(hotpath=None):
    hp = _table(hotpath)

    def forward(self, net, depth_cost_func, inv_depth, context, seq_len=4, scale_inv_depth=None):
        """Drop-in for BasicUpdateBlock.forward (models/update.py:114-141), SURVEY section 8(f) row 3: the 2-D
        convolutions stay cuDNN, the elementwise chains between them run in the glue kernels.
        Returns (net, mask_list, inv_depth_list) like upstream: mask_list[-1] = 0.25 * mask(net)."""
        from .net import update_block_forward_fused
        kw = getattr(scale_inv_depth, "keywords", None) or {}
        if "min_depth" not in kw or "max_depth" not in kw:
            raise NotImplementedError("scale_inv_depth must be upstream's partial(disp_to_depth, min_depth=, max_depth=) (Effi_MVS_plus.py:423)")
        B = inv_depth.shape[0]
        lo, hi = (1 / kw["max_depth"]).reshape(-1).expand(B), (1 / kw["min_depth"]).reshape(-1).expand(B)
        net, inv_seq, _, _, _, mask_pre = update_block_forward_fused(
            self, hp, net, lambda depth, it: depth_cost_func(depth, iter=it), inv_depth, context, seq_len,
            lo.contiguous(), hi.contiguous())
        masks = list(inv_seq)
        if self.UpMask:
            masks[-1] = 0.25 * (mask_pre + self.mask[2].bias.reshape(1, -1, 1, 1))
        return net, masks, inv_seq
    return forward


def make_upsample_depth(hotpath=None):
    hp = _table(hotpath)

    def upsample_depth(depth, mask, ratio=8):
        """Drop-in for upsample_depth (models/Effi_MVS_plus.py:167-178): (N,1,H,W), (N,9*ratio^2,H,W) -> (N,ratio*H,ratio*W)."""
        one = torch.ones(depth.shape[0], device=depth.device, dtype=torch.float32)
        return hp.convex_upsample(mask, None, 1.0, depth, one, one, ratio)[0]
    return upsample_depth


def patch(model, hotpath=None, upstream_module: str = "models.Effi_MVS_plus"):
    """Swap the hot-path call sites of an upstream ``Effi_MVS_plus`` instance in place.

    Instance-level ``forward`` attributes are replaced (the classes stay untouched) and the module
    global ``pro_bilinear_sampler`` is rebound in ``sys.modules[upstream_module]`` -- upstream does
    ``from .module import *`` (Effi_MVS_plus.py:4), so that is the namespace its forward reads.
    Returns a callable that undoes the patch."""
    hp = _table(hotpath)
    undo = []

    def swap(obj, name, value):
        had = name in vars(obj)
        old = vars(obj).get(name)
        setattr(obj, name, value)
        undo.append((obj, name, had, old))

    swap(model.depthnet, "forward", types.MethodType(make_depthnet_forward(hp), model.depthnet))
    swap(model.GetCost_initvolume, "forward", types.MethodType(make_getcost_initvolume_forward(hp), model.GetCost_initvolume))
    swap(model.GetCost, "forward", types.MethodType(make_getcost_forward(hp), model.GetCost))
    swap(model.cost_regularization, "forward", types.MethodType(make_costreg_forward(hp), model.cost_regularization))
    for net in list(model.CSP_R) + list(model.CSP_C):
        swap(net, "forward", types.MethodType(make_cost_up_forward(hp), net))
    mod = sys.modules.get(upstream_module)
    if mod is not None:
        swap(mod, "pro_bilinear_sampler", make_pro_bilinear_sampler(hp))
    if getattr(hp, "fused_update", False):     # section 8(f) row 3 (CUDA table only)
        for blk in model.update_block:
            swap(blk, "forward", types.MethodType(make_update_block_forward(hp), blk))
        if mod is not None and all(int(r) == 2 for r in getattr(model, "feat_ratio", [2])):
            swap(mod, "upsample_depth", make_upsample_depth(hp))

    def restore():
        for obj, name, had, old in reversed(undo):
            if had:
                setattr(obj, name, old)
            else:
                delattr(obj, name)
    return restore
