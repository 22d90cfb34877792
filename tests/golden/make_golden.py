"""Generates the golden fixtures in this directory by running the UPSTREAM code
(bdwsq1996/Effi-MVS-plus, mounted read-only at /root/reference in the build
container) on seeded inputs.  Upstream cannot travel to the GPU box, so the
outputs are committed as small .npz files and replayed by tests/test_oracle.py
and tests/test_gpu_parity.py against the oracle and the CUDA path.

    python tests/golden/make_golden.py        # needs /root/reference; CPU only

Upstream functions executed (unmodified): models/module.py homo_warping_new,
depth_regression, get_depth_range_samples, CostRegNet_2_sample_FPN3D_Fast,
cost_up_small; models/Effi_MVS_plus.py DepthNet, GetCost_initvolume, GetCost,
pro_bilinear_sampler, Effi_MVS_plus.forward; misc/fusion.py get_reproj_dynamic,
vis_filter_dynamic (+ helpers), and the vote/average/back-projection statements of
test_tank.py:473-515 replayed line by line on upstream's helpers.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import models  # noqa: E402  (upstream)

UE = sys.modules["models.Effi_MVS_plus"]
UM = sys.modules["models.module"]
torch.Tensor.cuda = lambda self, *a, **k: self      # misc/fusion.py:9-10 hard-codes .cuda()
import misc.fusion as ufusion  # noqa: E402  (upstream)

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import synthetic  # noqa: E402


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: tuple(v.shape) for k, v in out.items()})


def module_arrays(mod, prefix):
    return {prefix + k.replace(".", "__"): v for k, v in mod.state_dict().items() if "num_batches" not in k}


def randomize_bn(mod, g):
    for m in mod.modules():
        if isinstance(m, (torch.nn.BatchNorm3d, torch.nn.BatchNorm2d)):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
            m.running_mean.data = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)


@torch.no_grad()
def main():
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1234)

    # ---- warp + correlation + aggregation (a1-a3, a5), incl. out-of-frustum and z<=0 samples
    C, D, H, W, V, G = 8, 6, 20, 28, 4, 2
    feats, cams, hyp, wts = synthetic.microbench_inputs(C, D, H, W, views=V, seed=3)
    hyp = hyp.clone()
    hyp[:, 0] = 40.0                                   # far too close: mostly out of frustum
    hyp[:, 1, :4] = -500.0                             # behind the camera (upstream has no z<0 test)
    hyp[:, 2, 5, 5] = 0.0                              # depth 0 -> z == t_z
    P = [UE.torch.matmul(cams[:, v, 1, :3, :3], cams[:, v, 0, :3, :4]) for v in range(V)]
    projs = []
    for v in range(V):
        pv = cams[:, v, 0].clone()
        pv[:, :3, :4] = P[v]
        projs.append(pv)
    warped = [UM.homo_warping_new(feats[v], projs[v], projs[0], hyp) for v in range(1, V)]
    sims = [(w_.view(1, G, C // G, D, H, W) * feats[0].view(1, G, C // G, 1, H, W)).mean(2) for w_ in warped]
    num = sum(s * wts[:, i].unsqueeze(1).unsqueeze(1) for i, s in enumerate(sims))
    den = sum(wts[:, i].unsqueeze(1).unsqueeze(1) for i in range(V - 1))
    save("warp_corr", feats=torch.stack(feats), cams=cams, hyp=hyp, wts=wts, G=G,
         warped1=warped[0], sims=torch.stack(sims), agg=num / (den + 1e-6))

    # ---- GetCost_initvolume (a5) and GetCost / pro_bilinear_sampler (a6, a7)
    C, H, W, V, D = 16, 24, 32, 5, 8
    feats, cams, _, wts = synthetic.microbench_inputs(C, D, H, W, views=V, seed=5)
    cur = 500.0 + 300.0 * torch.rand(1, 1, H, W, generator=g)
    interval = torch.full((1, 1, 1, 1), (1 / 425.0 - 1 / 935.0) / 384 * 2)
    dmin = torch.full((1, 1, 1, 1), 1 / 935.0)
    dmax = torch.full((1, 1, 1, 1), 1 / 425.0)
    gci = UE.GetCost_initvolume().eval()
    sim, samples = gci(cur, feats, cams, interval, dmax, dmin, wts, CostNum=D, Inverse=True, G=1)
    sim_nw, _ = gci(cur, feats, cams, interval, dmax, dmin, None, CostNum=D, Inverse=True, G=1)
    save("local_volume", feats=torch.stack(feats), cams=cams, cur=cur, interval=interval, wts=wts,
         sim=sim, samples=samples, sim_noweights=sim_nw)

    vol_a = torch.randn(1, D, H, W, generator=g)
    vol_b = torch.randn(1, D, H, W, generator=g)
    pro = [vol_b.permute(0, 2, 3, 1).reshape(H * W, 1, 1, D), vol_a.permute(0, 2, 3, 1).reshape(H * W, 1, 1, D)]
    vmax, vmin = samples[:, 0:1], samples[:, -1:]
    cur2 = cur * (1 + 0.01 * torch.randn(1, 1, H, W, generator=g))
    gc = UE.GetCost().eval()
    out6 = gc(cur2, pro, feats, cams, interval, dmax, dmin, wts, CostNum=3, Inverse=True, G=1,
              depth_max_cur_volume=vmax, depth_min_cur_volume=vmin)
    gmin = torch.full((1, 1, 1, 1), 425.0)
    gmax = torch.full((1, 1, 1, 1), 935.0)
    out6_global = gc(cur2, pro, feats, cams, interval * 6, dmax, dmin, wts, CostNum=3, Inverse=True, G=1,
                     depth_max_cur_volume=gmax, depth_min_cur_volume=gmin)
    look = UE.pro_bilinear_sampler(pro[1], samples[:, :, ::1], gmin, gmax)
    save("lookup", vol_raw=vol_a, vol_reg=vol_b, cur=cur2, interval=interval, vmin=vmin, vmax=vmax,
         out6=out6, out6_global=out6_global, samples=samples, look_global=look)

    # ---- 3-D regularization nets (a9, a10) with random weights and non-trivial BN statistics
    reg = UM.CostRegNet_2_sample_FPN3D_Fast(1, 8).eval()
    csp = UM.cost_up_small(1, 8).eval()
    randomize_bn(reg, g)
    randomize_bn(csp, g)
    x = torch.randn(1, 1, 16, 12, 20, generator=g)
    y, pro_feat = reg(x)
    xs = torch.randn(1, 1, 8, 12, 16, generator=g)
    prev = torch.randn(1, 1, 8, 6, 8, generator=g)
    up, mid = csp(xs, prev)
    save("regnets", x=x, y=y, pro=pro_feat, xs=xs, prev=prev, up=up, mid=mid,
         **module_arrays(reg, "reg__"), **module_arrays(csp, "csp__"))

    # ---- DepthNet.forward (a2-a4, a9, a11, a12) with the same nets
    C, H, W, V, D = 32, 16, 24, 5, 48
    feats, cams, _, _ = synthetic.microbench_inputs(C, D, H, W, views=V, seed=9)
    pwn = torch.nn.Sequential(UM.ConvBnReLU(1, 16), UM.ConvBnReLU(16, 16), UM.ConvBnReLU(16, 8),
                              torch.nn.Conv2d(8, 1, 1), torch.nn.Sigmoid()).eval()
    randomize_bn(pwn, g)
    dv = torch.linspace(1 / 935.0, 1 / 425.0, 384).unsqueeze(0)
    hyp = 1.0 / UM.get_depth_range_samples(dv, D, None, "cpu", torch.float32, [1, H, W])
    out = UE.DepthNet().eval()([f * 0.5 for f in feats], cams, hyp, D, reg, pwn, G=1)
    save("stage1", feats=torch.stack(feats) * 0.5, cams=cams, hyp=hyp, depth=out["depth"],
         conf=out["photometric_confidence"], view_weights=out["view_weights"], reg_volume=out["reg_volume"],
         volume=out["volume"], **module_arrays(pwn, "pwn__"))

    # ---- whole model with the shipped DTU checkpoint (weights stored de-duplicated)
    args = types.SimpleNamespace(ndepths="48,8,8", GRUiters="3,3,3", CostNum=3)
    model = UE.Effi_MVS_plus(args).eval()
    sd = torch.load("/root/reference/checkpoints/Effi_MVS_plus/model_dtu.ckpt", map_location="cpu", weights_only=False)["model"]
    model.load_state_dict(sd, strict=True)
    keep = {k: v for k, v in sd.items() if "num_batches" not in k
            and not k.startswith(("update_block.", "CSP_R.", "CSP_C."))}
    torch.save(keep, os.path.join(HERE, "dtu_weights.pt"))
    s = synthetic.make_sample("plumbing", seed=0, width=256, height=192)
    out = model(s["imgs"], s["proj_matrices"], s["depth_values"])
    save("model_forward", conf=out["photometric_confidence"], width=256, height=192, seed=0,
         **{"depth{:02d}".format(i): d for i, d in enumerate(out["depth"])})

    # ---- fusion (a13-a15)
    h, w, v = 48, 64, 4
    E, K = synthetic.camera_ring(v + 1, w, h)
    depths = synthetic.render_plane_scene(E, K, w, h, noise=0.15, seed=2)
    depths[2, 10:20, 10:30] = 0.0                       # invalid source depth
    depths[0, 30:34, 40:50] *= 1.05                     # inconsistent reference depth
    cams = synthetic.stage_cameras(E, K, 1)["stage4"]
    ref_depth, srcs_depth = depths[0][None, None], depths[1:][None, :, None]
    ref_cam, srcs_cam = cams[:, 0], cams[:, 1:]
    conf = torch.rand(1, h // 2, w // 2, generator=g)
    for tag, (dist_base, rel_base, prob_thr) in {"mm": (2, 6, 0.3), "tank": (4, 6000, 0.5)}.items():
        xyd, a, b = ufusion.get_reproj_dynamic(ref_depth, srcs_depth, ref_cam, srcs_cam)
        masks, last = ufusion.vis_filter_dynamic(ref_depth, xyd, a, b, dist_base=dist_base, rel_diff_base=rel_base, thres_view=2)
        xyd_raw = xyd.clone()
        # test_tank.py:473-515, statement by statement
        cf = F.interpolate(conf.unsqueeze(1), size=[h, w], mode="nearest")
        prob_mask = (cf > prob_thr).squeeze(1)
        reproj_depth = xyd[:, :, -1]
        reproj_depth[~last.squeeze(2)] = 0
        geo_sums = masks.sum(dim=1)
        geo_sum = last.sum(dim=1)
        avg = (torch.sum(reproj_depth, dim=1, keepdim=True) + ref_depth) / (geo_sum + 1)
        geo = geo_sum >= v + 1
        for i in range(2, v + 1):
            geo = torch.logical_or(geo, geo_sums[:, i - 2] >= i)
        final = ufusion.bin_op_reduce([prob_mask, geo], torch.min)
        idx_img = ufusion.get_pixel_grids(h, w).unsqueeze(0)
        pts = ufusion.idx_cam2world(ufusion.idx_img2cam(idx_img, avg, ref_cam), ref_cam)[..., :3, 0].permute(0, 3, 1, 2)
        save("fusion_" + tag, ref_depth=ref_depth, srcs_depth=srcs_depth, ref_cam=ref_cam, srcs_cam=srcs_cam, conf=conf,
             dist_base=dist_base, rel_diff_base=rel_base, prob_threshold=prob_thr, thres_view=2,
             reproj_xyd=xyd_raw, masks=masks, final=final, geo=geo, depth_avg=avg, points=pts)


if __name__ == "__main__":
    main()
