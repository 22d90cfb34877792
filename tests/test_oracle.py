"""CPU tests: the oracle restatement replayed against the golden fixtures that
tests/golden/make_golden.py produced by running upstream itself (the parity pin),
plus the explicit ATen restatements (grid_sample, conv3d, conv_transpose3d) against
the ATen calls they restate."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import fusion as ofu
from oracle import hotpath as ohp
from util import cspnet, dtu_model, golden, pixelwise, regnet, rel_max

torch.set_grad_enabled(False)


def test_warp_corr_matches_upstream():
    g = golden("warp_corr")
    feats, cams, hyp, G = list(g["feats"]), g["cams"], g["hyp"], g["G"]
    P = [ohp.compose_projection(cams[:, v]) for v in range(len(feats))]
    w1 = ohp.homo_warp(feats[1], P[1], P[0], hyp)
    assert rel_max(w1.reshape(g["warped1"].shape), g["warped1"]) < 1e-6
    sims = [ohp.view_similarity(feats[0], feats[v], cams[:, 0], cams[:, v], hyp, G) for v in range(1, len(feats))]
    assert rel_max(torch.stack(sims), g["sims"]) < 1e-6
    agg = ohp.weighted_aggregate(sims, [g["wts"][:, i:i + 1] for i in range(len(sims))])
    assert rel_max(agg, g["agg"]) < 1e-6


def test_local_volume_matches_upstream():
    g = golden("local_volume")
    feats = list(g["feats"])
    sim, samples = ohp.local_volume(g["cur"], feats, g["cams"], g["interval"], g["wts"], 8, 1)
    assert torch.equal(samples, g["samples"])
    assert rel_max(sim, g["sim"]) < 1e-6
    sim2, _ = ohp.local_volume(g["cur"], feats, g["cams"], g["interval"], None, 8, 1)
    assert rel_max(sim2, g["sim_noweights"]) < 1e-6


def test_lookup_matches_upstream():
    g = golden("lookup")
    out = ohp.dynamic_cost(g["cur"], g["vol_raw"], g["vol_reg"], g["interval"], g["vmin"], g["vmax"], 3)
    assert rel_max(out, g["out6"]) < 1e-6
    gmin, gmax = torch.full((1, 1, 1, 1), 425.0), torch.full((1, 1, 1, 1), 935.0)
    out = ohp.dynamic_cost(g["cur"], g["vol_raw"], g["vol_reg"], g["interval"] * 6, gmin, gmax, 3)
    assert rel_max(out, g["out6_global"]) < 1e-6
    look = ohp.volume_lookup(g["vol_raw"], g["samples"], gmin, gmax)
    assert rel_max(look, g["look_global"]) < 1e-6


def test_regnets_match_upstream():
    g = golden("regnets")
    y, pro = ohp.cost_regularization(regnet(g), g["x"])
    assert rel_max(y, g["y"]) < 1e-5 and rel_max(pro, g["pro"]) < 1e-5
    up, mid = ohp.cross_scale_net(cspnet(g), g["xs"], g["prev"])
    assert rel_max(up, g["up"]) < 1e-5 and rel_max(mid, g["mid"]) < 1e-5


def test_stage1_matches_upstream():
    g = golden("stage1")
    r = golden("regnets")
    out = ohp.OracleHotPath().stage1(list(g["feats"]), g["cams"], g["hyp"], pixelwise(g), regnet(r), 1)
    assert rel_max(out["volume"], g["volume"]) < 1e-5
    assert rel_max(out["view_weights"], g["view_weights"]) < 1e-5
    assert rel_max(out["reg_volume"], g["reg_volume"]) < 1e-4
    assert float((out["depth"] - g["depth"]).abs().max()) < 1e-3 * (935 - 425)
    assert float((out["photometric_confidence"] - g["conf"]).abs().max()) < 1e-4


def test_model_forward_matches_upstream():
    from effimvs_b200 import synthetic
    g = golden("model_forward")
    s = synthetic.make_sample("plumbing", seed=g["seed"], width=g["width"], height=g["height"])
    out = dtu_model(ohp.OracleHotPath())(s["imgs"], s["proj_matrices"], s["depth_values"])
    assert len(out["depth"]) == 13
    for i, d in enumerate(out["depth"]):
        ref = g["depth{:02d}".format(i)]
        assert d.shape == ref.shape
        assert float((d - ref).abs().max()) < 1e-3 * (935 - 425), i
    assert float((out["photometric_confidence"] - g["conf"]).abs().max()) < 1e-4


@pytest.mark.parametrize("tag", ["mm", "tank"])
def test_fusion_matches_upstream(tag):
    g = golden("fusion_" + tag)
    out = ofu.fuse_view(g["ref_depth"], g["conf"], g["srcs_depth"], g["ref_cam"], g["srcs_cam"],
                        g["dist_base"], g["rel_diff_base"], g["thres_view"], g["prob_threshold"])
    ok = torch.isfinite(g["reproj_xyd"])
    assert torch.equal(torch.isfinite(out["reproj_xyd"]), ok)
    big = g["reproj_xyd"].abs() > 1e6                 # zero-depth samples: ~1e11 coordinates, compare loosely
    assert rel_max(out["reproj_xyd"][ok & ~big], g["reproj_xyd"][ok & ~big]) < 1e-5
    assert torch.equal(out["masks"], g["masks"].bool())
    assert torch.equal(out["final"], g["final"].bool()) and torch.equal(out["geo"], g["geo"].bool())
    assert rel_max(out["depth_avg"], g["depth_avg"]) < 1e-6
    assert rel_max(out["points"], g["points"]) < 1e-5


# ---- restated third-party arithmetic vs the ATen calls upstream makes -----------------
def test_grid_sample_restatement():
    gen = torch.Generator().manual_seed(0)
    img = torch.randn(2, 3, 7, 9, generator=gen)
    grid = torch.rand(2, 5, 6, 2, generator=gen) * 2.6 - 1.3
    grid[0, 0, 0] = float("nan")
    grid[0, 0, 1] = float("inf")
    grid[1, 2, 3] = torch.tensor([1.0, -1.0])
    a = ohp.grid_sample_zeros_ac(img, grid)
    b = F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    # NaN / inf coordinates: ATen's CUDA kernel (the one upstream runs on) skips every corner -> 0;
    # ATen's CPU kernel multiplies a masked 0 by a NaN weight -> NaN.  The restatement follows CUDA.
    assert torch.all(a[0, :, 0, :2] == 0) and torch.isnan(b[0, :, 0, :2]).all()
    a[0, :, 0, :2] = 0
    b[0, :, 0, :2] = 0
    assert torch.allclose(a, b, atol=1e-6)


@pytest.mark.parametrize("stride", [(1, 1, 1), (2, 2, 2), (1, 2, 2)])
def test_conv3d_restatement(stride):
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 8, 6, 10, generator=gen)
    w = torch.randn(5, 3, 3, 3, 3, generator=gen)
    assert torch.allclose(ohp.conv3d_taps(x, w, stride), F.conv3d(x, w, None, stride=stride, padding=1), atol=1e-4)


@pytest.mark.parametrize("stride,op", [((2, 2, 2), (1, 1, 1)), ((1, 2, 2), (0, 1, 1))])
def test_deconv3d_restatement(stride, op):
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(1, 4, 4, 3, 5, generator=gen)
    w = torch.randn(4, 2, 3, 3, 3, generator=gen)
    ref = F.conv_transpose3d(x, w, None, stride=stride, padding=1, output_padding=op)
    assert torch.allclose(ohp.deconv3d_taps(x, w, stride, op), ref, atol=1e-4)


def test_restatements_through_hotpath():
    """Whole stage-1 path with ATEN=False (explicit restatements) == ATEN=True."""
    g, r = golden("stage1"), golden("regnets")
    args = (list(g["feats"]), g["cams"], g["hyp"], pixelwise(g), regnet(r), 1)
    a = ohp.OracleHotPath().stage1(*args)
    ohp.ATEN = False
    try:
        b = ohp.OracleHotPath().stage1(*args)
    finally:
        ohp.ATEN = True
    assert rel_max(b["volume"], a["volume"]) < 1e-5 and rel_max(b["reg_volume"], a["reg_volume"]) < 1e-4


# ---- DTU geometric filter -----------------------------------------------------
def test_remap_restatement_equals_cv2():
    """oracle.dtu_filter.remap_bilinear against cv2.remap itself (OpenCV is the absent third-party dependency of
    this row; it is present in the build container, skipped where it is not)"""
    cv2 = pytest.importorskip("cv2")
    import numpy as np
    from oracle import dtu_filter as o
    g = np.random.default_rng(0)
    src = g.random((37, 53)).astype(np.float32) * 900
    mx = (g.random((64, 80)) * 60 - 4).astype(np.float32)
    my = (g.random((64, 80)) * 45 - 4).astype(np.float32)
    mx[0, :8] = [0.0, 52.0, 51.984375, -1.0, -0.015625, 52.5, 1e9, -1e9]      # borders, exact grid points, far outside
    my[0, :8] = [0.0, 36.0, 35.984375, -1.0, 36.5, 10.015625, 3.0, 3.0]
    want = cv2.remap(src, mx, my, interpolation=cv2.INTER_LINEAR)
    got = o.remap_bilinear(src, mx, my)
    assert np.array_equal(got, want)


def test_dtu_filter_matches_upstream():
    import numpy as np
    from oracle import dtu_filter as o
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dtu_filter.npz"))
    d, K, E = z["depths"], z["K"], z["E"]
    v = d.shape[0] - 1
    out = o.filter_view(d[0], z["confidence"], d[1:], K, E[0], [K] * v, list(E[1:]), float(z["conf_thres"]))
    assert np.array_equal(out["masks"], z["masks"])
    assert np.array_equal(out["reproj_depth"], z["reproj_depth"])
    assert np.array_equal(out["final"], z["final_mask"]) and np.array_equal(out["geo"], z["geo_mask"])
    assert np.array_equal(out["depth_avg"], z["depth_avg"])
    assert np.allclose(out["points"], z["points"], rtol=0, atol=1e-3)
