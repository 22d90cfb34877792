"""The upstream-named adapters of dropin.py, checked against UNPATCHED upstream on the CPU with the
oracle's hot-path table injected (upstream is only present in the build container, so this test is
skipped on the GPU box; the CUDA table behind the same adapters is covered by test_gpu_parity)."""
import os
import sys
import types

import pytest
import torch

UPSTREAM = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(UPSTREAM, "models")), reason="upstream checkout not present")


def test_patch_matches_unpatched_upstream():
    sys.path.insert(0, UPSTREAM)
    import models  # noqa: F401  (upstream)
    UE = sys.modules["models.Effi_MVS_plus"]
    import effimvs_b200  # noqa: F401
    from effimvs_b200 import dropin, synthetic
    from oracle import hotpath as ohp
    from util import GOLDEN

    args = types.SimpleNamespace(ndepths="48,8,8", GRUiters="3,3,3", CostNum=3)
    model = UE.Effi_MVS_plus(args).eval()
    sd = torch.load(os.path.join(GOLDEN, "dtu_weights.pt"), map_location="cpu")
    model.load_state_dict(sd, strict=False)
    s = synthetic.make_sample("plumbing", seed=2, width=256, height=192)
    with torch.no_grad():
        want = model(s["imgs"], s["proj_matrices"], s["depth_values"])
        restore = dropin.patch(model, hotpath=ohp.OracleHotPath())
        got = model(s["imgs"], s["proj_matrices"], s["depth_values"])
        restore()
        again = model(s["imgs"], s["proj_matrices"], s["depth_values"])
    for a, b, c in zip(got["depth"], want["depth"], again["depth"]):
        assert float((a - b).abs().max()) < 1e-3 * (935 - 425)
        assert torch.equal(b, c)                      # the patch is fully undone
    assert float((got["photometric_confidence"] - want["photometric_confidence"]).abs().max()) < 1e-4
