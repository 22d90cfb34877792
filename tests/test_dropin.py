"""The upstream-named adapters of dropin.py, checked against UNPATCHED upstream on the CPU with the
oracle's hot-path table injected (upstream is only present in the build container, so this test is
skipped on the GPU box; the CUDA table behind the same adapters is covered by test_gpu_parity)."""
import os
import pytest
import torch

from oracle import upstream

pytestmark = pytest.mark.skipif(not upstream.available(), reason="oracle/_ref (upstream byte copy) not staged")


@pytest.mark.parametrize("fused_update", [False, True])
def test_patch_matches_unpatched_upstream(fused_update):
    """fused_update: the table additionally offers the update-block glue entries (plain-torch stand-ins here), so
    that BasicUpdateBlock.forward and upsample_depth are swapped too -- the wiring CudaHotPath uses on the GPU."""
    import effimvs_b200  # noqa: F401
    from effimvs_b200 import dropin, synthetic
    from oracle import hotpath as ohp
    from util import GOLDEN, TorchGlue

    sd = torch.load(os.path.join(GOLDEN, "dtu_weights.pt"), map_location="cpu")
    model = upstream.build_model(sd)
    table = ohp.OracleHotPath()
    if fused_update:
        class Table(ohp.OracleHotPath, TorchGlue):
            fused_update = True
        table = Table()
    s = synthetic.make_sample("plumbing", seed=2, width=256, height=192)
    with torch.no_grad():
        want = model(s["imgs"], s["proj_matrices"], s["depth_values"])
        restore = dropin.patch(model, hotpath=table)
        got = model(s["imgs"], s["proj_matrices"], s["depth_values"])
        restore()
        again = model(s["imgs"], s["proj_matrices"], s["depth_values"])
    for a, b, c in zip(got["depth"], want["depth"], again["depth"]):
        assert float((a - b).abs().max()) < 1e-3 * (935 - 425)
        assert torch.equal(b, c)                      # the patch is fully undone
    assert float((got["photometric_confidence"] - want["photometric_confidence"]).abs().max()) < 1e-4
