"""Host-side (stock PyTorch) network pieces: algebraic rewrites must equal the plain upstream form."""
import pytest
import torch

import effimvs_b200  # noqa: F401
from effimvs_b200 import net


def _randomise_bn(mod):
    for m in mod.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.normal_()
            m.bias.data.normal_()


@pytest.mark.parametrize("chans,heads", [((8, 16, 32, 64), (32, 16, 8)), ((4, 8, 16, 32), (60, 40, 20))])
@pytest.mark.parametrize("fmt", [torch.contiguous_format, torch.channels_last])
def test_fused_topdown_equals_upstream_form(chans, heads, fmt):
    """FeaturePyramid's fused top-down path (sub-pixel head, composed lateral conv, bias map, addmm
    lateral) against the literal upstream sequence (models/module.py:395-410), incl. odd borders."""
    torch.manual_seed(0)
    f = net.FeaturePyramid(chans, heads).eval()
    _randomise_bn(f)
    x = torch.randn(2, 3, 64, 96).contiguous(memory_format=fmt)
    with torch.no_grad():
        f.fused_topdown = False
        want = f(x)
        f.fused_topdown = True
        got = f(x)
    for w, g in zip(want, got):
        assert w.shape == g.shape
        assert float((w - g).abs().max()) <= 2e-5 * float(w.abs().max())


def test_fused_topdown_cache_follows_weight_updates():
    torch.manual_seed(1)
    f = net.FeaturePyramid((4, 8, 16, 32), (12, 8, 4)).eval()
    x = torch.randn(1, 3, 32, 32)
    with torch.no_grad():
        f.fused_topdown = True
        a = f(x)[2].clone()
        f.out3.weight.mul_(2.0)
        f.inner2.bias.add_(1.0)
        b = f(x)[2]
        f.fused_topdown = False
        want = f(x)[2]
    assert float((b - want).abs().max()) <= 2e-5 * float(want.abs().max())
    assert float((a - want).abs().max()) > 1e-3
