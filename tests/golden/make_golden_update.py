"""Golden fixture for the ConvGRU update block + convex upsampling, generated
by running the UPSTREAM code (models/update.py BasicUpdateBlock, models/Effi_MVS_plus.py disp_to_depth and
upsample_depth, unmodified) on seeded inputs in the build container.

    python tests/golden/make_golden_update.py        # needs /root/reference; CPU only

The dynamic cost lookup is replaced by a fixed analytic function of the depth (cost_fn below, repeated in
the tests) so that the fixture exercises exactly the update-block arithmetic.
"""
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

import models  # noqa: E402,F401  (upstream)

UE = sys.modules["models.Effi_MVS_plus"]
UU = sys.modules["models.update"]


def cost_fn(depth, iter=0):
    """(B,1,H,W) depth -> (B,6,H,W): smooth, bounded, deterministic."""
    k = torch.arange(1, 7, dtype=depth.dtype).reshape(1, 6, 1, 1)
    return torch.sin(depth * (k * 0.01)) * 0.5


def main():
    g = torch.Generator().manual_seed(11)
    hidden, ctx, B, H, W = 16, 4, 2, 20, 28
    torch.manual_seed(11)    # torch's default (Kaiming-uniform) initialisation: well-conditioned activations
    blk = UU.BasicUpdateBlock(hidden_dim=hidden, cost_dim=2, ratio=2, context_dim=ctx, UpMask=True, cost_num=3).eval()
    for p in blk.mask.parameters():   # a softmax with some contrast
        p.data = p.data * 8.0
    net0 = torch.tanh(torch.randn(B, hidden, H, W, generator=g))
    context = torch.relu(torch.randn(B, ctx, H, W, generator=g))
    inv0 = torch.rand(B, 1, H, W, generator=g)
    dmin = torch.tensor([425.0, 400.0]).reshape(B, 1, 1, 1)
    dmax = torch.tensor([935.0, 1000.0]).reshape(B, 1, 1, 1)
    scale = partial(UE.disp_to_depth, min_depth=dmin, max_depth=dmax)
    with torch.no_grad():
        net, masks, invs = blk(net0, cost_fn, inv0, context, seq_len=3, scale_inv_depth=scale)
        depths = [scale(i)[1] for i in invs]
        up = UE.upsample_depth(invs[-1], masks[-1], ratio=2)
        depth_up = scale(up.unsqueeze(1))[1].squeeze(1)
    out = {"w__" + k.replace(".", "__"): v.numpy() for k, v in blk.state_dict().items()}
    out.update(net0=net0.numpy(), context=context.numpy(), inv0=inv0.numpy(), dmin=dmin.numpy(), dmax=dmax.numpy(),
               net=net.numpy(), mask=masks[-1].numpy(), up=up.numpy(), depth_up=depth_up.numpy())
    for i in range(3):
        out["inv{}".format(i + 1)] = invs[i].numpy()
        out["depth{}".format(i + 1)] = depths[i].numpy()
    np.savez_compressed(os.path.join(HERE, "update_block.npz"), **out)
    print("wrote update_block", {k: v.shape for k, v in out.items() if not k.startswith("w__")})


if __name__ == "__main__":
    main()
