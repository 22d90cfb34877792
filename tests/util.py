"""Shared helpers for the tests: golden-fixture loading and module reconstruction."""
import os
import types

import numpy as np
import torch

import effimvs_b200  # noqa: F401
from effimvs_b200 import net

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name, device="cpu"):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]).to(device) if z[k].ndim else z[k].item() for k in z.files}


def load_module(mod, arrays, prefix, device="cpu"):
    sd = {k[len(prefix):].replace("__", "."): v for k, v in arrays.items() if k.startswith(prefix)}
    missing = mod.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if "num_batches" not in k], missing
    assert not missing.unexpected_keys, missing
    return mod.to(device).eval()


def regnet(arrays, device="cpu"):
    return load_module(net.RegNet3D(1, 8), arrays, "reg__", device)


def cspnet(arrays, device="cpu"):
    return load_module(net.CrossScaleNet3D(1, 8), arrays, "csp__", device)


def pixelwise(arrays, device="cpu"):
    m = torch.nn.Sequential(net.ConvBNReLU2d(1, 16, 3, 1, 1), net.ConvBNReLU2d(16, 16, 3, 1, 1),
                            net.ConvBNReLU2d(16, 8, 3, 1, 1), torch.nn.Conv2d(8, 1, 1), torch.nn.Sigmoid())
    return load_module(m, arrays, "pwn__", device)


def dtu_model(hotpath, device="cpu", ndepths="48,8,8"):
    args = types.SimpleNamespace(ndepths=ndepths, GRUiters="3,3,3", CostNum=3)
    m = net.EffiMVSPlus(args, hotpath=hotpath)
    load_dtu_weights(m)
    return m.to(device).eval()


def load_dtu_weights(m):
    sd = torch.load(os.path.join(GOLDEN, "dtu_weights.pt"), map_location="cpu")
    full = dict(sd)
    for k, v in sd.items():          # re-create upstream's duplicate registrations
        for a, b in (("update_block_depth1.", "update_block.0."), ("update_block_depth2.", "update_block.1."),
                     ("update_block_depth3.", "update_block.2."), ("CSP_R1.", "CSP_R.0."), ("CSP_R2.", "CSP_R.1."),
                     ("CSP_C1.", "CSP_C.0."), ("CSP_C2.", "CSP_C.1.")):
            if k.startswith(a):
                full[b + k[len(a):]] = v
    res = m.load_state_dict(full, strict=False)
    assert not res.unexpected_keys and all("num_batches" in k for k in res.missing_keys), res


def rel_max(a, b):
    """max|a-b| / max|b|  -- the north_star's cost-volume metric (SURVEY.md section 7 'hard parts')."""
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
