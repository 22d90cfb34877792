"""One launch of conv2d_tc for ncu: python tools/conv2d_one.py H W cin cout [rows]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import capi, ops  # noqa: E402

H, W, cin, cout = (int(v) for v in sys.argv[1:5])
if len(sys.argv) > 5:
    os.environ["EFFIMVS_CONV2D_ROWS"] = sys.argv[5]
x = torch.randn(1, cin, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
b = torch.randn(cout, device="cuda")
out = torch.empty(1, cout, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
pk = ops.conv2d_tc_pack(w)
for _ in range(4):
    ops.conv2d_tc(x, None, pk, b, cout, capi.CONV2D_BIAS_RELU, out, None, None)
torch.cuda.synchronize()
