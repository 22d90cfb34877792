"""Times single tcgen05 layers with parts of the kernel disabled (EFFIMVS_TC_DEBUG bits) to find the
binding stage of the pipeline.  Results with debug bits set are numerically meaningless."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import effimvs_b200
from effimvs_b200 import capi, ops

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

cases = [("conv1 8->8 s1 48x148x200", 8, 8, 1, False, (48, 148, 200)),
         ("conv3 16->16 s1 24x74x100", 16, 16, 1, False, (24, 74, 100)),
         ("conv7 deconv 16->8 24x74x100", 16, 8, 2, True, (24, 74, 100)),
         ("csp conv1 16->8 8x296x400", 16, 8, 1, False, (8, 296, 400))]
for prec in (capi.PREC_BF16, capi.PREC_BF16X3):
    for name, cin, cout, sd, tr, (D, H, W) in cases:
        x = torch.randn(1, cin, D, H, W, device="cuda")
        w = torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3), device="cuda") * 0.1
        b = torch.randn(cout, device="cuda")
        row = []
        for dbg in (0, 1, 2, 3, 4, 8, 15):
            os.environ["EFFIMVS_TC_DEBUG"] = str(dbg)
            row.append("%d:%.0f" % (dbg, t(lambda: ops.conv3d_bf16(x, w, b, None, sd, tr, True, prec))))
        print("prec", prec, name, " us incl. to/from c8 conversion ->", " ".join(row))
os.environ["EFFIMVS_TC_DEBUG"] = "0"
