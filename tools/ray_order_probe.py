"""Which fp32 evaluation order reproduces torch.matmul(rot (B,3,3), xyz (B,3,HW)) on this GPU bit for bit?
(upstream models/module.py:324).  Used once to pick the order the warp kernels use for the ray."""
import torch
torch.backends.cuda.matmul.allow_tf32 = False
import sys
dev = "cuda"


def probe(H, W):
    y, x = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(H * W, device=dev))).unsqueeze(0)
    g = torch.Generator().manual_seed(0)
    rot = (torch.eye(3) + 0.05 * torch.randn(3, 3, generator=g)).unsqueeze(0).to(dev)
    rot[0, 0, 2] = 37.123
    rot[0, 1, 2] = -12.5
    want = torch.matmul(rot, xyz)[0]                      # (3, HW)
    r = rot[0].double()
    X, Y = xyz[0, 0].double(), xyz[0, 1].double()
    f32 = lambda t: t.float().double()
    def fma(a, b, c): return f32(a * b + c)
    cands = {
        "fma(r2,1,fma(r1,y,r0*x))": lambda r0, r1, r2: fma(r2, 1.0, fma(r1, Y, f32(r0 * X))),
        "fma(r0,x,fma(r1,y,r2))": lambda r0, r1, r2: fma(r0, X, fma(r1, Y, r2)),
        "fma(r1,y,fma(r0,x,r2))": lambda r0, r1, r2: fma(r1, Y, fma(r0, X, r2)),
        "(r0*x+r1*y)+r2 no fma": lambda r0, r1, r2: f32(f32(f32(r0 * X) + f32(r1 * Y)) + r2),
        "fma(r1,y,r0*x)+r2": lambda r0, r1, r2: f32(fma(r1, Y, f32(r0 * X)) + r2),
        "fma(r0,x,r1*y)+r2": lambda r0, r1, r2: f32(fma(r0, X, f32(r1 * Y)) + r2),
        "fma(r2,1,fma(r0,x,r1*y))": lambda r0, r1, r2: fma(r2, 1.0, fma(r0, X, f32(r1 * Y))),
    }
    for name, fn in cands.items():
        bad = 0
        for row in range(3):
            got = fn(r[row, 0], r[row, 1], r[row, 2]).float()
            bad += int((got != want[row]).sum())
        if bad == 0:
            print("H=%d W=%d HW=%d: %s" % (H, W, H * W, name))


if __name__ == "__main__":
    sizes = [(int(a.split("x")[0]), int(a.split("x")[1])) for a in sys.argv[1:]] or [(1184, 1600)]
    for H, W in sizes:
        probe(H, W)
