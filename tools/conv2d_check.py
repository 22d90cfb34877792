"""Correctness and timing of the tcgen05 TF32 3x3 convolution (csrc/conv2d_tc.cu) against torch on a B200.

Reference for correctness: weights and activations rounded to IEEE half as the packing kernel and the producers round
them, convolved in fp64 -- products of such numbers are exact in fp32, so only the accumulation order differs and the
bar is 5e-6 of max|ref|.  Timing: CUDA events around graph-free launches, L2 flushed, against cuDNN (TF32 allowed).

    python tools/conv2d_check.py [quick | case N | probe | scale | rows | timeline]

probe / scale / rows time the kernel with parts switched off (EFFIMVS_CONV2D_DEBUG bits) and over CTA / row counts; timeline
prints clock64 stamps of CTA 0's three roles.  The switches and the stamps are compiled out of the shipped kernel: build a
profiling library with -DEFFIMVS_CONV2D_PROFILING -DEFFIMVS_CONV2D_TIMELINE and point EFFIMVS_LIB at it.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import capi  # noqa: E402

lib = capi.lib


def tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def pack(w):
    cout, cin = w.shape[:2]
    buf = torch.empty(lib.effimvs_conv2d_tf32_packed_bytes(cin, cout) // 4, device=w.device, dtype=torch.float32)
    capi.check(lib.effimvs_conv2d_tf32_pack(w.contiguous().data_ptr(), cin, cout, buf.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return buf


def run(x0, x1, packed, bias, cout, mode, out, aux0=None, aux1=None):
    """x0 / x1 / out / aux: (B,H,W,C) views of channels-last storage (last stride 1)."""
    B, H, W, c0 = x0.shape
    ps = lambda t: t.stride(2)   # noqa: E731
    capi.check(lib.effimvs_conv2d_tf32(
        x0.data_ptr(), ps(x0), c0, x1.data_ptr() if x1 is not None else None, ps(x1) if x1 is not None else 0,
        x1.shape[3] if x1 is not None else 0, packed.data_ptr(), bias.data_ptr() if bias is not None else None, cout, B, H, W, mode,
        out.data_ptr(), ps(out), aux0.data_ptr() if aux0 is not None else None, ps(aux0) if aux0 is not None else 0,
        aux1.data_ptr() if aux1 is not None else None, ps(aux1) if aux1 is not None else 0, torch.cuda.current_stream().cuda_stream))


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def f16(x):
    return x.half().float()


ACT_ROUND = f16      # operands are staged / packed as IEEE half, round to nearest


def ref_conv(xs, w):
    x = torch.cat(xs, dim=3).permute(0, 3, 1, 2)
    return F.conv2d(ACT_ROUND(x).double(), f16(w).double(), padding=1).permute(0, 2, 3, 1)


def check(B, H, W, c0, c1, cout, mode, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    # inputs as channel slices of wider maps, to exercise pixel strides
    big0 = torch.randn(B, H, W, c0 + 16, device=dev, generator=g)
    x0 = big0[..., 8:8 + c0]
    x1 = torch.randn(B, H, W, c1, device=dev, generator=g) if c1 else None
    w = torch.randn(cout, c0 + c1, 3, 3, device=dev, generator=g) * 0.1
    bias = torch.randn(cout, device=dev, generator=g)
    acc = ref_conv([x0] + ([x1] if c1 else []), w)
    pk = pack(w)
    bar = 5e-6
    if mode in (capi.CONV2D_BIAS, capi.CONV2D_BIAS_RELU):
        outbig = torch.full((B, H, W, cout + 4), 7.0, device=dev)
        out = outbig[..., 4:]
        run(x0, x1, pk, bias, cout, mode, out)
        want = acc + bias.double()
        if mode == capi.CONV2D_BIAS_RELU:
            want = want.relu()
        assert (outbig[..., :4] == 7.0).all()
        errs = [((out.double() - want).abs().max() / want.abs().max()).item()]
    elif mode == capi.CONV2D_ADD_RELU:
        add = torch.randn(B, H, W, cout, device=dev, generator=g)
        out = torch.empty(B, H, W, cout, device=dev)
        run(x0, x1, pk, None, cout, mode, out, aux0=add)
        want = (acc + add.double()).relu()
        errs = [((out.double() - want).abs().max() / want.abs().max()).item()]
    elif mode == capi.CONV2D_GRU_GATES:
        h = cout // 2
        hprev = torch.randn(B, H, W, h, device=dev, generator=g)
        z = torch.empty(B, H, W, h, device=dev)
        out = torch.empty(B, H, W, h, device=dev)
        run(x0, x1, pk, bias, cout, mode, out, aux0=hprev, aux1=z)
        pre = acc + bias.double()
        wz, wr = torch.sigmoid(pre[..., :h]), torch.sigmoid(pre[..., h:]) * hprev.double()
        errs = [(z.double() - wz).abs().max().item(), ((out.double() - wr).abs().max() / wr.abs().max()).item()]
        bar = 5e-6
    else:
        h = cout
        z = torch.rand(B, H, W, h, device=dev, generator=g)
        net = torch.randn(B, H, W, h, device=dev, generator=g)
        want = (1 - z.double()) * net.double() + z.double() * torch.tanh(acc + bias.double())
        run(x0, x1, pk, bias, cout, mode, net, aux0=z)
        errs = [((net.double() - want).abs().max() / want.abs().max()).item()]
        bar = 5e-6
    torch.cuda.synchronize()
    ok = all(e < bar for e in errs)
    print("{} B{} {}x{} cin {}+{} cout {} mode {}: err {}".format("ok  " if ok else "FAIL", B, H, W, c0, c1, cout, mode,
                                                              " ".join("{:.1e}".format(e) for e in errs)), flush=True)
    return ok


def _graph_ms(body, n):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            body()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best / n


def timeit(fn, flush, n=10):
    """microseconds per call with a cold L2 (graph of n x [flush, call] minus graph of n x [flush]) and back to back (warm)."""
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        t_flush = _graph_ms(lambda: flush.zero_(), n)
        t_cold = _graph_ms(lambda: (flush.zero_(), fn()), n) - t_flush
        t_warm = _graph_ms(fn, n)
    torch.cuda.synchronize()
    return t_cold * 1e3, t_warm * 1e3


def bench(H, W, cin, cout, flush, mode="relu"):
    dev = "cuda"
    x = torch.randn(1, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.1
    bias = torch.randn(cout, device=dev)
    pk = pack(w)
    h = cout // 2 if mode == "gates" else cout
    out = torch.empty(1, H, W, h, device=dev)
    aux0 = torch.rand(1, H, W, h, device=dev)
    aux1 = torch.empty(1, H, W, h, device=dev)
    m = {"relu": capi.CONV2D_BIAS_RELU, "add": capi.CONV2D_ADD_RELU, "gates": capi.CONV2D_GRU_GATES, "update": capi.CONV2D_GRU_UPDATE}[mode]
    t_own, w_own = timeit(lambda: run(x, None, pk, bias, cout, m, out, aux0 if mode != "relu" else None, aux1 if mode == "gates" else None), flush)
    xn = x.permute(0, 3, 1, 2)
    wn = w.contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.allow_tf32 = True
    t_cudnn, w_cudnn = timeit(lambda: torch.cudnn_convolution_relu(xn, wn, bias, (1, 1), (1, 1), (1, 1), 1), flush)
    mb = 4e-6 * H * W * (cin + cout)
    print("time {}x{} {}->{} {}: own cold {:.1f} us ({:.0f} GB/s) warm {:.1f} us | cudnn-tf32 conv+relu alone cold {:.1f} warm {:.1f} us".format(
        H, W, cin, cout, mode, t_own, mb / t_own * 1e3, w_own, t_cudnn, w_cudnn), flush=True)


def probe():
    """Where the time goes: the kernel with parts switched off (EFFIMVS_CONV2D_DEBUG), CTAs per SM, rows per unit."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for H, W, cin, cout in [(4, 128, 32, 32), (15, 128, 32, 32), (15, 1280, 32, 32), (592, 800, 32, 32), (296, 400, 64, 64)]:
        x = torch.randn(1, H, W, cin, device="cuda")
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
        bias = torch.randn(cout, device="cuda")
        pk = pack(w)
        out = torch.empty(1, H, W, cout, device="cuda")
        fn = lambda: run(x, None, pk, bias, cout, capi.CONV2D_BIAS_RELU, out)   # noqa: E731
        for env in [{}, {"EFFIMVS_CONV2D_DEBUG": "1"}, {"EFFIMVS_CONV2D_DEBUG": "2"}, {"EFFIMVS_CONV2D_DEBUG": "4"},
                    {"EFFIMVS_CONV2D_DEBUG": "12"}, {"EFFIMVS_CONV2D_DEBUG": "3"}, {"EFFIMVS_CONV2D_DEBUG": "7"},
                    {"EFFIMVS_CONV2D_DEBUG": "15"}, {"EFFIMVS_CONV2D_CTAS": "1"}, {"EFFIMVS_CONV2D_ROWS": "8"},
                    {"EFFIMVS_CONV2D_ROWS": "30"}, {"EFFIMVS_CONV2D_RING": "4"}, {"EFFIMVS_CONV2D_STAGES": "4"}]:
            if H < 100 and env and env != {"EFFIMVS_CONV2D_DEBUG": "15"}:
                continue
            for k in ("EFFIMVS_CONV2D_DEBUG", "EFFIMVS_CONV2D_CTAS", "EFFIMVS_CONV2D_ROWS", "EFFIMVS_CONV2D_RING", "EFFIMVS_CONV2D_STAGES"):
                os.environ.pop(k, None)
            os.environ.update(env)
            cold, warm = timeit(fn, flush)
            print("probe {}x{} {}->{} {}: cold {:.1f} warm {:.1f} us".format(H, W, cin, cout, env, cold, warm), flush=True)
    return 0


def scale():
    """Identical work per CTA (one unit of 15 rows x 128 pixels), growing number of CTAs."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cin = cout = 32
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
    bias = torch.randn(cout, device="cuda")
    pk = pack(w)
    for k in (1, 148, 296):
        H, W = 15, 128 * k
        x = torch.randn(1, H, W, cin, device="cuda")
        out = torch.empty(1, H, W, cout, device="cuda")
        fn = lambda: run(x, None, pk, bias, cout, capi.CONV2D_BIAS_RELU, out)   # noqa: E731
        res = []
        for dbg in ("0", "15", "14", "1"):
            os.environ["EFFIMVS_CONV2D_DEBUG"] = dbg
            os.environ["EFFIMVS_CONV2D_ROWS"] = "15"
            res.append("dbg{} {:.1f}".format(dbg, timeit(fn, flush)[1]))
        print("scale units {}: warm us: {}".format(k, "  ".join(res)), flush=True)
    return 0


def timeline():
    """clock64 stamps of CTA 0's three roles for one unit of 15 rows (profiling hook effimvs_conv2d_debug_timeline)."""
    import ctypes
    cin = cout = 32
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
    bias = torch.randn(cout, device="cuda")
    pk = pack(w)
    os.environ["EFFIMVS_CONV2D_ROWS"] = "15"
    for k, dbg in ((1, "0"), (1, "15"), (296, "0")):
        os.environ["EFFIMVS_CONV2D_DEBUG"] = dbg
        x = torch.randn(1, 15, 128 * k, cin, device="cuda")
        out = torch.empty(1, 15, 128 * k, cout, device="cuda")
        buf = torch.zeros(3, 64, 4, dtype=torch.int64, device="cuda")
        run(x, None, pk, bias, cout, capi.CONV2D_BIAS_RELU, out)
        torch.cuda.synchronize()
        lib.effimvs_conv2d_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
        run(x, None, pk, bias, cout, capi.CONV2D_BIAS_RELU, out)
        torch.cuda.synchronize()
        lib.effimvs_conv2d_debug_timeline(ctypes.c_void_p(0))
        t = buf.cpu()
        t0 = int(t[t > 0].min())
        print("timeline units {} debug {} (clocks since the first stamp)".format(k, dbg))
        for step in range(18):
            row = []
            for role, name in ((0, "prod"), (1, "mma"), (2, "epi")):
                row.append(name + " " + " ".join("{:6d}".format(int(v) - t0 if v > 0 else -1) for v in t[role, step]))
            print("  step {:2d}: {}".format(step, " | ".join(row)), flush=True)
    return 0


def rows():
    """One CTA, one unit of H rows x 128 pixels: time against H (slope = per-row cost of the slowest role, intercept = launch + prologue)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for cin, cout, ctas in ((32, 32, "3"), (32, 32, "2"), (32, 32, "1"), (16, 16, "1")):
        os.environ["EFFIMVS_CONV2D_CTAS"] = ctas
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
        bias = torch.randn(cout, device="cuda")
        pk = pack(w)
        for H in (8, 64):
            x = torch.randn(1, H, 128, cin, device="cuda")
            out = torch.empty(1, H, 128, cout, device="cuda")
            fn = lambda: run(x, None, pk, bias, cout, capi.CONV2D_BIAS_RELU, out)   # noqa: E731
            res = []
            for dbg in ("0", "15", "14", "1", "2", "4"):
                os.environ["EFFIMVS_CONV2D_DEBUG"] = dbg
                os.environ["EFFIMVS_CONV2D_ROWS"] = str(H)
                res.append("dbg{} {:.1f}".format(dbg, timeit(fn, flush)[1]))
            print("rows {}->{} ctas/SM cfg {} H={}: warm us: {}".format(cin, cout, ctas, H, "  ".join(res)), flush=True)
    return 0


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "rows":
        return rows()
    if len(sys.argv) > 1 and sys.argv[1] == "timeline":
        return timeline()
    if len(sys.argv) > 1 and sys.argv[1] == "scale":
        return scale()
    if len(sys.argv) > 1 and sys.argv[1] == "probe":
        return probe()
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ok = True
    M = capi
    global ACT_ROUND
    if os.environ.get("CHECK_ACT_RNA"):
        ACT_ROUND = tf32
    cases = [
        (1, 8, 40, 16, 0, 16, M.CONV2D_BIAS),          # one partial strip, one unit
        (1, 37, 200, 16, 0, 16, M.CONV2D_BIAS_RELU),   # two strips
        (2, 19, 130, 32, 0, 32, M.CONV2D_BIAS_RELU),   # batch, strip of 2 pixels
        (1, 64, 128, 16, 16, 32, M.CONV2D_GRU_GATES),  # two segments in one phase
        (1, 64, 300, 32, 32, 64, M.CONV2D_GRU_GATES),  # two phases, one segment each
        (1, 33, 257, 16, 16, 16, M.CONV2D_GRU_UPDATE),
        (1, 50, 100, 32, 0, 16, M.CONV2D_ADD_RELU),
        (1, 50, 100, 64, 0, 32, M.CONV2D_ADD_RELU),
        (1, 30, 70, 48, 0, 48, M.CONV2D_BIAS_RELU),    # three phases of 16
        (1, 30, 70, 16, 0, 12, M.CONV2D_BIAS),         # cout not a multiple of 16
        (1, 2, 5, 16, 0, 16, M.CONV2D_BIAS),           # tiny
    ]
    if not quick:
        cases += [
            (1, 592, 800, 32, 0, 32, M.CONV2D_BIAS_RELU),
            (1, 592, 800, 16, 16, 32, M.CONV2D_GRU_GATES),
            (1, 296, 400, 64, 0, 64, M.CONV2D_BIAS_RELU),
            (1, 296, 400, 32, 32, 32, M.CONV2D_GRU_UPDATE),
            (1, 148, 200, 96, 0, 48, M.CONV2D_ADD_RELU),
            (1, 148, 200, 48, 0, 96, M.CONV2D_BIAS_RELU),
        ]
    if len(sys.argv) > 2 and sys.argv[1] == "case":
        cases = [cases[int(sys.argv[2])]]
        quick = True
    for c in cases:
        ok = check(*c) and ok
    if not quick:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for H, W, cin, cout in [(592, 800, 32, 32), (592, 800, 32, 16), (592, 800, 16, 16), (592, 800, 16, 32),
                                (296, 400, 64, 64), (296, 400, 64, 32), (296, 400, 32, 32), (296, 400, 32, 64),
                                (148, 200, 96, 48), (148, 200, 48, 48), (148, 200, 48, 96)]:
            bench(H, W, cin, cout, flush)
        for H, W, cin, cout, mode in [(592, 800, 32, 16, "add"), (592, 800, 32, 32, "gates"), (592, 800, 32, 16, "update"),
                                      (296, 400, 64, 32, "add"), (296, 400, 64, 64, "gates"), (296, 400, 64, 32, "update")]:
            for pf in ("0", "4096"):
                os.environ["EFFIMVS_CONV2D_DEBUG"] = pf
                bench(H, W, cin, cout, flush, mode)
        os.environ["EFFIMVS_CONV2D_DEBUG"] = "0"
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
