"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every function that
include/effimvs.h declares, the ctypes table covers all of them, argument validation returns the
documented error codes (no compute is launched), and CPU tensors are rejected (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import effimvs_b200  # noqa: F401
from effimvs_b200 import capi, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "effimvs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(effimvs_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(capi.lib, n), "libeffimvs.so does not export " + n
        assert n in capi.SIGNATURES, "capi.SIGNATURES has no binding for " + n
    assert sorted(capi.SIGNATURES) == names          # and nothing is bound that the header does not declare


def test_argument_validation_returns_error_codes():
    lib = capi.lib
    assert lib.effimvs_version() >= 100
    # null pointers / bad sizes are rejected before any launch
    assert lib.effimvs_weighted_agg_f32(None, None, 1, 1, 1, 1, 1, None, None) == capi.EINVAL
    assert "null pointer" in capi.last_error()
    one = ctypes.c_void_p(16)
    assert lib.effimvs_volume_lookup_f32(one, one, one, one, 0, 3, 1, 8, 3, 4, 4, one, None) == capi.EINVAL
    assert "sample_stride" in capi.last_error()
    arr, keep = capi.ptr_array([16] * 4)
    arr9, keep9 = capi.ptr_array([16] * 9)
    assert lib.effimvs_warp_corr_agg_f32(one, arr, 4, one, one, 0, None, None, 1, 12, 8, 8, 4, 1, 0, one, None, None) == capi.EUNSUPPORTED
    assert "C=12" in capi.last_error()
    assert lib.effimvs_warp_corr_agg_f32(one, arr, 4, one, one, 0, None, None, 1, 8, 8, 8, 4, 3, 0, one, None, None) == capi.EINVAL
    assert lib.effimvs_costreg_fpn3d(one, arr9, arr9, 1, 6, 8, 8, 0, one, 1 << 30, one, None) == capi.EUNSUPPORTED   # D not a multiple of 4
    assert lib.effimvs_costreg_workspace_bytes(1, 48, 148, 200, capi.PREC_F32) > 0
    assert lib.effimvs_costreg_workspace_bytes(1, 48, 148, 200, capi.PREC_BF16X3) > lib.effimvs_costreg_workspace_bytes(1, 48, 148, 200, capi.PREC_BF16)
    # phased forms: unknown phase bits, RUN without data pointers, PREPARE-only accepts NULL data pointers up to the workspace check
    assert lib.effimvs_costreg_fpn3d_ex(one, arr9, arr9, 1, 8, 8, 8, capi.PREC_BF16X3, 4, one, 1 << 30, one, None) == capi.EINVAL
    assert "phases" in capi.last_error()
    assert lib.effimvs_costreg_fpn3d_ex(None, arr9, arr9, 1, 8, 8, 8, capi.PREC_BF16X3, capi.WS_RUN, one, 1 << 30, None, None) == capi.EINVAL
    assert lib.effimvs_costreg_fpn3d_ex(None, arr9, arr9, 1, 8, 8, 8, capi.PREC_BF16X3, capi.WS_PREPARE, one, 16, None, None) == capi.EWORKSPACE
    assert lib.effimvs_cost_up_small_ex(None, None, arr, arr, 1, 8, 8, 8, capi.PREC_BF16X3, capi.WS_PREPARE, one, 16, None, None) == capi.EWORKSPACE
    assert lib.effimvs_cost_up_small_ex(None, None, arr, arr, 1, 8, 8, 8, capi.PREC_BF16X3, 0, one, 16, None, None) == capi.EINVAL
    assert lib.effimvs_cost_up_small_ex(None, None, arr, arr, 1, 8, 8, 8, capi.PREC_F32, capi.WS_PREPARE, one, 1 << 30, None, None) == capi.OK
    with pytest.raises(capi.EffiMVSError):
        capi.check(capi.EWORKSPACE)
    # section 8(f) entry points
    assert lib.effimvs_gru_reset_f32(one, one, one, 10, 18, 16, one, None) == capi.EINVAL and "multiples of 4" in capi.last_error()
    assert lib.effimvs_gru_update_f32(one, one, None, one, one, 10, 16, 16, one, None) == capi.EINVAL
    assert lib.effimvs_gru_delta_f32(one, None, one, one, one, 1, 16, one, one, None) == capi.EINVAL        # pre without bias
    assert lib.effimvs_delta_head_f32(one, one, one, one, one, one, 1, 24, 8, 8, one, one, None) == capi.EUNSUPPORTED
    assert "hidden channels 24" in capi.last_error()
    assert lib.effimvs_delta_head_f32(one, one, None, one, one, one, 1, 16, 8, 8, one, one, None) == capi.EINVAL
    assert lib.effimvs_convex_upsample_f32(one, None, 0.25, one, one, one, 1, 4, 4, 8, one, one, None) == capi.EUNSUPPORTED
    assert "ratio=8" in capi.last_error()
    assert lib.effimvs_convex_upsample_conv_f32(one, 30, one, None, 0.25, one, one, one, 1, 4, 4, 2, one, one, None) == capi.EUNSUPPORTED
    assert "K=30" in capi.last_error()
    assert lib.effimvs_encoder_head_f32(one, one, one, one, one, one, 1, 6, 20, 4, 4, one, None) == capi.EUNSUPPORTED   # hidden not a multiple of 16
    assert lib.effimvs_encoder_tail_f32(one, one, one, 10, 10, 16, one, None) == capi.EUNSUPPORTED                      # hm not a multiple of 4
    assert lib.effimvs_encoder_tail_ctx_f32(one, one, one, 20, 5, 1, one, one, 10, 12, 16, one, None) == capi.EUNSUPPORTED           # cx = 5
    assert lib.effimvs_encoder_tail_ctx_f32(one, one, ctypes.c_void_p(20), 20, 4, 1, one, one, 10, 12, 16, one, None) == capi.EINVAL  # misaligned ctx
    assert lib.effimvs_gru_init_f32(one, 10, 18, 4, one, None) == capi.EINVAL
    assert lib.effimvs_inv_init_f32(one, one, one, 1, 16, None, one, None) == capi.EINVAL
    assert lib.effimvs_depth_ranges_f32(one, 1, 1, 48, one, one, None) == capi.EINVAL                      # a single depth value
    assert lib.effimvs_gru_init_ctx_f32(one, 10, 16, 6, one, one, one, one, None) == capi.EINVAL            # cx not a multiple of 4
    assert lib.effimvs_gru_init_ctx_f32(one, 10, 16, 4, one, None, one, one, None) == capi.EINVAL           # no bias
    td, tf = (ctypes.c_double * 2)(1.0, 0.5), (ctypes.c_float * 2)(0.1, 0.2)
    assert lib.effimvs_dtu_filter_f32(one, one, one, one, td, tf, 2, 1, 11, 0.5, 0.75, 3, 8, 8, one, None, one, one, None, None, None) == capi.EUNSUPPORTED
    assert "non-decreasing" in capi.last_error()
    assert lib.effimvs_dtu_filter_f32(one, one, one, one, td, tf, 2, 1, 11, 0.5, 0.75, 40, 8, 8, one, None, one, one, None, None, None) == capi.EINVAL


def test_encoder_head_host_tables_layout():
    """effimvs_encoder_head_pack_host is a pure host function: the per-chunk tables (kernel parameters of the constant-bank
    encoder head) hold convd1's taps tap-major, convc1's rows padded to 8, and the two biases."""
    lib = capi.lib
    h, CD = 32, 6
    gen = torch.Generator().manual_seed(5)
    wc1, bc1 = torch.randn(h, CD, 1, 1, generator=gen), torch.randn(h, generator=gen)
    wd1, bd1 = torch.randn(h, 1, 7, 7, generator=gen), torch.randn(h, generator=gen)
    n = lib.effimvs_encoder_head_table_floats(h)
    assert n == 2 * 944 and lib.effimvs_encoder_head_table_floats(24) == 0
    tab = torch.full((n,), float("nan"))
    assert lib.effimvs_encoder_head_pack_host(wc1.data_ptr(), bc1.data_ptr(), wd1.data_ptr(), bd1.data_ptr(), CD, h, tab.data_ptr()) == capi.OK
    t = tab.reshape(2, 944)
    for k in range(2):
        ch = slice(16 * k, 16 * k + 16)
        assert torch.equal(t[k, :784].reshape(49, 16), wd1[ch, 0].reshape(16, 49).t())
        assert torch.equal(t[k, 784:784 + 16 * CD].reshape(CD, 16), wc1[ch, :, 0, 0].t())
        assert torch.count_nonzero(t[k, 784 + 16 * CD:912]) == 0
        assert torch.equal(t[k, 912:928], bc1[ch]) and torch.equal(t[k, 928:], bd1[ch])
    assert lib.effimvs_encoder_head_pack_host(wc1.data_ptr(), bc1.data_ptr(), wd1.data_ptr(), bd1.data_ptr(), 9, h, tab.data_ptr()) == capi.EUNSUPPORTED
    one = ctypes.c_void_p(1)
    assert lib.effimvs_encoder_head_hostw_f32(one, one, None, 1, 6, 16, 4, 4, one, None) == capi.EINVAL
    assert lib.effimvs_encoder_head_hostw_f32(one, one, one, 1, 6, 20, 4, 4, one, None) == capi.EUNSUPPORTED


def test_cpu_tensors_are_rejected_not_emulated():
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.volume_lookup(torch.zeros(1, 8, 4, 4), torch.ones(1, 3, 4, 4), torch.ones(1), torch.ones(1), 1)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.warp_corr_agg(torch.zeros(1, 8, 4, 4), [torch.zeros(1, 8, 4, 4)], torch.zeros(1, 1, 12), torch.ones(1, 2, 4, 4), 0, None, None, 2, 1, False)


def test_fake_implementations_give_shapes_without_a_device():
    with torch.device("meta"):
        zr, hx = torch.empty(2, 32, 6, 8), torch.empty(2, 32, 6, 8)
        assert torch.ops.effimvs.gru_reset(zr, torch.empty(16), hx).shape == (2, 32, 6, 8)
        assert torch.ops.effimvs.encoder_head(torch.empty(2, 6, 6, 8), torch.empty(2, 1, 6, 8), torch.empty(16, 6, 1, 1), torch.empty(16),
                                              torch.empty(16, 1, 7, 7), torch.empty(16)).shape == (2, 32, 6, 8)
        iv, dp = torch.ops.effimvs.delta_head(torch.empty(2, 16, 6, 8), torch.empty(1, 16, 3, 3), torch.empty(1), torch.empty(2, 1, 6, 8),
                                              torch.empty(2), torch.empty(2))
        assert iv.shape == (2, 1, 6, 8) and dp.shape == (2, 1, 6, 8)
        up, dep = torch.ops.effimvs.convex_upsample_conv(torch.empty(2, 32, 6, 8), torch.empty(36, 32, 1, 1), None, 0.25, torch.empty(2, 1, 6, 8),
                                                         torch.empty(2), torch.empty(2), 2)
        assert up.shape == (2, 12, 16) and dep.shape == (2, 12, 16)
        assert torch.ops.effimvs.gru_init(torch.empty(2, 20, 6, 8), 16).shape == (2, 32, 6, 8)
        assert torch.ops.effimvs.depth_ranges(torch.empty(2, 48), 48, [4.0, 2.0, 1.0]).shape == (2 * 56,)
        iv0, dp0 = torch.ops.effimvs.inv_init(torch.empty(2, 1, 6, 8), torch.empty(2), torch.empty(2))
        assert iv0.shape == (2, 1, 6, 8) and dp0.shape == (2, 1, 6, 8)
        hx0, term = torch.ops.effimvs.gru_init_ctx(torch.empty(2, 20, 6, 8), 16, torch.empty(16, 4, 1, 1), torch.empty(16))
        assert hx0.shape == (2, 32, 6, 8) and term.shape == (2, 16, 6, 8)
        up, dep = torch.ops.effimvs.convex_upsample(torch.empty(2, 36, 6, 8), None, 0.25, torch.empty(2, 1, 6, 8), torch.empty(2), torch.empty(2), 2)
        assert up.shape == (2, 12, 16) and dep.shape == (2, 12, 16)
        ref = torch.empty(2, 16, 24, 32)
        sim, hyp = torch.ops.effimvs.warp_corr_agg(ref, [ref, ref], torch.empty(2, 2, 12), torch.empty(2, 1, 24, 32), 2,
                                                   torch.empty(2), None, 8, 1, True)
        assert sim.shape == (2, 1, 8, 24, 32) and hyp.shape == (2, 8, 24, 32)
        ws = torch.empty(64, dtype=torch.uint8)
        assert torch.ops.effimvs.costreg_run(torch.empty(1, 1, 8, 8, 8), [ref] * 9, [ref] * 8, 2, ws).shape == (1, 1, 8, 8, 8)
        assert torch.ops.effimvs.cost_up_run(torch.empty(1, 1, 8, 8, 8), torch.empty(1, 1, 8, 4, 4), [ref] * 4, [ref] * 4, 2, ws).shape == (1, 1, 8, 8, 8)
        # ops without outputs (in-place on a caller-kept buffer) trace as no-ops
        assert torch.ops.effimvs.costreg_prepare([ref] * 9, [ref] * 8, 1, 8, 8, 8, 2, ws) is None
        assert torch.ops.effimvs.cost_up_prepare([ref] * 4, [ref] * 4, 1, 8, 8, 8, 2, ws) is None
        assert torch.ops.effimvs.encoder_tail_ctx(torch.empty(1, 12, 4, 4), torch.empty(16, 12, 1, 1), torch.empty(1, 20, 4, 4), 16, 4, True,
                                                  torch.empty(16, 4, 1, 1), torch.empty(16), torch.empty(1, 32, 4, 4)) is None
        out = torch.ops.effimvs.conv3d_bf16(torch.empty(1, 16, 4, 6, 8), torch.empty(16, 8, 3, 3, 3), None, None, 2, True, True, 2)
        assert out.shape == (1, 8, 8, 12, 16)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/effimvs.h compiles as C99 (no C++ or torch types in the boundary) and a C program linked against
    libeffimvs.so reaches the library: version, and an argument error reported through effimvs_last_error()."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "effimvs.h"\n'
        "int main(void) {\n"
        "    if (effimvs_version() < 100) return 1;\n"
        "    int rc = effimvs_weighted_agg_f32(NULL, NULL, 1, 1, 1, 1, 1, NULL, NULL);\n"
        "    if (rc != EFFIMVS_EINVAL) return 2;\n"
        '    if (!strstr(effimvs_last_error(), "null pointer")) return 3;\n'
        "    size_t n = effimvs_costreg_workspace_bytes(1, 48, 148, 200, EFFIMVS_PREC_BF16X3);\n"
        "    rc = effimvs_costreg_fpn3d_ex(NULL, NULL, NULL, 1, 48, 148, 200, EFFIMVS_PREC_BF16X3, EFFIMVS_WS_PREPARE, NULL, n, NULL, NULL);\n"
        "    if (rc != EFFIMVS_EINVAL) return 4;\n"
        '    printf("%zu\\n", n);\n'
        "    return 0;\n}\n")
    exe = tmp_path / "probe"
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-leffimvs", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True)
    assert int(out.stdout.strip()) > 100 << 20      # the DTU-shape bf16x3 workspace is a few hundred MB


def test_documents_name_only_declared_entry_points():
    """every effimvs_* function that INTEGRATION.md, DESIGN.md or README.md mentions is declared in include/effimvs.h"""
    declared = set(declared_functions())
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for name in set(re.findall(r"\b(effimvs_[a-z0-9_]*[a-z0-9])\b", text)):
            if name in ("effimvs_b200",) or name.startswith("effimvs_b200"):
                continue
            # shorthand like `effimvs_*_ex` / `effimvs_*_workspace_bytes` is written with a star and never matches; `_conv_f32` style suffixes neither
            assert name in declared or any(d.startswith(name) for d in declared), "{} names {} which the header does not declare".format(doc, name)


def test_conv2d_entry_points_validate_on_the_host():
    """effimvs_conv2d_tf32*: shape support, packed size and argument checks are host-side (nothing is launched)"""
    lib = capi.lib
    sup = lib.effimvs_conv2d_tf32_supported
    # the layers of the three update blocks (hidden 16 / 32 / 48): 2h -> 2h, 2h -> h, h -> h, h -> 2h
    for h in (16, 32, 48):
        for cin, cout in ((2 * h, 2 * h), (2 * h, h), (h, h), (h, 2 * h)):
            assert sup(cin, cout) == 1, (cin, cout)
            assert lib.effimvs_conv2d_tf32_packed_bytes(cin, cout) == 18 * cin * ((cout + 15) // 16 * 16)
    assert sup(12, 16) == 0 and sup(16, 6) == 0 and sup(256, 256) == 0 and sup(16, 12) == 1
    assert lib.effimvs_conv2d_tf32_packed_bytes(0, 16) == 0
    P = ctypes.c_void_p
    a = P(1 << 20)                                     # 32-byte aligned dummy addresses; every call below fails before a launch
    mode = capi.CONV2D_BIAS
    assert lib.effimvs_conv2d_tf32(None, 32, 32, None, 0, 0, a, None, 32, 1, 8, 8, mode, a, 32, None, 0, None, 0, None) == capi.EINVAL
    assert lib.effimvs_conv2d_tf32(a, 32, 32, None, 0, 0, a, None, 32, 1, 1, 8, mode, a, 32, None, 0, None, 0, None) == capi.EINVAL      # H < 2
    assert lib.effimvs_conv2d_tf32(a, 32, 12, None, 0, 0, a, None, 32, 1, 8, 8, mode, a, 32, None, 0, None, 0, None) == capi.EINVAL      # segment % 8
    assert lib.effimvs_conv2d_tf32(a, 32, 24, None, 0, 0, a, None, 32, 1, 8, 8, mode, a, 32, None, 0, None, 0, None) == capi.EUNSUPPORTED  # cin % 16
    assert lib.effimvs_conv2d_tf32(P((1 << 20) + 16), 32, 32, None, 0, 0, a, None, 32, 1, 8, 8, mode, a, 32, None, 0, None, 0, None) == capi.EINVAL
    assert "32-byte aligned" in capi.last_error()
    assert lib.effimvs_conv2d_tf32(a, 32, 32, None, 0, 0, a, None, 32, 1, 8, 8, 9, a, 32, None, 0, None, 0, None) == capi.EINVAL        # mode
    assert lib.effimvs_conv2d_tf32(a, 32, 32, None, 0, 0, a, None, 32, 1, 8, 8, capi.CONV2D_ADD_RELU, a, 32, None, 0, None, 0, None) == capi.EINVAL
    assert "addend" in capi.last_error()
    assert lib.effimvs_conv2d_tf32(a, 32, 32, None, 0, 0, a, None, 32, 1, 8, 8, capi.CONV2D_GRU_GATES, a, 16, a, 16, None, 0, None) == capi.EINVAL
    assert lib.effimvs_conv2d_tf32_pack(None, 16, 16, a, None) == capi.EINVAL
    assert lib.effimvs_conv2d_tf32_pack(a, 12, 16, a, None) == capi.EUNSUPPORTED
    with pytest.raises(RuntimeError):
        ops.conv2d_tc(torch.zeros(1, 16, 4, 4), None, torch.zeros(10), None, 16, mode, torch.zeros(1, 16, 4, 4), None, None)
