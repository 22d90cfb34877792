// Net-level entry points of the 3-D regularization: the 9-layer cost-regularization FPN and the
// 4-layer cross-scale propagation net, sequenced on the caller's stream with a caller-provided
// workspace (no allocation, no synchronisation).
//
//   CostRegNet_2_sample_FPN3D_Fast.forward   upstream models/module.py:453-463
//   cost_up_small.forward                    upstream models/module.py:509-516
#include "common.cuh"

namespace effimvs {

int conv3d_f32(const float* x, const float* weight, const float* bias, const float* residual, int B, int Cin, int Cout,
               int D, int H, int W, int sd, int sh, int sw, int transposed, int relu, float* y, int y_coff, int y_ctot,
               cudaStream_t st);

// bf16 tcgen05 implementations (conv3d_tc.cu)
size_t costreg_bf16_workspace_bytes(int B, int D, int H, int W, bool hilo);
int costreg_bf16(const float* x, const float* const* weights, const float* const* biases, int B, int D, int H, int W,
                 bool hilo, int phases, void* ws, size_t ws_bytes, float* prob_out, cudaStream_t st);
size_t cost_up_bf16_workspace_bytes(int B, int D, int H, int W, bool hilo);
int cost_up_bf16(const float* x, const float* prev, const float* const* weights, const float* const* biases, int B,
                 int D, int H, int W, bool hilo, int phases, void* ws, size_t ws_bytes, float* out, cudaStream_t st);

namespace {
size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
}  // namespace

}  // namespace effimvs

using namespace effimvs;

extern "C" size_t effimvs_costreg_workspace_bytes(int B, int D, int H, int W, int precision) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    if (precision != EFFIMVS_PREC_F32) return costreg_bf16_workspace_bytes(B, D, H, W, precision == EFFIMVS_PREC_BF16X3);
    size_t v = (size_t)B * D * H * W;
    // c0, c1 (8 ch, full) + c2, c3, c6 (16 ch, 1/8 of the voxels) + c4, c5 (32 ch, 1/64); c7 reuses c0
    return align256(8 * v * 4) * 2 + align256(16 * (v / 8) * 4) * 3 + align256(32 * (v / 64) * 4) * 2;
}

extern "C" int effimvs_costreg_fpn3d(const float* x, const float* const* weights, const float* const* biases,
                                     int B, int D, int H, int W, int precision, void* workspace, size_t workspace_bytes,
                                     float* prob_out, void* stream) {
    return effimvs_costreg_fpn3d_ex(x, weights, biases, B, D, H, W, precision, EFFIMVS_WS_PREPARE | EFFIMVS_WS_RUN, workspace,
                                    workspace_bytes, prob_out, stream);
}

extern "C" int effimvs_costreg_fpn3d_ex(const float* x, const float* const* weights, const float* const* biases,
                                        int B, int D, int H, int W, int precision, int phases, void* workspace,
                                        size_t workspace_bytes, float* prob_out, void* stream) {
    EFFI_REQUIRE(phases > 0 && !(phases & ~(EFFIMVS_WS_PREPARE | EFFIMVS_WS_RUN)), EFFIMVS_EINVAL, "costreg_fpn3d: phases=%d", phases);
    const bool run = (phases & EFFIMVS_WS_RUN) != 0;
    EFFI_REQUIRE(weights && biases && workspace && (!run || (x && prob_out)), EFFIMVS_EINVAL, "costreg_fpn3d: null pointer");
    EFFI_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, EFFIMVS_EINVAL, "costreg_fpn3d: bad sizes");
    EFFI_REQUIRE(D % 4 == 0 && H % 4 == 0 && W % 4 == 0, EFFIMVS_EUNSUPPORTED,
                 "costreg_fpn3d: D,H,W = %d,%d,%d must be multiples of 4 (two stride-2 levels)", D, H, W);
    for (int i = 0; i < 9; ++i) EFFI_REQUIRE(weights[i], EFFIMVS_EINVAL, "costreg_fpn3d: weights[%d] is NULL", i);
    size_t need = effimvs_costreg_workspace_bytes(B, D, H, W, precision);
    EFFI_REQUIRE(workspace_bytes >= need, EFFIMVS_EWORKSPACE, "costreg_fpn3d: workspace %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == EFFIMVS_PREC_BF16 || precision == EFFIMVS_PREC_BF16X3)
        return costreg_bf16(x, weights, biases, B, D, H, W, precision == EFFIMVS_PREC_BF16X3, phases, workspace, workspace_bytes, prob_out, st);
    EFFI_REQUIRE(precision == EFFIMVS_PREC_F32, EFFIMVS_EINVAL, "costreg_fpn3d: precision=%d", precision);
    if (!run) return EFFIMVS_OK;   // nothing to prepare for the fp32 layers

    size_t v = (size_t)B * D * H * W;
    char* p = (char*)workspace;
    float* c0 = (float*)p; p += align256(8 * v * 4);
    float* c1 = (float*)p; p += align256(8 * v * 4);
    float* c2 = (float*)p; p += align256(16 * (v / 8) * 4);
    float* c3 = (float*)p; p += align256(16 * (v / 8) * 4);
    float* c6 = (float*)p; p += align256(16 * (v / 8) * 4);
    float* c4 = (float*)p; p += align256(32 * (v / 64) * 4);
    float* c5 = (float*)p;
    float* c7 = c0;
    int rc;
    const int D2 = D / 2, H2 = H / 2, W2 = W / 2, D4 = D / 4, H4 = H / 4, W4 = W / 4;
    if ((rc = conv3d_f32(x, weights[0], biases[0], nullptr, B, 1, 8, D, H, W, 1, 1, 1, 0, 1, c0, 0, 8, st))) return rc;
    if ((rc = conv3d_f32(c0, weights[1], biases[1], nullptr, B, 8, 8, D, H, W, 1, 1, 1, 0, 1, c1, 0, 8, st))) return rc;
    if ((rc = conv3d_f32(c1, weights[2], biases[2], nullptr, B, 8, 16, D, H, W, 2, 2, 2, 0, 1, c2, 0, 16, st))) return rc;
    if ((rc = conv3d_f32(c2, weights[3], biases[3], nullptr, B, 16, 16, D2, H2, W2, 1, 1, 1, 0, 1, c3, 0, 16, st))) return rc;
    if ((rc = conv3d_f32(c3, weights[4], biases[4], nullptr, B, 16, 32, D2, H2, W2, 2, 2, 2, 0, 1, c4, 0, 32, st))) return rc;
    if ((rc = conv3d_f32(c4, weights[5], biases[5], nullptr, B, 32, 32, D4, H4, W4, 1, 1, 1, 0, 1, c5, 0, 32, st))) return rc;
    if ((rc = conv3d_f32(c5, weights[6], biases[6], c3, B, 32, 16, D4, H4, W4, 2, 2, 2, 1, 1, c6, 0, 16, st))) return rc;
    if ((rc = conv3d_f32(c6, weights[7], biases[7], c1, B, 16, 8, D2, H2, W2, 2, 2, 2, 1, 1, c7, 0, 8, st))) return rc;
    return conv3d_f32(c7, weights[8], nullptr, nullptr, B, 8, 1, D, H, W, 1, 1, 1, 0, 0, prob_out, 0, 1, st);
}

extern "C" size_t effimvs_cost_up_workspace_bytes(int B, int D, int H, int W, int precision) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    if (precision != EFFIMVS_PREC_F32) return cost_up_bf16_workspace_bytes(B, D, H, W, precision == EFFIMVS_PREC_BF16X3);
    size_t v = (size_t)B * D * (H / 2) * (W / 2);
    return align256(16 * v * 4) + align256(8 * v * 4);
}

extern "C" int effimvs_cost_up_small(const float* x, const float* prev, const float* const* weights,
                                     const float* const* biases, int B, int D, int H, int W, int precision,
                                     void* workspace, size_t workspace_bytes, float* out, void* stream) {
    return effimvs_cost_up_small_ex(x, prev, weights, biases, B, D, H, W, precision, EFFIMVS_WS_PREPARE | EFFIMVS_WS_RUN, workspace,
                                    workspace_bytes, out, stream);
}

extern "C" int effimvs_cost_up_small_ex(const float* x, const float* prev, const float* const* weights,
                                        const float* const* biases, int B, int D, int H, int W, int precision, int phases,
                                        void* workspace, size_t workspace_bytes, float* out, void* stream) {
    EFFI_REQUIRE(phases > 0 && !(phases & ~(EFFIMVS_WS_PREPARE | EFFIMVS_WS_RUN)), EFFIMVS_EINVAL, "cost_up_small: phases=%d", phases);
    const bool run = (phases & EFFIMVS_WS_RUN) != 0;
    EFFI_REQUIRE(weights && biases && workspace && (!run || (x && prev && out)), EFFIMVS_EINVAL, "cost_up_small: null pointer");
    EFFI_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, EFFIMVS_EINVAL, "cost_up_small: bad sizes");
    EFFI_REQUIRE(H % 2 == 0 && W % 2 == 0, EFFIMVS_EUNSUPPORTED, "cost_up_small: H,W = %d,%d must be even", H, W);
    for (int i = 0; i < 4; ++i)
        EFFI_REQUIRE(weights[i] && biases[i], EFFIMVS_EINVAL, "cost_up_small: weights/biases[%d] is NULL", i);
    size_t need = effimvs_cost_up_workspace_bytes(B, D, H, W, precision);
    EFFI_REQUIRE(workspace_bytes >= need, EFFIMVS_EWORKSPACE, "cost_up_small: workspace %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == EFFIMVS_PREC_BF16 || precision == EFFIMVS_PREC_BF16X3)
        return cost_up_bf16(x, prev, weights, biases, B, D, H, W, precision == EFFIMVS_PREC_BF16X3, phases, workspace, workspace_bytes, out, st);
    EFFI_REQUIRE(precision == EFFIMVS_PREC_F32, EFFIMVS_EINVAL, "cost_up_small: precision=%d", precision);
    if (!run) return EFFIMVS_OK;

    const int H2 = H / 2, W2 = W / 2;
    size_t v = (size_t)B * D * H2 * W2;
    float* cat = (float*)workspace;
    float* c1 = (float*)((char*)workspace + align256(16 * v * 4));
    int rc;
    if ((rc = conv3d_f32(x, weights[0], biases[0], nullptr, B, 1, 8, D, H, W, 1, 2, 2, 0, 1, cat, 0, 16, st))) return rc;
    if ((rc = conv3d_f32(prev, weights[1], biases[1], nullptr, B, 1, 8, D, H2, W2, 1, 1, 1, 0, 1, cat, 8, 16, st))) return rc;
    if ((rc = conv3d_f32(cat, weights[2], biases[2], nullptr, B, 16, 8, D, H2, W2, 1, 1, 1, 0, 1, c1, 0, 8, st))) return rc;
    return conv3d_f32(c1, weights[3], biases[3], nullptr, B, 8, 1, D, H2, W2, 1, 2, 2, 1, 1, out, 0, 1, st);
}
