// Per-pixel glue of the ConvGRU update block and the convex upsampling: the
// 2-D convolutions stay cuDNN (stock PyTorch); everything between them -- gate activations, the
// reset product written straight into the next convolution's input, the state update, the inverse-
// depth step with its depth conversion, and the softmax-weighted 3x3 upsampling -- is one pass each
// instead of a chain of a dozen elementwise launches over the same maps.
//
//   ConvGRU.forward            upstream models/update.py:33-49
//   DepthHead.forward (tail)   upstream models/update.py:19-27
//   BasicUpdateBlock.forward   upstream models/update.py:114-127 (inv_depth + delta)
//   disp_to_depth              upstream models/Effi_MVS_plus.py:138-148
//   upsample_depth             upstream models/Effi_MVS_plus.py:167-178
//
// Multi-channel maps are channels-last (B,H,W,C), what cuDNN's NHWC convolutions emit; the arithmetic
// repeats torch's elementwise kernels operation by operation (no contraction across torch ops):
// sigmoid = 1 / (1 + exp(-x)), tanh = tanhf, reciprocal = 1 / x, bias added first as cuDNN's epilogue does.
#include <stdlib.h>

#include "common.cuh"

namespace effimvs {
namespace {

__device__ __forceinline__ float sigmoid_t(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// disp_to_depth: 1 / clamp(lo + (hi - lo) * inv, 1e-4)
__device__ __forceinline__ float to_depth(float inv, float lo, float hi) {
    const float s = fmaxf(__fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), inv)), 1e-4f);
    return __fdiv_rn(1.0f, s);
}

// rhx = [ sigmoid(r_pre + b_r) * h ; x ]   with zr_pre = (pixels, 2h) = [z_pre ; r_pre], hx = (pixels, h + cx) = [h ; x]
__global__ void __launch_bounds__(256)
gru_reset_kernel(const float4* __restrict__ zr_pre, const float* __restrict__ bias_r, const float4* __restrict__ hx, long long n_pix, int h,
                 int cx, float4* __restrict__ rhx) {
    pdl_enter();
    const int ct4 = (h + cx) >> 2, h4 = h >> 2;
    const long long total = n_pix * ct4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / ct4;
        const int c4 = (int)(i - p * ct4);
        float4 v = __ldg(hx + i);
        if (c4 < h4) {
            const float4 r = __ldg(zr_pre + p * (2 * h4) + h4 + c4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias_r) + c4);
            v.x = __fmul_rn(sigmoid_t(__fadd_rn(r.x, b.x)), v.x);
            v.y = __fmul_rn(sigmoid_t(__fadd_rn(r.y, b.y)), v.y);
            v.z = __fmul_rn(sigmoid_t(__fadd_rn(r.z, b.z)), v.z);
            v.w = __fmul_rn(sigmoid_t(__fadd_rn(r.w, b.w)), v.w);
        }
        rhx[i] = v;
    }
}

__device__ __forceinline__ float gru_mix(float zp, float bz, float qp, float bq, float hv) {
    const float z = sigmoid_t(__fadd_rn(zp, bz));
    const float q = tanhf(__fadd_rn(qp, bq));
    return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, z), hv), __fmul_rn(z, q));   // (1 - z) * h + z * q
}

// h' = (1 - z) * h + z * tanh(q_pre + b_q), z = sigmoid(z_pre + b_z); written to hx[:, :h] in place and to net (pixels, h)
__global__ void __launch_bounds__(256)
gru_update_kernel(const float4* __restrict__ zr_pre, const float* __restrict__ bias_z, const float4* __restrict__ q_pre,
                  const float* __restrict__ bias_q, float4* __restrict__ hx, long long n_pix, int h, int cx, float4* __restrict__ net) {
    pdl_enter();
    const int ct4 = (h + cx) >> 2, h4 = h >> 2;
    const long long total = n_pix * h4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / h4;
        const int c4 = (int)(i - p * h4);
        const float4 zp = __ldg(zr_pre + p * (2 * h4) + c4), qp = __ldg(q_pre + i);
        const float4 bz = __ldg(reinterpret_cast<const float4*>(bias_z) + c4), bq = __ldg(reinterpret_cast<const float4*>(bias_q) + c4);
        const float4 hv = hx[p * ct4 + c4];
        float4 o;
        o.x = gru_mix(zp.x, bz.x, qp.x, bq.x, hv.x);
        o.y = gru_mix(zp.y, bz.y, qp.y, bq.y, hv.y);
        o.z = gru_mix(zp.z, bz.z, qp.z, bq.z, hv.z);
        o.w = gru_mix(zp.w, bz.w, qp.w, bq.w, hv.w);
        hx[p * ct4 + c4] = o;
        net[i] = o;
    }
}

// inv' = inv + tanh(pre + b)  (pre == nullptr: inv' = inv), depth = disp_to_depth(inv')
__global__ void __launch_bounds__(256)
gru_delta_kernel(const float* __restrict__ pre, const float* __restrict__ bias, const float* __restrict__ inv, const float* __restrict__ lo,
                 const float* __restrict__ hi, int HW, float* __restrict__ inv_out, float* __restrict__ depth_out) {
    pdl_enter();
    const int b = blockIdx.y;
    const float l = __ldg(lo + b), hh = __ldg(hi + b);
    const float bv = pre ? __ldg(bias) : 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const size_t o = (size_t)b * HW + i;
        float v = __ldg(inv + o);
        if (pre) v = __fadd_rn(v, tanhf(__fadd_rn(__ldg(pre + o), bv)));
        if (inv_out) inv_out[o] = v;
        depth_out[o] = to_depth(v, l, hh);
    }
}

// Start of a stage's refinement (models/Effi_MVS_plus.py:151-164 depth_to_disp, then :138-148 disp_to_depth): the normalised
// inverse depth of the current estimate, inv = (1 / depth - lo) / ((hi - lo) + 1e-10), and the depth it maps back to -- the same
// IEEE operations torch issues as four kernels (reciprocal, sub, div, and the clone + disp_to_depth of the first gru_delta call).
__global__ void __launch_bounds__(256)
inv_init_kernel(const float* __restrict__ cur_depth, const float* __restrict__ lo, const float* __restrict__ hi, int HW,
                float* __restrict__ inv_out, float* __restrict__ depth_out) {
    pdl_enter();
    const int b = blockIdx.y;
    const float l = __ldg(lo + b), hh = __ldg(hi + b);
    const float span = __fadd_rn(__fsub_rn(hh, l), 1e-10f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const size_t o = (size_t)b * HW + i;
        const float v = __fdiv_rn(__fsub_rn(__frcp_rn(__ldg(cur_depth + o)), l), span);
        inv_out[o] = v;
        depth_out[o] = to_depth(v, l, hh);
    }
}

// upsample_depth with ratio R (2): mask (B,H,W,9*R*R) = scale * (mask_pre + bias), channel = (k*R + ry)*R + rx;
// softmax over the 9 neighbours k, weighted sum of the zero-padded 3x3 neighbourhood of inv.
// CONV: mask_pre is not given; the kernel forms it from t (B,H,W,K) = relu(mask[0](net)) and mask[2].weight (CH, K) -- a
// per-pixel K x CH product against weights held in shared memory -- so the CH-channel mask map is never written or read.
template <int R, bool CONV>
__global__ void __launch_bounds__(128)
convex_upsample_kernel(const float* __restrict__ mask_pre, const float* __restrict__ mask_w, int K, const float* __restrict__ mask_bias,
                       float scale, const float* __restrict__ inv, const float* __restrict__ lo, const float* __restrict__ hi, int H, int W,
                       float* __restrict__ up_out, float* __restrict__ depth_out) {
    constexpr int CH = 9 * R * R;
    __shared__ float sb[CH];
    extern __shared__ __align__(16) float s_mw[];        // CONV: [K][CH]
    pdl_trigger();
    for (int i = threadIdx.x; i < CH; i += blockDim.x) sb[i] = mask_bias ? mask_bias[i] : 0.0f;
    if (CONV)
        for (int i = threadIdx.x; i < K * CH; i += blockDim.x) s_mw[i] = mask_w[(i % CH) * K + i / CH];
    pdl_wait();
    __syncthreads();
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    const float l = __ldg(lo + b), hh = __ldg(hi + b);
    const float* ib = inv + (size_t)b * H * W;
    float nb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        nb[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(ib + (size_t)yy * W + xx) : 0.0f;
    }
    float m[CH];
    if (CONV) {
        const float4* tp = reinterpret_cast<const float4*>(mask_pre + ((size_t)b * H * W + pix) * K);
#pragma unroll
        for (int c = 0; c < CH; ++c) m[c] = 0.0f;
        for (int k4 = 0; k4 < (K >> 2); ++k4) {
            const float4 v = __ldg(tp + k4);
            const float tv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4* wr = reinterpret_cast<const float4*>(s_mw + (4 * k4 + j) * CH);
#pragma unroll
                for (int q = 0; q < CH / 4; ++q) {
                    const float4 wv = wr[q];
                    m[4 * q] = fmaf(tv[j], wv.x, m[4 * q]); m[4 * q + 1] = fmaf(tv[j], wv.y, m[4 * q + 1]);
                    m[4 * q + 2] = fmaf(tv[j], wv.z, m[4 * q + 2]); m[4 * q + 3] = fmaf(tv[j], wv.w, m[4 * q + 3]);
                }
            }
        }
    } else {
        const float4* mp = reinterpret_cast<const float4*>(mask_pre + ((size_t)b * H * W + pix) * CH);
#pragma unroll
        for (int q = 0; q < CH / 4; ++q) {
            const float4 v = __ldg(mp + q);
            m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
        }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) m[c] = __fmul_rn(scale, __fadd_rn(m[c], sb[c]));
    const int WO = W * R;
#pragma unroll
    for (int ry = 0; ry < R; ++ry) {
        float res[R];
#pragma unroll
        for (int rx = 0; rx < R; ++rx) {
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 9; ++k) mx = fmaxf(mx, m[(k * R + ry) * R + rx]);
            float e[9], z = 0.0f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                e[k] = expf(__fsub_rn(m[(k * R + ry) * R + rx], mx));
                z = __fadd_rn(z, e[k]);
            }
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 9; ++k) acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(e[k], z), nb[k]));
            res[rx] = acc;
        }
        const size_t o = ((size_t)b * H * R + (size_t)y * R + ry) * WO + (size_t)x * R;
#pragma unroll
        for (int rx = 0; rx < R; ++rx) {
            if (up_out) up_out[o + rx] = res[rx];
            if (depth_out) depth_out[o + rx] = to_depth(res[rx], l, hh);
        }
    }
}

int grid_for(long long work, int block) {
    long long g = (work + block - 1) / block;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace
}  // namespace effimvs

using namespace effimvs;

extern "C" int effimvs_gru_reset_f32(const float* zr_pre, const float* bias_r, const float* hx, long long n_pix, int h, int cx,
                                     float* rhx, void* stream) {
    EFFI_REQUIRE(zr_pre && bias_r && hx && rhx, EFFIMVS_EINVAL, "gru_reset: null pointer");
    EFFI_REQUIRE(n_pix > 0 && h > 0 && cx >= 0 && h % 4 == 0 && cx % 4 == 0, EFFIMVS_EINVAL, "gru_reset: h=%d, cx=%d must be multiples of 4", h, cx);
    const long long work = n_pix * ((h + cx) / 4);
    launch_kernel(gru_reset_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)zr_pre, bias_r, (const float4*)hx, n_pix, h, cx,
                                                                           (float4*)rhx);
    return check_launch("gru_reset_kernel");
}

extern "C" int effimvs_gru_update_f32(const float* zr_pre, const float* bias_z, const float* q_pre, const float* bias_q, float* hx,
                                      long long n_pix, int h, int cx, float* net_out, void* stream) {
    EFFI_REQUIRE(zr_pre && bias_z && q_pre && bias_q && hx && net_out, EFFIMVS_EINVAL, "gru_update: null pointer");
    EFFI_REQUIRE(n_pix > 0 && h > 0 && cx >= 0 && h % 4 == 0 && cx % 4 == 0, EFFIMVS_EINVAL, "gru_update: h=%d, cx=%d must be multiples of 4", h, cx);
    const long long work = n_pix * (h / 4);
    launch_kernel(gru_update_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)zr_pre, bias_z, (const float4*)q_pre, bias_q,
                                                                            (float4*)hx, n_pix, h, cx, (float4*)net_out);
    return check_launch("gru_update_kernel");
}

extern "C" int effimvs_inv_init_f32(const float* cur_depth, const float* lo_disp, const float* hi_disp, int B, int HW, float* inv_out,
                                    float* depth_out, void* stream) {
    EFFI_REQUIRE(cur_depth && lo_disp && hi_disp && inv_out && depth_out, EFFIMVS_EINVAL, "inv_init: null pointer");
    EFFI_REQUIRE(B > 0 && B <= 65535 && HW > 0, EFFIMVS_EINVAL, "inv_init: bad sizes");
    dim3 grid(grid_for(HW, 256), B);
    launch_kernel(inv_init_kernel, grid, dim3(256), 0, (cudaStream_t)stream, cur_depth, lo_disp, hi_disp, HW, inv_out, depth_out);
    return check_launch("inv_init_kernel");
}

extern "C" int effimvs_gru_delta_f32(const float* pre, const float* bias, const float* inv, const float* lo_disp, const float* hi_disp,
                                     int B, int HW, float* inv_out, float* depth_out, void* stream) {
    EFFI_REQUIRE(inv && lo_disp && hi_disp && depth_out && (!pre || bias), EFFIMVS_EINVAL, "gru_delta: null pointer");
    EFFI_REQUIRE(B > 0 && B <= 65535 && HW > 0, EFFIMVS_EINVAL, "gru_delta: bad sizes");
    dim3 grid(grid_for(HW, 256), B);
    launch_kernel(gru_delta_kernel, grid, dim3(256), 0, (cudaStream_t)stream, pre, bias, inv, lo_disp, hi_disp, HW, inv_out, depth_out);
    return check_launch("gru_delta_kernel");
}

// ------------------------------------------------------------------------------------------------
// DepthHead.conv2 (hidden -> 1, 3x3, zero padding; upstream models/update.py:19-27) fused with the step above:
// a one-channel convolution is 9 * h multiply-adds per pixel -- a stream over the hidden map, not a GEMM (one
// output channel would leave 15 of 16 accumulator columns of an MMA tile idle, and cuDNN pays an NHWC -> NCHW
// conversion kernel for the single-channel result).  Four lanes share a pixel column (lane q owns channel quads
// q, q+4, ...), a thread sweeps DH_ROWS output rows so that each input row is loaded once for the three rows it
// contributes to, the four partial sums are reduced with two shuffles and lane q finishes row q:
// inv' = inv + tanh(conv + bias), depth = disp_to_depth(inv').
// ------------------------------------------------------------------------------------------------
namespace effimvs {
namespace {

constexpr int DH_ROWS = 4, DH_COLS = 64;   // output tile of a 256-thread block: 64 columns x 4 rows

template <int J>   // hidden channels / 16
__global__ void __launch_bounds__(256)
delta_head_kernel(const float4* __restrict__ t, const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ inv,
                  const float* __restrict__ lo, const float* __restrict__ hi, int H, int W, float* __restrict__ inv_out,
                  float* __restrict__ depth_out) {
    constexpr int h = 16 * J, h4 = 4 * J;
    __shared__ float4 s_w[9 * h4];                       // [tap][channel quad]
    float* s_wf = reinterpret_cast<float*>(s_w);
    pdl_trigger();
    for (int i = threadIdx.x; i < 9 * h; i += blockDim.x) s_wf[i] = w[(i % h) * 9 + i / h];   // weight (1, h, 3, 3)
    pdl_wait();
    __syncthreads();
    const int q = threadIdx.x & 3;
    const int px = blockIdx.x * DH_COLS + (threadIdx.x >> 2), y0 = blockIdx.y * DH_ROWS, b = blockIdx.z;
    // the kernel is bound by instruction issue (1.9 M threads of ~400 instructions at 800 x 592): the products run on the packed
    // fp32 pipe (two channels per FFMA2, the two halves added at the end).  Measured and dropped: with h = 16 the nine weight quads
    // of a thread in registers instead of shared memory (70 registers, three 256-thread blocks per SM: 26.2 against 23.8 us).
    ulonglong2 acc2[DH_ROWS];
#pragma unroll
    for (int o = 0; o < DH_ROWS; ++o) acc2[o] = make_ulonglong2(0ull, 0ull);
    if (px < W) {
#pragma unroll
        for (int r = -1; r <= DH_ROWS; ++r) {
            const int yy = y0 + r;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = px + kx - 1;
                if (xx < 0 || xx >= W) continue;
                const ulonglong2* base = reinterpret_cast<const ulonglong2*>(t + (((size_t)b * H + yy) * W + xx) * h4 + q);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const ulonglong2 v = __ldg(base + 4 * j);
#pragma unroll
                    for (int o = 0; o < DH_ROWS; ++o) {
                        const int ky = r - o + 1;
                        if (ky < 0 || ky > 2) continue;
                        const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(&s_w[(ky * 3 + kx) * h4 + 4 * j + q]);
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[o].x) : "l"(v.x), "l"(wv.x));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[o].y) : "l"(v.y), "l"(wv.y));
                    }
                }
            }
        }
    }
    float mine = 0.0f;
#pragma unroll
    for (int o = 0; o < DH_ROWS; ++o) {
        float a0, a1, a2, a3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc2[o].x));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(acc2[o].y));
        float a = (a0 + a1) + (a2 + a3);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (o == q) mine = a;
    }
    const int y = y0 + q;
    if (px >= W || y >= H) return;
    const size_t o = ((size_t)b * H + y) * W + px;
    const float v = __fadd_rn(__ldg(inv + o), tanhf(__fadd_rn(mine, __ldg(bias))));
    inv_out[o] = v;
    depth_out[o] = to_depth(v, __ldg(lo + b), __ldg(hi + b));
}

}  // namespace
}  // namespace effimvs

extern "C" int effimvs_delta_head_f32(const float* t, const float* weight, const float* bias, const float* inv, const float* lo_disp,
                                      const float* hi_disp, int B, int h, int H, int W, float* inv_out, float* depth_out, void* stream) {
    using namespace effimvs;
    EFFI_REQUIRE(t && weight && bias && inv && lo_disp && hi_disp && inv_out && depth_out, EFFIMVS_EINVAL, "delta_head: null pointer");
    EFFI_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, EFFIMVS_EINVAL, "delta_head: bad sizes");
    dim3 grid(ceil_div(W, DH_COLS), ceil_div(H, DH_ROWS), B);
    EFFI_REQUIRE(grid.y <= 65535, EFFIMVS_EUNSUPPORTED, "delta_head: H=%d too large", H);
    cudaStream_t st = (cudaStream_t)stream;
#define EFFI_DH_CASE(JJ)                                                                                                     \
    case 16 * JJ:                                                                                                            \
        launch_kernel(delta_head_kernel<JJ>, grid, dim3(256), 0, st, (const float4*)t, weight, bias, inv, lo_disp, hi_disp, H, W, inv_out, depth_out); \
        break;
    switch (h) {
        EFFI_DH_CASE(1) EFFI_DH_CASE(2) EFFI_DH_CASE(3) EFFI_DH_CASE(4) EFFI_DH_CASE(6) EFFI_DH_CASE(8)
        default:
            set_error("delta_head: hidden channels %d not in {16,32,48,64,96,128}", h);
            return EFFIMVS_EUNSUPPORTED;
    }
#undef EFFI_DH_CASE
    return check_launch("delta_head_kernel");
}

extern "C" int effimvs_convex_upsample_f32(const float* mask_pre, const float* mask_bias, float mask_scale, const float* inv,
                                           const float* lo_disp, const float* hi_disp, int B, int H, int W, int ratio, float* up_out,
                                           float* depth_out, void* stream) {
    EFFI_REQUIRE(mask_pre && inv && lo_disp && hi_disp && (up_out || depth_out), EFFIMVS_EINVAL, "convex_upsample: null pointer");
    EFFI_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, EFFIMVS_EINVAL, "convex_upsample: bad sizes");
    EFFI_REQUIRE(ratio == 2, EFFIMVS_EUNSUPPORTED, "convex_upsample: ratio=%d (only 2, the value upstream uses, is built)", ratio);
    dim3 grid(ceil_div(H * W, 128), B);
    launch_kernel(convex_upsample_kernel<2, false>, grid, dim3(128), 0, (cudaStream_t)stream, mask_pre, nullptr, 0, mask_bias, mask_scale, inv, lo_disp, hi_disp,
                                                                            H, W, up_out, depth_out);
    return check_launch("convex_upsample_kernel");
}

extern "C" int effimvs_convex_upsample_conv_f32(const float* t, int K, const float* mask_w, const float* mask_bias, float mask_scale,
                                                const float* inv, const float* lo_disp, const float* hi_disp, int B, int H, int W, int ratio,
                                                float* up_out, float* depth_out, void* stream) {
    EFFI_REQUIRE(t && mask_w && inv && lo_disp && hi_disp && (up_out || depth_out), EFFIMVS_EINVAL, "convex_upsample_conv: null pointer");
    EFFI_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, EFFIMVS_EINVAL, "convex_upsample_conv: bad sizes");
    EFFI_REQUIRE(ratio == 2, EFFIMVS_EUNSUPPORTED, "convex_upsample_conv: ratio=%d (only 2, the value upstream uses, is built)", ratio);
    EFFI_REQUIRE(K >= 4 && K % 4 == 0 && K <= 256, EFFIMVS_EUNSUPPORTED, "convex_upsample_conv: K=%d must be a multiple of 4 up to 256", K);
    dim3 grid(ceil_div(H * W, 128), B);
    const size_t smem = (size_t)K * 36 * sizeof(float);
    launch_kernel(convex_upsample_kernel<2, true>, grid, dim3(128), smem, (cudaStream_t)stream, t, mask_w, K, mask_bias, mask_scale, inv, lo_disp, hi_disp, H, W,
                                                                              up_out, depth_out);
    return check_launch("convex_upsample_kernel");
}

// ------------------------------------------------------------------------------------------------
// ProjectionInput head (upstream models/update.py:88-91): the two input-side convolutions of the cost
// encoder, relu(convc1(cost)) (1x1, CD -> h) and relu(convd1(inv)) (7x7 pad 3, 1 -> h), written side by
// side into one channels-last (B,H,W,2h) map -- the input of the (block-diagonal) second layer, so the
// concatenation upstream performs later (update.py:92) never happens.  With 6 and 1 input channels these
// layers are far too thin for cuDNN's implicit GEMMs; here a thread owns one pixel, the 7x7 window comes
// from a shared-memory tile and the weights are broadcast 128-bit shared loads, 16 output channels at a time.
// ------------------------------------------------------------------------------------------------
namespace effimvs {
namespace {

constexpr int EH_TX = 32, EH_R = 3, EH_PX = 2;   // a thread owns EH_PX horizontally adjacent pixels: weights loaded once for both

// EH_TY: rows of a block's tile.  4 by default: 64 x 8 tiles leave SMs idle on the small stages (200 x 148 is 76 such tiles)
// and at two 256-thread blocks per SM (126 registers) the tail wave of the large one costs more than the extra halo reads.
// HT: the hidden size as a compile-time constant (16 / 32; 0 = run-time h), so that the 49 x h / 4 weight reads of the unrolled window
// loop are shared-memory loads at immediate offsets; with h = 16 the channel loop is a single pass and the kernel needs 72 registers
// instead of 127 (stage 3: 57 -> 47 us).  Measured no gain for 32 (neither as a constant nor with its two chunks spread over
// threadIdx.z: 31.7 / 34.8 us); 48 takes the z form (23.5 -> 20.4 us).
// ZC: the 16-channel chunks of the hidden size are spread over threadIdx.z (block = 32 x EH_TY x h / 16 threads) instead of being a loop
// of one thread: the same 72 registers for every h, h / 16 times the warps to hide the broadcast-load latency.
template <int EH_TY, int HT, bool ZC>
__global__ void __launch_bounds__(EH_TX * EH_TY * (ZC ? (HT ? HT / 16 : 1) : 1))
encoder_head_kernel(const float* __restrict__ cost, int CD, const float* __restrict__ inv, const float* __restrict__ wc1,
                    const float* __restrict__ bc1, const float* __restrict__ wd1, const float* __restrict__ bd1, int h_rt, int H, int W,
                    float* __restrict__ out) {
    const int h = HT ? HT : h_rt;
    extern __shared__ __align__(16) float esm[];
    float* s_wd = esm;                       // [49][h]   tap-major, channels contiguous
    float* s_wc = s_wd + 49 * h;             // [CD][h]
    float* s_b = s_wc + CD * h;              // [2h]      convc1 bias, convd1 bias
    float* s_inv = s_b + 2 * h;              // [EH_TY + 6][EH_TX * EH_PX + 6]
    constexpr int TWP = EH_TX * EH_PX, SW = TWP + 2 * EH_R, SH = EH_TY + 2 * EH_R;
    const int tid = (threadIdx.z * EH_TY + threadIdx.y) * EH_TX + threadIdx.x, nt = EH_TX * EH_TY * blockDim.z;
    const int b = blockIdx.z;
    pdl_trigger();       // the weight tables (constants) are staged while the predecessor drains; inv / cost after the wait
    for (int i = tid; i < 49 * h; i += nt) s_wd[i] = wd1[(i % h) * 49 + i / h];
    for (int i = tid; i < CD * h; i += nt) s_wc[i] = wc1[(i % h) * CD + i / h];
    for (int i = tid; i < 2 * h; i += nt) s_b[i] = i < h ? bc1[i] : bd1[i - h];
    pdl_wait();
    const int x0 = blockIdx.x * TWP - EH_R, y0 = blockIdx.y * EH_TY - EH_R;
    const float* ib = inv + (size_t)b * H * W;
    for (int i = tid; i < SW * SH; i += nt) {
        const int yy = y0 + i / SW, xx = x0 + i % SW;
        s_inv[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(ib + (size_t)yy * W + xx) : 0.0f;
    }
    __syncthreads();
    const int x = blockIdx.x * TWP + threadIdx.x * EH_PX, y = blockIdx.y * EH_TY + threadIdx.y;
    if (x >= W || y >= H) return;
    const bool has[EH_PX] = {true, x + 1 < W};
    float4* o[EH_PX];
#pragma unroll
    for (int p = 0; p < EH_PX; ++p) o[p] = reinterpret_cast<float4*>(out + (((size_t)b * H + y) * W + (has[p] ? x + p : x)) * (2 * h));
    // relu(convc1(cost) + b): 1x1
    float cv[EH_PX][8];
#pragma unroll
    for (int p = 0; p < EH_PX; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) cv[p][c] = (c < CD && has[p]) ? __ldg(cost + (((size_t)b * CD + c) * H + y) * W + x + p) : 0.0f;
    for (int ch = ZC ? threadIdx.z * 16 : 0; ch < (ZC ? threadIdx.z * 16 + 16 : h); ch += 16) {
        float acc[EH_PX][16];
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[p][j] = s_b[ch + j];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c < CD) {
                const float4* wp = reinterpret_cast<const float4*>(s_wc + c * h + ch);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = wp[q];
#pragma unroll
                    for (int p = 0; p < EH_PX; ++p) {
                        acc[p][4 * q] = fmaf(cv[p][c], wv.x, acc[p][4 * q]); acc[p][4 * q + 1] = fmaf(cv[p][c], wv.y, acc[p][4 * q + 1]);
                        acc[p][4 * q + 2] = fmaf(cv[p][c], wv.z, acc[p][4 * q + 2]); acc[p][4 * q + 3] = fmaf(cv[p][c], wv.w, acc[p][4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
            if (has[p])
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[p][(ch >> 2) + q] = make_float4(fmaxf(acc[p][4 * q], 0.0f), fmaxf(acc[p][4 * q + 1], 0.0f), fmaxf(acc[p][4 * q + 2], 0.0f),
                                                      fmaxf(acc[p][4 * q + 3], 0.0f));
    }
    // relu(convd1(inv) + b): 7x7, zero padding 3; the two pixels of a thread share a row of 8 window values
    const float* win = s_inv + threadIdx.y * SW + threadIdx.x * EH_PX;
    for (int ch = ZC ? threadIdx.z * 16 : 0; ch < (ZC ? threadIdx.z * 16 + 16 : h); ch += 16) {
        float acc[EH_PX][16];
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[p][j] = s_b[h + ch + j];
#pragma unroll
        for (int ty = 0; ty < 7; ++ty) {
            float v[7 + EH_PX - 1];
#pragma unroll
            for (int i = 0; i < 7 + EH_PX - 1; ++i) v[i] = win[ty * SW + i];
#pragma unroll
            for (int tx = 0; tx < 7; ++tx) {
                const float4* wp = reinterpret_cast<const float4*>(s_wd + (ty * 7 + tx) * h + ch);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = wp[q];
#pragma unroll
                    for (int p = 0; p < EH_PX; ++p) {
                        acc[p][4 * q] = fmaf(v[tx + p], wv.x, acc[p][4 * q]); acc[p][4 * q + 1] = fmaf(v[tx + p], wv.y, acc[p][4 * q + 1]);
                        acc[p][4 * q + 2] = fmaf(v[tx + p], wv.z, acc[p][4 * q + 2]); acc[p][4 * q + 3] = fmaf(v[tx + p], wv.w, acc[p][4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
            if (has[p])
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[p][((h + ch) >> 2) + q] = make_float4(fmaxf(acc[p][4 * q], 0.0f), fmaxf(acc[p][4 * q + 1], 0.0f), fmaxf(acc[p][4 * q + 2], 0.0f),
                                                            fmaxf(acc[p][4 * q + 3], 0.0f));
    }
}

// The same head with the weights as a KERNEL PARAMETER (constant bank): every FFMA of the unrolled 7x7 window takes its weight
// as a c[0][imm] operand, so there is no weight staging prologue per block (ncu on the shared-memory version at 800 x 592:
// ~30 % of the stall samples sit in the strided weight gather and the tile load before the first barrier) and no broadcast
// LDS.128 per four FMAs (short-scoreboard / MIO-throttle stalls of the main loop).  One launch covers 16 of the h output
// channels of each half (the table of a chunk is 3.7 KB of the 4 KB parameter space; offsets into it must be immediates, so
// the chunk cannot be a run-time index): h / 16 launches per call.  The host needs the weights in HOST memory at launch time
// (effimvs_encoder_head_pack_host builds the tables once per weight set); a CUDA graph keeps the copy made at capture.
struct EHTable {
    float wd[49][16];       // convd1.weight[ch0 + j][0][tap]   -> [tap][j]
    float wc[8][16];        // convc1.weight[ch0 + j][c]        -> [c][j], rows >= CD zero
    float bc[16], bd[16];   // convc1.bias, convd1.bias
};
static_assert(sizeof(EHTable) == 944 * sizeof(float), "EHTable layout is part of the C ABI (effimvs_encoder_head_pack_host)");

template <int EH_TY>
__global__ void __launch_bounds__(EH_TX * EH_TY)
encoder_head_const_kernel(const __grid_constant__ EHTable T, const float* __restrict__ cost, int CD, const float* __restrict__ inv,
                          int h, int ch0, int H, int W, float* __restrict__ out) {
    constexpr int TWP = EH_TX * EH_PX, SW = TWP + 2 * EH_R, SH = EH_TY + 2 * EH_R;
    static_assert(SW % 2 == 0 && EH_PX == 2, "window rows are read as float2 pairs");
    __shared__ __align__(16) float s_inv[SH * SW];
    pdl_enter();
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * TWP - EH_R, y0 = blockIdx.y * EH_TY - EH_R;
    const float* ib = inv + (size_t)b * H * W;
    for (int r = threadIdx.y; r < SH; r += EH_TY) {            // a warp per tile row: no integer division, all loads in flight
        const int yy = y0 + r;
        const bool rok = yy >= 0 && yy < H;
        const float* rowp = ib + (size_t)(rok ? yy : 0) * W;
#pragma unroll
        for (int c0 = 0; c0 < SW; c0 += EH_TX) {
            const int c = c0 + threadIdx.x, xx = x0 + c;
            if (c < SW) s_inv[r * SW + c] = (rok && xx >= 0 && xx < W) ? __ldg(rowp + xx) : 0.0f;
        }
    }
    const int x = blockIdx.x * TWP + threadIdx.x * EH_PX, y = blockIdx.y * EH_TY + threadIdx.y;
    const bool inside = x < W && y < H;
    const bool has[EH_PX] = {true, x + 1 < W};
    // relu(convc1(cost) + b): 1x1 (loads issued before the barrier)
    float cv[EH_PX][8];
#pragma unroll
    for (int p = 0; p < EH_PX; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) cv[p][c] = (c < CD && inside && has[p]) ? __ldg(cost + (((size_t)b * CD + c) * H + y) * W + x + p) : 0.0f;
    __syncthreads();
    if (!inside) return;
    float4* o[EH_PX];
#pragma unroll
    for (int p = 0; p < EH_PX; ++p) o[p] = reinterpret_cast<float4*>(out + (((size_t)b * H + y) * W + (has[p] ? x + p : x)) * (2 * h) + ch0);
    {
        float acc[EH_PX][16];
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[p][j] = T.bc[j];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c < CD) {
#pragma unroll
                for (int p = 0; p < EH_PX; ++p)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(cv[p][c], T.wc[c][j], acc[p][j]);
            }
        }
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
            if (has[p])
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[p][q] = make_float4(fmaxf(acc[p][4 * q], 0.0f), fmaxf(acc[p][4 * q + 1], 0.0f), fmaxf(acc[p][4 * q + 2], 0.0f),
                                          fmaxf(acc[p][4 * q + 3], 0.0f));
    }
    // relu(convd1(inv) + b): 7x7, zero padding 3; the two pixels of a thread share a row of 8 window values (four 64-bit loads)
    {
        const float2* win = reinterpret_cast<const float2*>(s_inv + threadIdx.y * SW + threadIdx.x * EH_PX);
        float acc[EH_PX][16];
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[p][j] = T.bd[j];
#pragma unroll
        for (int ty = 0; ty < 7; ++ty) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 t = win[ty * (SW / 2) + i];
                v[2 * i] = t.x;
                v[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int tx = 0; tx < 7; ++tx)
#pragma unroll
                for (int p = 0; p < EH_PX; ++p)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(v[tx + p], T.wd[ty * 7 + tx][j], acc[p][j]);
        }
#pragma unroll
        for (int p = 0; p < EH_PX; ++p)
            if (has[p])
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[p][(h >> 2) + q] = make_float4(fmaxf(acc[p][4 * q], 0.0f), fmaxf(acc[p][4 * q + 1], 0.0f), fmaxf(acc[p][4 * q + 2], 0.0f),
                                                     fmaxf(acc[p][4 * q + 3], 0.0f));
    }
}

}  // namespace
}  // namespace effimvs

extern "C" int effimvs_encoder_head_table_floats(int h) { return (h > 0 && h % 16 == 0) ? (h / 16) * 944 : 0; }

extern "C" int effimvs_encoder_head_pack_host(const float* wc1, const float* bc1, const float* wd1, const float* bd1, int CD, int h,
                                              float* tables_out) {
    using namespace effimvs;
    EFFI_REQUIRE(wc1 && bc1 && wd1 && bd1 && tables_out, EFFIMVS_EINVAL, "encoder_head_pack_host: null pointer");
    EFFI_REQUIRE(CD >= 1 && CD <= 8 && h >= 16 && h % 16 == 0 && h <= 128, EFFIMVS_EUNSUPPORTED,
                 "encoder_head_pack_host: cost channels %d must be in [1,8], hidden %d a multiple of 16 up to 128", CD, h);
    EHTable* T = reinterpret_cast<EHTable*>(tables_out);
    for (int k = 0; k < h / 16; ++k)
        for (int j = 0; j < 16; ++j) {
            const int ch = k * 16 + j;
            for (int t = 0; t < 49; ++t) T[k].wd[t][j] = wd1[ch * 49 + t];
            for (int c = 0; c < 8; ++c) T[k].wc[c][j] = c < CD ? wc1[ch * CD + c] : 0.0f;
            T[k].bc[j] = bc1[ch];
            T[k].bd[j] = bd1[ch];
        }
    return EFFIMVS_OK;
}

extern "C" int effimvs_encoder_head_hostw_f32(const float* cost, const float* inv, const float* host_tables, int B, int CD, int h, int H,
                                              int W, float* out, void* stream) {
    using namespace effimvs;
    EFFI_REQUIRE(cost && inv && host_tables && out, EFFIMVS_EINVAL, "encoder_head_hostw: null pointer");
    EFFI_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, EFFIMVS_EINVAL, "encoder_head_hostw: bad sizes");
    EFFI_REQUIRE(CD >= 1 && CD <= 8 && h >= 16 && h % 16 == 0 && h <= 128, EFFIMVS_EUNSUPPORTED,
                 "encoder_head_hostw: cost channels %d must be in [1,8], hidden %d a multiple of 16 up to 128", CD, h);
    constexpr int TY = 4;
    dim3 block(EH_TX, TY), grid(ceil_div(W, EH_TX * EH_PX), ceil_div(H, TY), B);
    const EHTable* T = reinterpret_cast<const EHTable*>(host_tables);
    for (int k = 0; k < h / 16; ++k)       // the table is copied into the launch's parameter buffer: the host array may change afterwards
        launch_kernel(encoder_head_const_kernel<TY>, grid, block, 0, (cudaStream_t)stream, T[k], cost, CD, inv, h, k * 16, H, W, out);
    return check_launch("encoder_head_const_kernel");
}

extern "C" int effimvs_encoder_head_f32(const float* cost, const float* inv, const float* wc1, const float* bc1, const float* wd1,
                                        const float* bd1, int B, int CD, int h, int H, int W, float* out, void* stream) {
    EFFI_REQUIRE(cost && inv && wc1 && bc1 && wd1 && bd1 && out, EFFIMVS_EINVAL, "encoder_head: null pointer");
    EFFI_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, EFFIMVS_EINVAL, "encoder_head: bad sizes");
    EFFI_REQUIRE(CD >= 1 && CD <= 8 && h >= 16 && h % 16 == 0 && h <= 128, EFFIMVS_EUNSUPPORTED,
                 "encoder_head: cost channels %d must be in [1,8], hidden %d a multiple of 16 up to 128", CD, h);
    const int tiles_x = effimvs::ceil_div(W, effimvs::EH_TX * effimvs::EH_PX);
    static const int force_ty = [] { const char* e = getenv("EFFIMVS_EH_TY"); return e ? atoi(e) : 0; }();   // tuning switch: 4 or 8
    const int ty = force_ty == 8 ? 8 : 4;   // measured per DTU depth map (9 launches): 64 x 4 tiles 6.65 ms, 64 x 8 tiles 6.70 ms
    dim3 block(effimvs::EH_TX, ty), grid(tiles_x, effimvs::ceil_div(H, ty), B);
    const size_t smem = (size_t)(49 * h + CD * h + 2 * h + (effimvs::EH_TX * effimvs::EH_PX + 6) * (ty + 6)) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    static const int zc = [] { const char* e = getenv("EFFIMVS_EH_ZCHUNKS"); return e ? atoi(e) : 1; }();   // tuning switch
    if (ty == 8) launch_kernel(effimvs::encoder_head_kernel<8, 0, false>, grid, block, smem, st, cost, CD, inv, wc1, bc1, wd1, bd1, h, H, W, out);
    else if (h == 16) launch_kernel(effimvs::encoder_head_kernel<4, 16, false>, grid, block, smem, st, cost, CD, inv, wc1, bc1, wd1, bd1, h, H, W, out);
    else if (h == 48 && zc) launch_kernel(effimvs::encoder_head_kernel<4, 48, true>, grid, dim3(block.x, block.y, 3), smem, st, cost, CD, inv, wc1, bc1, wd1, bd1, h, H, W, out);
    else if (h == 32) launch_kernel(effimvs::encoder_head_kernel<4, 32, false>, grid, block, smem, st, cost, CD, inv, wc1, bc1, wd1, bd1, h, H, W, out);
    else launch_kernel(effimvs::encoder_head_kernel<4, 0, false>, grid, block, smem, st, cost, CD, inv, wc1, bc1, wd1, bd1, h, H, W, out);
    return effimvs::check_launch("encoder_head_kernel");
}

// ------------------------------------------------------------------------------------------------
// ProjectionInput tail (upstream models/update.py:93-95): relu(convc(cat[cor_dfm, context])), a 1x1
// convolution.  Its context half (+ both biases) does not change over the GRU iterations and arrives as
// ctx_term; the rest is a tiny per-pixel GEMM (hm x h, hm = h - context channels) whose result is written
// straight into the x half of the GRU's input map hx = cat[h, x] (no separate output + copy).
// ------------------------------------------------------------------------------------------------
namespace effimvs {
namespace {

template <int HM4>   // input channels / 4
__global__ void __launch_bounds__(256)
encoder_tail_kernel(const float4* __restrict__ m, const float* __restrict__ w, const float4* __restrict__ ctx_term, long long n_pix,
                    int h, float4* __restrict__ hx) {
    extern __shared__ __align__(16) float tsm[];   // [hm][h]: input-major so that 4 outputs are one 128-bit broadcast load
    constexpr int HM = HM4 * 4;
    pdl_trigger();
    for (int i = threadIdx.x; i < HM * h; i += blockDim.x) tsm[i] = w[(i % h) * HM + i / h];
    pdl_wait();
    __syncthreads();
    const int h4 = h >> 2;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (long long)gridDim.x * blockDim.x) {
        float mv[HM];
#pragma unroll
        for (int q = 0; q < HM4; ++q) {
            const float4 v = __ldg(m + p * HM4 + q);
            mv[4 * q] = v.x; mv[4 * q + 1] = v.y; mv[4 * q + 2] = v.z; mv[4 * q + 3] = v.w;
        }
        for (int c4 = 0; c4 < h4; ++c4) {
            float4 acc = __ldg(ctx_term + p * h4 + c4);
#pragma unroll
            for (int k = 0; k < HM; ++k) {
                const float4 wv = *reinterpret_cast<const float4*>(tsm + k * h + c4 * 4);
                acc.x = fmaf(mv[k], wv.x, acc.x); acc.y = fmaf(mv[k], wv.y, acc.y);
                acc.z = fmaf(mv[k], wv.z, acc.z); acc.w = fmaf(mv[k], wv.w, acc.w);
            }
            hx[p * (2 * h4) + h4 + c4] = make_float4(fmaxf(acc.x, 0.0f), fmaxf(acc.y, 0.0f), fmaxf(acc.z, 0.0f), fmaxf(acc.w, 0.0f));
        }
    }
}

}  // namespace
}  // namespace effimvs

// ------------------------------------------------------------------------------------------------
// The same tail with the context half computed in the kernel: x = relu(Wm m + Wctx act(context) + bias).
// The context map has 4 / 8 / 12 channels, so reading it (instead of a precomputed h-channel ctx_term) is
// less traffic than the term it replaces, and the 1x1 convolution + bias-add launches that formed the term
// once per stage disappear.  ``ctx`` is read at a pixel stride of ctx_stride floats (a channel range of the
// context network's output map), optionally through relu (models/Effi_MVS_plus.py:466).
// gru_init: hx[:, :h] = tanh(ctx_map[:, :h]) -- the hidden state the GRU starts from (Effi_MVS_plus.py:465).
// ------------------------------------------------------------------------------------------------
namespace effimvs {
namespace {

template <int HM4>
__global__ void __launch_bounds__(256)
encoder_tail_ctx_kernel(const float4* __restrict__ m, const float* __restrict__ w_m, const float* __restrict__ ctx, int ctx_stride, int cx,
                        int ctx_relu, const float* __restrict__ w_ctx, const float* __restrict__ bias, long long n_pix, int h,
                        float4* __restrict__ hx) {
    extern __shared__ __align__(16) float tsm[];   // [hm][h] | [cx][h] | [h]
    constexpr int HM = HM4 * 4;
    float* tcx = tsm + HM * h;
    float* tb = tcx + cx * h;
    pdl_trigger();
    for (int i = threadIdx.x; i < HM * h; i += blockDim.x) tsm[i] = w_m[(i % h) * HM + i / h];
    for (int i = threadIdx.x; i < cx * h; i += blockDim.x) tcx[i] = w_ctx[(i % h) * cx + i / h];
    for (int i = threadIdx.x; i < h; i += blockDim.x) tb[i] = bias[i];
    pdl_wait();
    __syncthreads();
    const int h4 = h >> 2;
    // gridDim.y slices of the output channels (block-uniform, so the weight reads stay broadcasts): on the 1/8-resolution stage
    // one thread per pixel is 116 blocks of 2,300 dependent FMAs each -- a fraction of the machine (18 us for 29,600 pixels)
    const int c4_lo = (int)blockIdx.y * h4 / (int)gridDim.y, c4_hi = ((int)blockIdx.y + 1) * h4 / (int)gridDim.y;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (long long)gridDim.x * blockDim.x) {
        float mv[HM], cv[12];
#pragma unroll
        for (int q = 0; q < HM4; ++q) {
            const float4 v = __ldg(m + p * HM4 + q);
            mv[4 * q] = v.x; mv[4 * q + 1] = v.y; mv[4 * q + 2] = v.z; mv[4 * q + 3] = v.w;
        }
        const float4* cp = reinterpret_cast<const float4*>(ctx + p * ctx_stride);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (4 * q < cx) v = __ldg(cp + q);
            if (ctx_relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
            cv[4 * q] = v.x; cv[4 * q + 1] = v.y; cv[4 * q + 2] = v.z; cv[4 * q + 3] = v.w;
        }
        for (int c4 = c4_lo; c4 < c4_hi; ++c4) {
            float4 acc = *reinterpret_cast<const float4*>(tb + c4 * 4);
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                if (k < cx) {
                    const float4 wv = *reinterpret_cast<const float4*>(tcx + k * h + c4 * 4);
                    acc.x = fmaf(cv[k], wv.x, acc.x); acc.y = fmaf(cv[k], wv.y, acc.y);
                    acc.z = fmaf(cv[k], wv.z, acc.z); acc.w = fmaf(cv[k], wv.w, acc.w);
                }
            }
#pragma unroll
            for (int k = 0; k < HM; ++k) {
                const float4 wv = *reinterpret_cast<const float4*>(tsm + k * h + c4 * 4);
                acc.x = fmaf(mv[k], wv.x, acc.x); acc.y = fmaf(mv[k], wv.y, acc.y);
                acc.z = fmaf(mv[k], wv.z, acc.z); acc.w = fmaf(mv[k], wv.w, acc.w);
            }
            hx[p * (2 * h4) + h4 + c4] = make_float4(fmaxf(acc.x, 0.0f), fmaxf(acc.y, 0.0f), fmaxf(acc.z, 0.0f), fmaxf(acc.w, 0.0f));
        }
    }
}

__global__ void __launch_bounds__(256)
gru_init_kernel(const float4* __restrict__ ctx_map, long long n_pix, int h4, int ct4, float4* __restrict__ hx) {
    pdl_enter();
    const long long total = n_pix * h4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / h4;
        const int c4 = (int)(i - p * h4);
        const float4 v = __ldg(ctx_map + p * ct4 + c4);
        hx[p * (2 * h4) + c4] = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
    }
}

// gru_init + the context half of the encoder's 1x1 output convolution, formed ONCE per stage: ctx_term = w_ctx relu(ctx_map[:, h:]) + bias
// -- the map the ADD_RELU epilogue of conv2d_tc adds in every GRU iteration.  One pass over the context map instead of relu
// (a slice copy), a cuDNN 1x1 convolution and its bias add; a thread = (pixel, quad of output channels), the cx context values
// of a pixel are shared by its h / 4 threads through L1, the weights sit in shared memory as [k][h].
__global__ void __launch_bounds__(256)
gru_init_ctx_kernel(const float4* __restrict__ ctx_map, long long n_pix, int h4, int cx4, const float* __restrict__ w_ctx,
                    const float* __restrict__ bias, float4* __restrict__ hx, float4* __restrict__ ctx_term) {
    extern __shared__ __align__(16) float gsm[];   // [cx][h] | [h]
    const int h = h4 * 4, cx = cx4 * 4, ct4 = h4 + cx4;
    float* tb = gsm + cx * h;
    pdl_trigger();
    for (int i = threadIdx.x; i < cx * h; i += blockDim.x) gsm[i] = w_ctx[(i % h) * cx + i / h];   // weight (h, cx) -> [k][o]
    for (int i = threadIdx.x; i < h; i += blockDim.x) tb[i] = bias[i];
    pdl_wait();
    __syncthreads();
    const long long total = n_pix * h4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / h4;
        const int c4 = (int)(i - p * h4);
        const float4 v = __ldg(ctx_map + p * ct4 + c4);
        hx[p * (2 * h4) + c4] = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
        float4 acc = *reinterpret_cast<const float4*>(tb + c4 * 4);
        for (int k4 = 0; k4 < cx4; ++k4) {
            const float4 c = __ldg(ctx_map + p * ct4 + h4 + k4);
            const float cv[4] = {fmaxf(c.x, 0.0f), fmaxf(c.y, 0.0f), fmaxf(c.z, 0.0f), fmaxf(c.w, 0.0f)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(gsm + (k4 * 4 + j) * h + c4 * 4);
                acc.x = fmaf(cv[j], wv.x, acc.x); acc.y = fmaf(cv[j], wv.y, acc.y);
                acc.z = fmaf(cv[j], wv.z, acc.z); acc.w = fmaf(cv[j], wv.w, acc.w);
            }
        }
        ctx_term[i] = acc;
    }
}

}  // namespace
}  // namespace effimvs

extern "C" int effimvs_gru_init_ctx_f32(const float* ctx_map, long long n_pix, int h, int cx, const float* w_ctx, const float* bias,
                                        float* hx, float* ctx_term, void* stream) {
    using namespace effimvs;
    EFFI_REQUIRE(ctx_map && w_ctx && bias && hx && ctx_term, EFFIMVS_EINVAL, "gru_init_ctx: null pointer");
    EFFI_REQUIRE(n_pix > 0 && h >= 4 && h % 4 == 0 && h <= 128 && cx >= 4 && cx % 4 == 0 && cx <= 64, EFFIMVS_EINVAL,
                 "gru_init_ctx: h=%d (<= 128), cx=%d (4..64) must be multiples of 4", h, cx);
    const long long blocks = (n_pix * (h / 4) + 255) / 256;
    const int grid = (int)(blocks < (long long)kNumSMs * 16 ? blocks : (long long)kNumSMs * 16);
    const size_t smem = (size_t)(cx * h + h) * sizeof(float);
    launch_kernel(gru_init_ctx_kernel, grid, dim3(256), smem, (cudaStream_t)stream, (const float4*)ctx_map, n_pix, h / 4, cx / 4, w_ctx, bias,
                  (float4*)hx, (float4*)ctx_term);
    return check_launch("gru_init_ctx_kernel");
}

extern "C" int effimvs_gru_init_f32(const float* ctx_map, long long n_pix, int h, int cx, float* hx, void* stream) {
    using namespace effimvs;
    EFFI_REQUIRE(ctx_map && hx, EFFIMVS_EINVAL, "gru_init: null pointer");
    EFFI_REQUIRE(n_pix > 0 && h >= 4 && h % 4 == 0 && cx >= 0 && cx % 4 == 0, EFFIMVS_EINVAL,
                 "gru_init: h=%d, cx=%d must be multiples of 4", h, cx);
    const long long blocks = (n_pix * (h / 4) + 255) / 256;
    const int grid = (int)(blocks < (long long)kNumSMs * 16 ? blocks : (long long)kNumSMs * 16);
    launch_kernel(gru_init_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (const float4*)ctx_map, n_pix, h / 4, (h + cx) / 4, (float4*)hx);
    return check_launch("gru_init_kernel");
}

extern "C" int effimvs_encoder_tail_ctx_f32(const float* m, const float* w_m, const float* ctx, int ctx_stride, int cx, int ctx_relu,
                                            const float* w_ctx, const float* bias, long long n_pix, int hm, int h, float* hx,
                                            void* stream) {
    using namespace effimvs;
    EFFI_REQUIRE(m && w_m && ctx && w_ctx && bias && hx, EFFIMVS_EINVAL, "encoder_tail_ctx: null pointer");
    EFFI_REQUIRE(n_pix > 0 && h >= 4 && h % 4 == 0 && h <= 128, EFFIMVS_EINVAL, "encoder_tail_ctx: h=%d must be a multiple of 4 up to 128", h);
    EFFI_REQUIRE(cx >= 4 && cx <= 12 && cx % 4 == 0 && ctx_stride >= cx && ctx_stride % 4 == 0, EFFIMVS_EUNSUPPORTED,
                 "encoder_tail_ctx: context channels %d must be 4, 8 or 12 (pixel stride %d a multiple of 4)", cx, ctx_stride);
    EFFI_REQUIRE((reinterpret_cast<uintptr_t>(ctx) & 15) == 0, EFFIMVS_EINVAL, "encoder_tail_ctx: ctx must be 16-byte aligned");
    const long long blocks = (n_pix + 255) / 256;
    const int h4 = h / 4;
    // small maps: slice the output channels over gridDim.y until there are a few blocks per SM
    const int slices = blocks >= (long long)kNumSMs * 2 ? 1 : (h4 % 3 == 0 ? 3 : (h4 % 2 == 0 ? 2 : 1));
    const dim3 grid((unsigned)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8), slices);
    const size_t smem = (size_t)(hm * h + cx * h + h) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define EFFI_TAILC_CASE(Q)                                                                                                        \
    case Q * 4:                                                                                                                   \
        launch_kernel(encoder_tail_ctx_kernel<Q>, grid, dim3(256), smem, st, (const float4*)m, w_m, ctx, ctx_stride, cx, ctx_relu, w_ctx, bias, n_pix, h, \
                                                            (float4*)hx);                                                        \
        break;
    switch (hm) {
        EFFI_TAILC_CASE(2) EFFI_TAILC_CASE(3) EFFI_TAILC_CASE(4) EFFI_TAILC_CASE(5) EFFI_TAILC_CASE(6) EFFI_TAILC_CASE(7) EFFI_TAILC_CASE(8)
        EFFI_TAILC_CASE(9) EFFI_TAILC_CASE(10) EFFI_TAILC_CASE(11) EFFI_TAILC_CASE(12)
        default:
            set_error("encoder_tail_ctx: input channels %d must be a multiple of 4 in [8,48]", hm);
            return EFFIMVS_EUNSUPPORTED;
    }
#undef EFFI_TAILC_CASE
    return check_launch("encoder_tail_ctx_kernel");
}

extern "C" int effimvs_encoder_tail_f32(const float* m, const float* w, const float* ctx_term, long long n_pix, int hm, int h,
                                        float* hx, void* stream) {
    using namespace effimvs;
    EFFI_REQUIRE(m && w && ctx_term && hx, EFFIMVS_EINVAL, "encoder_tail: null pointer");
    EFFI_REQUIRE(n_pix > 0 && h >= 4 && h % 4 == 0 && h <= 128, EFFIMVS_EINVAL, "encoder_tail: h=%d must be a multiple of 4 up to 128", h);
    const long long blocks = (n_pix + 255) / 256;
    const int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
    const size_t smem = (size_t)hm * h * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define EFFI_TAIL_CASE(Q)                                                                                        \
    case Q * 4:                                                                                                  \
        launch_kernel(encoder_tail_kernel<Q>, grid, dim3(256), smem, st, (const float4*)m, w, (const float4*)ctx_term, n_pix, h, (float4*)hx); \
        break;
    switch (hm) {
        EFFI_TAIL_CASE(2) EFFI_TAIL_CASE(3) EFFI_TAIL_CASE(4) EFFI_TAIL_CASE(5) EFFI_TAIL_CASE(6) EFFI_TAIL_CASE(7) EFFI_TAIL_CASE(8)
        EFFI_TAIL_CASE(9) EFFI_TAIL_CASE(10) EFFI_TAIL_CASE(11) EFFI_TAIL_CASE(12)
        default:
            set_error("encoder_tail: input channels %d must be a multiple of 4 in [8,48]", hm);
            return EFFIMVS_EUNSUPPORTED;
    }
#undef EFFI_TAIL_CASE
    return check_launch("encoder_tail_kernel");
}
