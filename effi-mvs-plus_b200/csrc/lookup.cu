// Per-pixel kernels of the dynamic cost volume: relative projections, 1-D volume lookup,
// GRU-iteration cost lookup, softmax depth regression + confidence.
//
//   relative projection          upstream models/Effi_MVS_plus.py:34-37, models/module.py:314
//   pro_bilinear_sampler         upstream models/Effi_MVS_plus.py:102-134 (+ depth_to_disp :151-164)
//   GetCost.forward              upstream models/Effi_MVS_plus.py:257-303
//   softmax / depth_regression / photometric confidence
//                                upstream models/Effi_MVS_plus.py:78-88, models/module.py:518-524
//
// All of them are streaming, HBM/L2-bound kernels: one thread per reference pixel, the D values
// of the pixel are read with a stride of H*W (coalesced across the warp).
#include "common.cuh"

namespace effimvs {
namespace {

// ---------------------------------------------------------------------------------------------
// P_src @ inverse(P_ref), fp64 internally (Gauss-Jordan with partial pivoting), one thread / (b, src)
// ---------------------------------------------------------------------------------------------
__device__ void compose44(const float* cam, double P[16]) {
    const float* E = cam;        // 4x4 extrinsic
    const float* K = cam + 16;   // 4x4 holding the 3x3 intrinsic
    for (int i = 0; i < 16; ++i) P[i] = (double)E[i];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            // upstream forms K @ E[:3,:4] in fp32 (torch.matmul); round the product to fp32 likewise
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k) acc = fmaf(K[r * 4 + k], E[k * 4 + c], acc);
            P[r * 4 + c] = (double)acc;
        }
}

__device__ bool invert44(const double A[16], double inv[16]) {
    double M[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            M[r][c] = A[r * 4 + c];
            M[r][c + 4] = (r == c) ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(M[col][col]);
        for (int r = col + 1; r < 4; ++r)
            if (fabs(M[r][col]) > best) { best = fabs(M[r][col]); piv = r; }
        if (best == 0.0) return false;
        if (piv != col)
            for (int c = 0; c < 8; ++c) { double t = M[col][c]; M[col][c] = M[piv][c]; M[piv][c] = t; }
        double s = 1.0 / M[col][col];
        for (int c = 0; c < 8; ++c) M[col][c] *= s;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            double f = M[r][col];
            for (int c = 0; c < 8; ++c) M[r][c] -= f * M[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[r * 4 + c] = M[r][c + 4];
    return true;
}

__global__ void relative_projection_kernel(const float* __restrict__ cams, int B, int V, float* __restrict__ proj) {
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (V - 1)) return;
    int b = i / (V - 1), v = i % (V - 1) + 1;
    double Pr[16], Ps[16], Pi[16];
    compose44(cams + ((size_t)b * V) * 32, Pr);
    compose44(cams + ((size_t)b * V + v) * 32, Ps);
    float* out = proj + (size_t)i * 12;
    if (!invert44(Pr, Pi)) {
        for (int k = 0; k < 12; ++k) out[k] = __int_as_float(0x7fc00000);
        return;
    }
    double R[12];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += Ps[r * 4 + k] * Pi[k * 4 + c];
            R[r * 4 + c] = acc;
        }
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) out[r * 3 + c] = (float)R[r * 4 + c];
        out[9 + r] = (float)R[r * 4 + 3];
    }
}

// ---------------------------------------------------------------------------------------------
// 1-D lookup along D.   t = (1/depth - 1/dmax) / ((1/dmin - 1/dmax) + 1e-10) * (D-1), then the
// normalise / un-normalise round trip of bilinear_sampler + ATen grid_sample (CUDA semantics:
// division by the Python scalar (D-1) is a multiplication by its fp32 reciprocal).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lookup_1d(const float* __restrict__ vol, int D, int HW, float depth, float dmin, float dmax) {
    float inv_max = __fdiv_rn(1.0f, dmax);
    float num = __fsub_rn(__fdiv_rn(1.0f, depth), inv_max);
    float den = __fadd_rn(__fsub_rn(__fdiv_rn(1.0f, dmin), inv_max), 1e-10f);
    float t = __fmul_rn(__fdiv_rn(num, den), (float)(D - 1));
    float g = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, t), __fdiv_rn(1.0f, (float)(D - 1))), 1.0f);
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(D - 1));
    float f = floorf(ix);
    float w1 = __fsub_rn(ix, f);                      // weight of tap f+1
    float w0 = __fsub_rn(__fadd_rn(f, 1.0f), ix);     // weight of tap f
    float out = 0.0f;
    if (f >= 0.0f && f <= (float)(D - 1)) out = __fmul_rn(__ldg(vol + (size_t)((int)f) * HW), w0);
    if (f >= -1.0f && f <= (float)(D - 2)) out = fmaf(__ldg(vol + (size_t)((int)f + 1) * HW), w1, out);
    return out;
}

__global__ void volume_lookup_kernel(const float* __restrict__ volume, const float* __restrict__ sample,
                                     const float* __restrict__ dmin, const float* __restrict__ dmax, int range_mode,
                                     int sstride, int D, int d, int H, int W, float* __restrict__ out) {
    pdl_enter();
    const int b = blockIdx.z;
    const int HW = H * W;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const float lo = range_mode == EFFIMVS_RANGE_PIXEL ? __ldg(dmin + (size_t)b * HW + pix) : __ldg(dmin + b);
    const float hi = range_mode == EFFIMVS_RANGE_PIXEL ? __ldg(dmax + (size_t)b * HW + pix) : __ldg(dmax + b);
    const float* vol = volume + (size_t)b * D * HW + pix;
    size_t spix = pix, sHW = HW;
    if (sstride == 2) {
        int y = pix / W, x = pix - y * W;
        spix = (size_t)(2 * y) * (2 * W) + 2 * x;
        sHW = (size_t)4 * HW;
    }
    for (int k = blockIdx.y; k < d; k += gridDim.y) {
        float depth = __ldg(sample + ((size_t)b * d + k) * sHW + spix);
        out[((size_t)b * d + k) * HW + pix] = lookup_1d(vol, D, HW, depth, lo, hi);
    }
}

__device__ __forceinline__ float local_hypothesis(float cur_depth, float interval, int D, int d) {
    float inv = __fdiv_rn(1.0f, cur_depth);
    float half = __fmul_rn((float)(D / 2), interval);
    float lo = fmaxf(__fsub_rn(inv, half), 1e-4f);
    float hi = fminf(fmaxf(__fadd_rn(inv, half), 1e-4f), 1e4f);
    float step = __fmul_rn(__fsub_rn(hi, lo), __fdiv_rn(1.0f, (float)(D - 1)));   // torch (CUDA) divides by a Python scalar as a * (1/b)
    float s = fmaxf(__fadd_rn(lo, __fmul_rn((float)d, step)), 1e-5f);
    return __fdiv_rn(1.0f, s);
}

// get_cur_depth_range_samples (models/module.py:554-570) on its own: `cur` (B,H,W) in whatever space the
// caller works in (upstream passes inverse depth), `interval` (B) -> samples (B,ndepth,H,W), no reciprocal.
__global__ void range_samples_kernel(const float* __restrict__ cur, const float* __restrict__ interval, int ndepth, int HW,
                                     float* __restrict__ out) {
    pdl_enter();
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const float c = __ldg(cur + (size_t)b * HW + pix);
    const float half = __fmul_rn((float)(ndepth / 2), __ldg(interval + b));
    const float lo = fmaxf(__fsub_rn(c, half), 1e-4f);
    const float hi = fminf(fmaxf(__fadd_rn(c, half), 1e-4f), 1e4f);
    const float step = __fmul_rn(__fsub_rn(hi, lo), __fdiv_rn(1.0f, (float)(ndepth - 1)));   // torch (CUDA): a * (1/b)
    for (int d = 0; d < ndepth; ++d)
        out[((size_t)b * ndepth + d) * HW + pix] = fmaxf(__fadd_rn(lo, __fmul_rn((float)d, step)), 1e-5f);
}

__global__ void dynamic_cost_kernel(const float* __restrict__ cur_depth, const float* __restrict__ raw,
                                    const float* __restrict__ reg, const float* __restrict__ interval,
                                    const float* __restrict__ dmin, const float* __restrict__ dmax, int range_mode,
                                    int ndepth, int D, int HW, float* __restrict__ out) {
    pdl_enter();
    const int b = blockIdx.z;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const float lo = range_mode == EFFIMVS_RANGE_PIXEL ? __ldg(dmin + (size_t)b * HW + pix) : __ldg(dmin + b);
    const float hi = range_mode == EFFIMVS_RANGE_PIXEL ? __ldg(dmax + (size_t)b * HW + pix) : __ldg(dmax + b);
    const float cur = __ldg(cur_depth + (size_t)b * HW + pix);
    const float iv = __ldg(interval + b);
    const float* vraw = raw + (size_t)b * D * HW + pix;
    const float* vreg = reg + (size_t)b * D * HW + pix;
    float* o = out + (size_t)b * 2 * ndepth * HW + pix;
    for (int k = 0; k < ndepth; ++k) {
        float depth = local_hypothesis(cur, iv, ndepth, k);
        o[(size_t)k * HW] = lookup_1d(vraw, D, HW, depth, lo, hi);
        o[(size_t)(ndepth + k) * HW] = lookup_1d(vreg, D, HW, depth, lo, hi);
    }
}

// ---------------------------------------------------------------------------------------------
// softmax over D + expectation + 4-bin confidence, three passes over the pixel's D logits
// (second and third pass hit L1/L2).
// ---------------------------------------------------------------------------------------------
__global__ void softmax_regress_conf_kernel(const float* __restrict__ prob_pre, const float* __restrict__ hyp,
                                            int hyp_mode, int D, int HW, float* __restrict__ depth_out,
                                            float* __restrict__ conf_out) {
    pdl_enter();
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const float* p = prob_pre + (size_t)b * D * HW + pix;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(p + (size_t)d * HW));
    float z = 0.0f;
    for (int d = 0; d < D; ++d) z += expf(__ldg(p + (size_t)d * HW) - m);
    float depth = 0.0f, idxf = 0.0f;
    for (int d = 0; d < D; ++d) {
        float pr = __fdiv_rn(expf(__ldg(p + (size_t)d * HW) - m), z);
        float h = hyp_mode == EFFIMVS_HYP_PLANES ? __ldg(hyp + b * D + d) : __ldg(hyp + ((size_t)b * D + d) * HW + pix);
        depth = fmaf(pr, h, depth);
        idxf = fmaf(pr, (float)d, idxf);
    }
    int idx = (int)idxf;  // trunc, like .long()
    idx = idx < 0 ? 0 : (idx > D - 1 ? D - 1 : idx);
    float conf = 0.0f;
    for (int d = idx - 1; d <= idx + 2; ++d)
        if (d >= 0 && d < D) conf += __fdiv_rn(expf(__ldg(p + (size_t)d * HW) - m), z);
    depth_out[(size_t)b * HW + pix] = depth;
    // upstream: 4 * avg_pool3d(window 4) = (sum of the window) / 4 * 4
    conf_out[(size_t)b * HW + pix] = __fmul_rn(4.0f, __fmul_rn(conf, 0.25f));
}

// uint8 image -> fp32 in [0,1], the arithmetic of upstream's loaders (datasets/general_eval.py:83-87 read_img:
// np.array(img, dtype=np.float32) / 255.): an IEEE division, 16 pixels-channels per thread
__global__ void __launch_bounds__(256)
images_u8_kernel(const uint8_t* __restrict__ in, size_t n, float* __restrict__ out) {
    pdl_enter();
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i + 16 <= n && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + i));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 o;
            o.x = __fdiv_rn((float)(w[q] & 255u), 255.0f);
            o.y = __fdiv_rn((float)((w[q] >> 8) & 255u), 255.0f);
            o.z = __fdiv_rn((float)((w[q] >> 16) & 255u), 255.0f);
            o.w = __fdiv_rn((float)(w[q] >> 24), 255.0f);
            reinterpret_cast<float4*>(out + i)[q] = o;
        }
    } else {
        for (size_t k = i; k < n && k < i + 16; ++k) out[k] = __fdiv_rn((float)in[k], 255.0f);
    }
}

}  // namespace
}  // namespace effimvs

using namespace effimvs;

extern "C" int effimvs_relative_projection_f32(const float* cams, int B, int V, float* proj_out, void* stream) {
    EFFI_REQUIRE(cams && proj_out, EFFIMVS_EINVAL, "relative_projection: null pointer");
    EFFI_REQUIRE(B > 0 && V >= 2, EFFIMVS_EINVAL, "relative_projection: need B > 0 and V >= 2");
    int n = B * (V - 1);
    launch_kernel(relative_projection_kernel, dim3(ceil_div(n, 32)), dim3(32), 0, (cudaStream_t)stream, cams, B, V, proj_out);
    return check_launch("relative_projection_kernel");
}

extern "C" int effimvs_volume_lookup_f32(const float* volume, const float* depth_sample, const float* depth_min,
                                         const float* depth_max, int range_mode, int sample_stride,
                                         int B, int D, int d, int H, int W, float* out, void* stream) {
    EFFI_REQUIRE(volume && depth_sample && depth_min && depth_max && out, EFFIMVS_EINVAL, "volume_lookup: null pointer");
    EFFI_REQUIRE(B > 0 && D > 1 && d > 0 && H > 0 && W > 0, EFFIMVS_EINVAL, "volume_lookup: bad sizes (D must be > 1)");
    EFFI_REQUIRE(range_mode == 0 || range_mode == 1, EFFIMVS_EINVAL, "volume_lookup: range_mode=%d", range_mode);
    EFFI_REQUIRE(sample_stride == 1 || sample_stride == 2, EFFIMVS_EINVAL, "volume_lookup: sample_stride=%d", sample_stride);
    dim3 block(128), grid(ceil_div(H * W, 128), d < 8 ? d : 8, B);
    launch_kernel(volume_lookup_kernel, grid, block, 0, (cudaStream_t)stream, volume, depth_sample, depth_min, depth_max, range_mode,
                                                                  sample_stride, D, d, H, W, out);
    return check_launch("volume_lookup_kernel");
}

extern "C" int effimvs_dynamic_cost_f32(const float* cur_depth, const float* raw_volume, const float* reg_volume,
                                        const float* interval, const float* depth_min, const float* depth_max,
                                        int range_mode, int ndepth, int B, int D, int H, int W, float* out, void* stream) {
    EFFI_REQUIRE(cur_depth && raw_volume && reg_volume && interval && depth_min && depth_max && out, EFFIMVS_EINVAL,
                 "dynamic_cost: null pointer");
    EFFI_REQUIRE(B > 0 && D > 1 && ndepth > 1 && H > 0 && W > 0, EFFIMVS_EINVAL, "dynamic_cost: bad sizes");
    EFFI_REQUIRE(range_mode == 0 || range_mode == 1, EFFIMVS_EINVAL, "dynamic_cost: range_mode=%d", range_mode);
    dim3 block(128), grid(ceil_div(H * W, 128), 1, B);
    launch_kernel(dynamic_cost_kernel, grid, block, 0, (cudaStream_t)stream, cur_depth, raw_volume, reg_volume, interval, depth_min,
                                                                 depth_max, range_mode, ndepth, D, H * W, out);
    return check_launch("dynamic_cost_kernel");
}

extern "C" int effimvs_softmax_regress_conf_f32(const float* prob_pre, const float* hyp, int hyp_mode,
                                                int B, int D, int H, int W, float* depth_out, float* conf_out, void* stream) {
    EFFI_REQUIRE(prob_pre && hyp && depth_out && conf_out, EFFIMVS_EINVAL, "softmax_regress_conf: null pointer");
    EFFI_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, EFFIMVS_EINVAL, "softmax_regress_conf: bad sizes");
    EFFI_REQUIRE(hyp_mode == EFFIMVS_HYP_TENSOR || hyp_mode == EFFIMVS_HYP_PLANES, EFFIMVS_EINVAL,
                 "softmax_regress_conf: hyp_mode=%d", hyp_mode);
    dim3 block(128), grid(ceil_div(H * W, 128), B);
    launch_kernel(softmax_regress_conf_kernel, grid, block, 0, (cudaStream_t)stream, prob_pre, hyp, hyp_mode, D, H * W, depth_out, conf_out);
    return check_launch("softmax_regress_conf_kernel");
}

namespace effimvs {
namespace {
// Everything the cascade derives from depth_values alone (models/Effi_MVS_plus.py:409-424, models/module.py:577-585), one thread per number instead of
// ~16 single-element torch kernels per forward; the same IEEE operations: torch evaluates tensor / python-scalar on the device as
// a * (1 / b) with the reciprocal formed in fp32, tensor * python-scalar as a * float(b), reciprocal() as 1 / x.
// out = 8 rows of B scalars -- depth_far = 1 / min, depth_near = 1 / max, lo_disp = 1 / depth_far, hi_disp = 1 / depth_near,
// interval_s = ((max - min) * (1 / Dv)) * ratio_s (s = 0..2), 0 -- followed by the (B, D1) plane-sweep hypotheses
// 1 / (min + k * ((max - min) * (1 / (D1 - 1)))): every piece is contiguous for its consumer.
__global__ void depth_ranges_kernel(const float* __restrict__ depth_values, int Dv, int D1, float r0, float r1, float r2, float inv_dv, float inv_d1m1,
                                    float* __restrict__ out) {
    pdl_enter();
    const int b = blockIdx.x, t = threadIdx.x;
    const float mn = __ldg(depth_values + (size_t)b * Dv), mx = __ldg(depth_values + (size_t)b * Dv + Dv - 1);
    const int B = gridDim.x;
    const float span = __fsub_rn(mx, mn);
    if (t < 8) {
        const float far_ = __frcp_rn(mn), near_ = __frcp_rn(mx), unit = __fmul_rn(span, inv_dv);
        float v = 0.0f;
        if (t == 0) v = far_;
        else if (t == 1) v = near_;
        else if (t == 2) v = __frcp_rn(far_);
        else if (t == 3) v = __frcp_rn(near_);
        else if (t == 4) v = __fmul_rn(unit, r0);
        else if (t == 5) v = __fmul_rn(unit, r1);
        else if (t == 6) v = __fmul_rn(unit, r2);
        out[t * B + b] = v;
    }
    float* hyp = out + 8 * (size_t)B + (size_t)b * D1;
    for (int k = t; k < D1; k += blockDim.x) hyp[k] = __frcp_rn(__fadd_rn(mn, __fmul_rn((float)k, __fmul_rn(span, inv_d1m1))));
}
}  // namespace
}  // namespace effimvs

extern "C" int effimvs_depth_ranges_f32(const float* depth_values, int B, int Dv, int D1, const float* ratios3, float* out, void* stream) {
    using namespace effimvs;
    EFFI_REQUIRE(depth_values && ratios3 && out, EFFIMVS_EINVAL, "depth_ranges: null pointer");
    EFFI_REQUIRE(B > 0 && Dv > 1 && D1 > 1, EFFIMVS_EINVAL, "depth_ranges: B=%d, Dv=%d, D1=%d", B, Dv, D1);
    launch_kernel(depth_ranges_kernel, dim3(B), dim3(64), 0, (cudaStream_t)stream, depth_values, Dv, D1, ratios3[0], ratios3[1], ratios3[2],
                  1.0f / (float)Dv, 1.0f / (float)(D1 - 1), out);
    return check_launch("depth_ranges_kernel");
}

extern "C" int effimvs_depth_range_samples_f32(const float* cur, const float* interval, int B, int ndepth, int H, int W,
                                               float* samples_out, void* stream) {
    EFFI_REQUIRE(cur && interval && samples_out, EFFIMVS_EINVAL, "depth_range_samples: null pointer");
    EFFI_REQUIRE(B > 0 && ndepth > 1 && H > 0 && W > 0 && B <= 65535, EFFIMVS_EINVAL, "depth_range_samples: bad sizes");
    dim3 block(128), grid(ceil_div(H * W, 128), B);
    launch_kernel(range_samples_kernel, grid, block, 0, (cudaStream_t)stream, cur, interval, ndepth, H * W, samples_out);
    return check_launch("range_samples_kernel");
}

extern "C" int effimvs_images_u8_to_f32(const unsigned char* images, long long n, float* out, void* stream) {
    EFFI_REQUIRE(images && out, EFFIMVS_EINVAL, "images_u8_to_f32: null pointer");
    EFFI_REQUIRE(n > 0, EFFIMVS_EINVAL, "images_u8_to_f32: n=%lld", n);
    const long long threads = (n + 15) / 16;
    launch_kernel(images_u8_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, images, (size_t)n, out);
    return check_launch("images_u8_kernel");
}
