// Sample coordinates of the homography warp, shared by the warp kernels (warp_corr.cu, warp_tile.cu).
//
// Coordinates exactly as upstream computes them on a CUDA device (models/module.py:318-341; no FMA
// contraction where torch issues separate kernels): ray = rot @ (x,y,1) [torch.matmul -> cuBLAS],
// p = ray * depth + trans, z == 0 -> z + 1e-8, u = px / pz, normalise u * (1 / ((W-1)/2)) - 1 (torch's
// CUDA div-by-scalar multiplies by the reciprocal), ATen un-normalise ((g + 1) / 2) * (W - 1).
#pragma once
#include "common.cuh"

namespace effimvs {

struct Ray { float rx, ry, rz, tx, ty, tz; };   // rot @ (x, y, 1) and trans of one (pixel, source view)

// torch.matmul(rot (B,3,3), xyz (B,3,HW)) lands in different cuBLAS kernels depending on HW, and they
// round differently (tools/ray_order_probe.py, B200, torch 2.11 / cuBLAS 12.8): large maps accumulate
// with an FMA chain, fma(r1, y, r0*x) + r2; small maps form the products separately,
// (r0*x + r1*y) + r2.  The kernels reproduce either order; ray_unfused_for() mirrors the switch.
__device__ __forceinline__ Ray make_ray(const float* __restrict__ P, float x, float y, bool unfused) {
    Ray r;
    if (unfused) {
        r.rx = __fadd_rn(__fadd_rn(__fmul_rn(P[0], x), __fmul_rn(P[1], y)), P[2]);
        r.ry = __fadd_rn(__fadd_rn(__fmul_rn(P[3], x), __fmul_rn(P[4], y)), P[5]);
        r.rz = __fadd_rn(__fadd_rn(__fmul_rn(P[6], x), __fmul_rn(P[7], y)), P[8]);
    } else {
        r.rx = fmaf(P[2], 1.0f, fmaf(P[1], y, __fmul_rn(P[0], x)));
        r.ry = fmaf(P[5], 1.0f, fmaf(P[4], y, __fmul_rn(P[3], x)));
        r.rz = fmaf(P[8], 1.0f, fmaf(P[7], y, __fmul_rn(P[6], x)));
    }
    r.tx = P[9]; r.ty = P[10]; r.tz = P[11];
    return r;
}

constexpr int kRayUnfusedMaxPixels = 220000;   // probed: 400x512 (204,800 px) unfused, 400x600 (240,000 px) FMA chain
inline bool ray_unfused_for(int H, int W) { return (long long)H * W <= kRayUnfusedMaxPixels; }

// un-normalised sample position (ix, iy) in source pixels
__device__ __forceinline__ void sample_coords(const Ray& r, float depth, int H, int W, float inv_half_w, float inv_half_h,
                                              float& ix, float& iy) {
    float px = __fadd_rn(__fmul_rn(r.rx, depth), r.tx);
    float py = __fadd_rn(__fmul_rn(r.ry, depth), r.ty);
    float pz = __fadd_rn(__fmul_rn(r.rz, depth), r.tz);
    if (pz == 0.0f) pz = __fadd_rn(pz, 1e-8f);
    float u, v;
    div2_rn(px, py, pz, u, v);     // = px / pz, py / pz rounded to nearest, one shared reciprocal
    float gx = __fsub_rn(__fmul_rn(u, inv_half_w), 1.0f);
    float gy = __fsub_rn(__fmul_rn(v, inv_half_h), 1.0f);
    // ATen: ((g + 1) / 2) * (size - 1).  The halving is exact, so one multiplication by (size - 1) / 2 (exact too)
    // rounds the same real product: bit-identical, one instruction less per coordinate
    ix = __fmul_rn(__fadd_rn(gx, 1.0f), 0.5f * (float)(W - 1));
    iy = __fmul_rn(__fadd_rn(gy, 1.0f), 0.5f * (float)(H - 1));
}

// inverse-depth samples around cur_depth, exactly the op sequence of models/module.py:558-570
__device__ __forceinline__ float local_hypothesis(float cur_depth, float interval, int D, int d) {
    float inv = __fdiv_rn(1.0f, cur_depth);
    float half = __fmul_rn((float)(D / 2), interval);
    float lo = fmaxf(__fsub_rn(inv, half), 1e-4f);
    float hi = fminf(fmaxf(__fadd_rn(inv, half), 1e-4f), 1e4f);
    float step = __fmul_rn(__fsub_rn(hi, lo), __fdiv_rn(1.0f, (float)(D - 1)));   // torch (CUDA) divides by a Python scalar as a * (1/b)
    float s = fmaxf(__fadd_rn(lo, __fmul_rn((float)d, step)), 1e-5f);
    return __fdiv_rn(1.0f, s);
}

// the same split into the per-pixel part and the per-plane part (identical operations, the first five done once)
struct LocalHyp { float lo, step; };
__device__ __forceinline__ LocalHyp local_hypothesis_prepare(float cur_depth, float interval, int D, float inv_dm1) {
    const float inv = __fdiv_rn(1.0f, cur_depth);
    const float half = __fmul_rn((float)(D / 2), interval);
    LocalHyp h;
    h.lo = fmaxf(__fsub_rn(inv, half), 1e-4f);
    const float hi = fminf(fmaxf(__fadd_rn(inv, half), 1e-4f), 1e4f);
    h.step = __fmul_rn(__fsub_rn(hi, h.lo), inv_dm1);           // inv_dm1 = fl(1 / (D - 1))
    return h;
}
__device__ __forceinline__ float local_hypothesis_at(const LocalHyp& h, int d) {
    return __fdiv_rn(1.0f, fmaxf(__fadd_rn(h.lo, __fmul_rn((float)d, h.step)), 1e-5f));
}

__device__ __forceinline__ float fetch_hypothesis(const float* __restrict__ hyp, int mode, const float* __restrict__ interval,
                                                  int b, int d, int D, int pix, int HW) {
    if (mode == EFFIMVS_HYP_TENSOR) return __ldg(hyp + ((size_t)b * D + d) * HW + pix);
    if (mode == EFFIMVS_HYP_PLANES) return __ldg(hyp + b * D + d);
    return local_hypothesis(__ldg(hyp + (size_t)b * HW + pix), __ldg(interval + b), D, d);
}

}  // namespace effimvs
