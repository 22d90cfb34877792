// Tiled fused homography warp + group-wise correlation + view aggregation for channels-last feature
// maps: the source footprint of a block's reference tile is staged in shared memory by TMA and the
// bilinear taps are read from there.
//
// Replaces the same upstream code as warp_corr.cu (models/module.py:303-344, 554-570;
// models/Effi_MVS_plus.py:39-53, 65-67, 222-244); this file is the fast path for EFFIMVS_FEA_NHWC.
//
// Why a tile: the 4 taps x C channels of every (pixel, plane, view) sample are 16*C bytes of gather
// for 4*(C + 1) bytes of HBM traffic, so the kernel is bound by the on-chip gather rate, not HBM.
// Measured on B200: a warp-wide global load pays ~2 cycles per 128-byte line it touches (~64 B/clk/SM
// through L1), shared memory delivers 128 B/clk/SM when conflict free.  So per source view the block
//   1. computes the sample coordinates of its 32x4 reference pixels x DPT planes exactly as upstream
//      does on a CUDA device (same op order, no contraction) and reduces their bounding box,
//   2. has one thread issue a TMA box load (8 channels x BW x BH source pixels, 32-byte swizzle, zero
//      fill outside the image = grid_sample's zeros padding) per 8-channel block,
//   3. samples from shared memory with 128-bit loads (the swizzle spreads the 32-byte pixels of
//      neighbouring lanes over all banks) and packed fp32 FMAs (FFMA2).
// A (block, view) whose footprint does not fit the box (depth discontinuities, wild hypotheses)
// gathers from global memory with 256-bit loads instead -- same arithmetic, same results.
#include <cuda.h>  // CUtensorMap types only; the encoder is resolved through cudaGetDriverEntryPoint

#include <limits.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "warp_coords.cuh"

namespace effimvs {
namespace {

constexpr int TW = 32, TH = 4;            // reference tile of a block: one warp per tile row
constexpr int TILE_THREADS = TW * TH;
constexpr int BW = 64, BH = 12;           // staged source box in pixels (one 8-channel block of it)
constexpr int BOX_BYTES = BW * BH * 32;
constexpr int ROW_BYTES = BW * 32;        // multiple of 256: the swizzle bit (address bit 7) is row independent
constexpr int FLAG_FORCE_GATHER = 1;      // debugging / tests: never stage, always gather from global
constexpr int FLAG_RAY_UNFUSED = 2;       // ray = (r0*x + r1*y) + r2 without contraction (see warp_coords.cuh)

struct TileMaps {
    CUtensorMap m[EFFIMVS_MAX_SRC_VIEWS];
};

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int x, int y, int b) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(x), "r"(y), "r"(b)
                 : "memory");
}

struct Q2 { u64 a, b; };                 // four consecutive channels as two packed pairs
__device__ __forceinline__ Q2 lds_q2(uint32_t addr) {
    Q2 q;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(q.a), "=l"(q.b) : "r"(addr));
    return q;
}
struct O2 { u64 a, b, c, d; };           // eight consecutive channels
__device__ __forceinline__ O2 ldg_o2(const float* p) {
    O2 o;
    asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(o.a), "=l"(o.b), "=l"(o.c), "=l"(o.d) : "l"(p));
    return o;
}

// One sample: where its 2x2 cell sits and its four axis weights (products formed at use, in
// upstream's operand order).  Non-live samples (no corner inside the image, or NaN/inf coordinates:
// ATen's CUDA kernel gives zero) carry zero weights.
struct Cell {
    int x0, y0;
    float ex, dx, ey, dy;
    bool live;
};

__device__ __forceinline__ Cell make_cell(const Ray& r, float depth, int H, int W, float inv_half_w, float inv_half_h) {
    float ix, iy;
    sample_coords(r, depth, H, W, inv_half_w, inv_half_h, ix, iy);
    const float fx = floorf(ix), fy = floorf(iy);
    Cell c;
    c.live = (fx >= -1.0f) && (fx <= (float)(W - 1)) && (fy >= -1.0f) && (fy <= (float)(H - 1));   // false for NaN/inf
    c.x0 = c.live ? (int)fx : 0;
    c.y0 = c.live ? (int)fy : 0;
    c.ex = c.live ? __fsub_rn(__fadd_rn(fx, 1.0f), ix) : 0.0f;
    c.dx = c.live ? __fsub_rn(ix, fx) : 0.0f;
    c.ey = c.live ? __fsub_rn(__fadd_rn(fy, 1.0f), iy) : 0.0f;
    c.dy = c.live ? __fsub_rn(iy, fy) : 0.0f;
    return c;
}

// interpolate 8 channels (4 packed pairs) of one sample and multiply-accumulate with the reference
template <int NACC, int ACC0, int PPA>
__device__ __forceinline__ void accumulate8(const u64 (&t00)[4], const u64 (&t01)[4], const u64 (&t10)[4], const u64 (&t11)[4],
                                            float w00, float w01, float w10, float w11, const u64* __restrict__ ref2,
                                            u64 (&acc)[NACC]) {
    const u64 W00 = pack2(w00, w00), W01 = pack2(w01, w01), W10 = pack2(w10, w10), W11 = pack2(w11, w11);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        u64 s = mul2(t00[j], W00);
        s = fma2(t01[j], W01, s);
        s = fma2(t10[j], W10, s);
        s = fma2(t11[j], W11, s);
        const int a = ACC0 + j / PPA;
        acc[a] = fma2(s, ref2[j], acc[a]);
    }
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

template <int G> struct TilePlanes { static constexpr int value = G >= 4 ? 4 : 8; };

template <int C, int G>
__global__ void __launch_bounds__(TILE_THREADS, C == 32 ? 3 : 4)
warp_corr_tile_kernel(const __grid_constant__ TileMaps maps, const float* __restrict__ ref_fea,
                      const __grid_constant__ SrcPtrs srcs, int n_src, const float* __restrict__ proj, const float* __restrict__ hyp, int hyp_mode,
                      const float* __restrict__ interval, const float* __restrict__ weights, int H, int W, int D, int tiles_x,
                      int flags, float* __restrict__ sim_out, float* __restrict__ hyp_out) {
    constexpr int DPT = TilePlanes<G>::value;
    constexpr int CG = C / G;                              // channels per group
    constexpr int NACC = CG == 1 ? C / 2 : G;              // packed accumulators per plane
    constexpr int PPA = CG == 1 ? 1 : CG / 2;              // channel pairs per accumulator
    extern __shared__ uint8_t smem_raw[];
    __shared__ float sP[EFFIMVS_MAX_SRC_VIEWS * 12];
    __shared__ int s_box[2][4];                            // min x, min y, max x, max y of the live cells (double buffered over views)
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z;
    const int HW = H * W;
    const uint32_t tile = (smem_u32(smem_raw) + 1023u) & ~1023u;

    for (int i = tid; i < n_src * 12; i += TILE_THREADS) sP[i] = proj[(size_t)b * n_src * 12 + i];
    if (tid < 8) s_box[tid >> 2][tid & 3] = (tid & 2) ? INT_MIN : INT_MAX;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int xi = tx * TW + lane, yi = ty * TH + (tid >> 5);
    const bool inimg = xi < W && yi < H;
    const int pix = inimg ? yi * W + xi : 0;
    const int d0 = blockIdx.y * DPT;
    const float x = (float)xi, y = (float)yi;
    const float inv_half_w = __fdiv_rn(1.0f, (float)((double)(W - 1) / 2.0));
    const float inv_half_h = __fdiv_rn(1.0f, (float)((double)(H - 1) / 2.0));

    u64 ref2[C / 2];
    {
        const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(ref_fea + ((size_t)b * HW + pix) * C);
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            const ulonglong2 v = __ldg(rp + q);
            ref2[2 * q] = v.x;
            ref2[2 * q + 1] = v.y;
        }
    }
    float depth[DPT], num[DPT][G];
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const bool on = inimg && d0 + k < D;
        depth[k] = on ? fetch_hypothesis(hyp, hyp_mode, interval, b, d0 + k, D, pix, HW) : 1.0f;
        if (hyp_out && on) hyp_out[((size_t)b * D + d0 + k) * HW + pix] = depth[k];
#pragma unroll
        for (int g = 0; g < G; ++g) num[k][g] = 0.0f;
    }
    float den = 0.0f;
    uint32_t phase = 0;

    for (int v = 0; v < n_src; ++v) {
        const Ray ray = make_ray(sP + v * 12, x, y, (flags & FLAG_RAY_UNFUSED) != 0);
        const float w = (weights && inimg) ? __ldg(weights + ((size_t)b * n_src + v) * HW + pix) : 1.0f;
        Cell cell[DPT];
        int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
#pragma unroll
        for (int k = 0; k < DPT; ++k) {
            cell[k] = make_cell(ray, depth[k], H, W, inv_half_w, inv_half_h);
            if (!(inimg && d0 + k < D)) {
                cell[k].live = false;
                cell[k].x0 = cell[k].y0 = 0;
                cell[k].ex = cell[k].dx = cell[k].ey = cell[k].dy = 0.0f;
            }
            mnx = min(mnx, cell[k].live ? cell[k].x0 : INT_MAX);
            mny = min(mny, cell[k].live ? cell[k].y0 : INT_MAX);
            mxx = max(mxx, cell[k].live ? cell[k].x0 : INT_MIN);
            mxy = max(mxy, cell[k].live ? cell[k].y0 : INT_MIN);
        }
        mnx = __reduce_min_sync(0xffffffffu, mnx);
        mny = __reduce_min_sync(0xffffffffu, mny);
        mxx = __reduce_max_sync(0xffffffffu, mxx);
        mxy = __reduce_max_sync(0xffffffffu, mxy);
        int* box = s_box[v & 1];
        if (lane == 0 && mnx <= mxx) {
            atomicMin(&box[0], mnx);
            atomicMin(&box[1], mny);
            atomicMax(&box[2], mxx);
            atomicMax(&box[3], mxy);
        }
        __syncthreads();   // box complete; every thread is also done with the tile of the previous view
        const int bx = box[0], by = box[1], ux = box[2], uy = box[3];
        const bool any = bx <= ux;
        const bool fits = any && (ux - bx + 2 <= BW) && (uy - by + 2 <= BH) && !(flags & FLAG_FORCE_GATHER);

        u64 acc[DPT][NACC];
#pragma unroll
        for (int k = 0; k < DPT; ++k)
#pragma unroll
            for (int a = 0; a < NACC; ++a) acc[k][a] = 0ull;

        if (fits) {
            const int org = by * BW + bx;
            uint32_t lin[DPT];   // byte offset of the cell's north-west pixel in the (unswizzled) box
#pragma unroll
            for (int k = 0; k < DPT; ++k)   // non-live samples have zero weights; keep their address inside the box
                lin[k] = cell[k].live ? (uint32_t)(cell[k].y0 * BW + cell[k].x0 - org) * 32u : 0u;
            static_for<0, C / 8>([&](auto cb_c) {
                constexpr int cb = decltype(cb_c)::value;
                if (cb > 0) __syncthreads();   // everyone has sampled the previous channel block
                if (tid == 0) {
                    if (cb == 0) {             // the other box buffer was last read in the previous view: reset it for the next one
                        int* nb = s_box[(v + 1) & 1];
                        nb[0] = INT_MAX; nb[1] = INT_MAX; nb[2] = INT_MIN; nb[3] = INT_MIN;
                    }
                    mbar_expect_tx(&s_bar, BOX_BYTES);
                    tma_load_box(tile, &maps.m[v], &s_bar, cb * 8, bx, by, b);
                }
                mbar_wait(&s_bar, phase);
                phase ^= 1u;
#pragma unroll
                for (int k = 0; k < DPT; ++k) {
                    // 32-byte swizzle: address bit 4 ^= bit 7 (the two 16-byte halves of a pixel swap in every other 128-byte line)
                    const uint32_t a0 = lin[k], a1 = a0 + 32u;
                    const uint32_t p00 = tile + (a0 ^ ((a0 >> 3) & 16u)), p01 = tile + (a1 ^ ((a1 >> 3) & 16u));
                    u64 t00[4], t01[4], t10[4], t11[4];
                    Q2 q;
                    q = lds_q2(p00);                     t00[0] = q.a; t00[1] = q.b;
                    q = lds_q2(p00 ^ 16u);               t00[2] = q.a; t00[3] = q.b;
                    q = lds_q2(p01);                     t01[0] = q.a; t01[1] = q.b;
                    q = lds_q2(p01 ^ 16u);               t01[2] = q.a; t01[3] = q.b;
                    q = lds_q2(p00 + ROW_BYTES);         t10[0] = q.a; t10[1] = q.b;
                    q = lds_q2((p00 ^ 16u) + ROW_BYTES); t10[2] = q.a; t10[3] = q.b;
                    q = lds_q2(p01 + ROW_BYTES);         t11[0] = q.a; t11[1] = q.b;
                    q = lds_q2((p01 ^ 16u) + ROW_BYTES); t11[2] = q.a; t11[3] = q.b;
                    accumulate8<NACC, (CG == 1 ? cb * 4 : (cb * 8) / CG), PPA>(
                        t00, t01, t10, t11, __fmul_rn(cell[k].ex, cell[k].ey), __fmul_rn(cell[k].dx, cell[k].ey),
                        __fmul_rn(cell[k].ex, cell[k].dy), __fmul_rn(cell[k].dx, cell[k].dy), ref2 + cb * 4, acc[k]);
                }
            });
        } else {
            if (tid == 0) {
                int* nb = s_box[(v + 1) & 1];
                nb[0] = INT_MAX; nb[1] = INT_MAX; nb[2] = INT_MIN; nb[3] = INT_MIN;
            }
            __syncthreads();
            if (any) {
                const float* src = srcs.p[v] + (size_t)b * C * HW;
#pragma unroll
                for (int k = 0; k < DPT; ++k) {
                    // 2x2 block clamped into the image; the axis weights move with it (a corner outside the
                    // image gets weight zero, the inside one keeps upstream's weight)
                    const Cell& c = cell[k];
                    const int xc = min(max(c.x0, 0), W - 2), yc = min(max(c.y0, 0), H - 2);
                    const float wl = c.x0 == xc ? c.ex : (c.x0 + 1 == xc ? c.dx : 0.0f);
                    const float wr = c.x0 == xc ? c.dx : (c.x0 == xc + 1 ? c.ex : 0.0f);
                    const float wt = c.y0 == yc ? c.ey : (c.y0 + 1 == yc ? c.dy : 0.0f);
                    const float wb = c.y0 == yc ? c.dy : (c.y0 == yc + 1 ? c.ey : 0.0f);
                    const float* p = src + ((size_t)yc * W + xc) * C;
                    if (c.live) {
                        static_for<0, C / 8>([&](auto cb_c) {
                            constexpr int cb = decltype(cb_c)::value;
                            u64 t00[4], t01[4], t10[4], t11[4];
                            O2 o;
                            o = ldg_o2(p + cb * 8);                     t00[0] = o.a; t00[1] = o.b; t00[2] = o.c; t00[3] = o.d;
                            o = ldg_o2(p + C + cb * 8);                 t01[0] = o.a; t01[1] = o.b; t01[2] = o.c; t01[3] = o.d;
                            o = ldg_o2(p + (size_t)W * C + cb * 8);     t10[0] = o.a; t10[1] = o.b; t10[2] = o.c; t10[3] = o.d;
                            o = ldg_o2(p + (size_t)(W + 1) * C + cb * 8); t11[0] = o.a; t11[1] = o.b; t11[2] = o.c; t11[3] = o.d;
                            accumulate8<NACC, (CG == 1 ? cb * 4 : (cb * 8) / CG), PPA>(t00, t01, t10, t11, __fmul_rn(wl, wt), __fmul_rn(wr, wt),
                                                                                       __fmul_rn(wl, wb), __fmul_rn(wr, wb), ref2 + cb * 4,
                                                                                       acc[k]);
                        });
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < DPT; ++k) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float lo, hi, sim;
                if (CG == 1) {
                    unpack2(acc[k][g / 2], lo, hi);
                    sim = (g & 1) ? hi : lo;
                } else {
                    unpack2(acc[k][g], lo, hi);
                    sim = __fmul_rn(__fadd_rn(lo, hi), 1.0f / CG);
                }
                num[k][g] = weights ? __fadd_rn(num[k][g], __fmul_rn(sim, w)) : __fadd_rn(num[k][g], sim);
            }
        }
        den = __fadd_rn(den, w);
    }
    if (!inimg) return;
    const float div = weights ? __fadd_rn(den, 1e-6f) : (float)n_src;
#pragma unroll
    for (int k = 0; k < DPT; ++k)
        if (d0 + k < D) {
#pragma unroll
            for (int g = 0; g < G; ++g) sim_out[(((size_t)b * G + g) * D + d0 + k) * HW + pix] = __fdiv_rn(num[k][g], div);
        }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// (C, W, H, B) view of a channels-last map; box = 8 channels x BW x BH pixels, 32-byte swizzle, zero fill
int encode_map(CUtensorMap* m, const float* base, int B, int C, int H, int W) {
    EncodeTiledFn enc = tensor_map_encoder();
    EFFI_REQUIRE(enc, EFFIMVS_ECUDA, "warp_corr: cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {8, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EFFI_REQUIRE(r == CUDA_SUCCESS, EFFIMVS_ECUDA, "warp_corr: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EFFIMVS_OK;
}

template <int C, int G>
int launch_tile(const TileMaps& maps, const float* ref, const SrcPtrs& srcs, int n_src, const float* proj, const float* hyp,
                int hyp_mode, const float* interval, const float* weights, int B, int H, int W, int D, int flags, float* sim_out,
                float* hyp_out, cudaStream_t st) {
    const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
    dim3 block(TILE_THREADS), grid(tiles_x * tiles_y, ceil_div(D, TilePlanes<G>::value), B);
    const size_t smem = BOX_BYTES + 1024;
    cudaFuncSetAttribute(warp_corr_tile_kernel<C, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    warp_corr_tile_kernel<C, G><<<grid, block, smem, st>>>(maps, ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, H, W, D,
                                                            tiles_x, flags, sim_out, hyp_out);
    return check_launch("warp_corr_tile_kernel");
}

template <int C>
int tile_dispatch_g(int G, const TileMaps& maps, const float* ref, const SrcPtrs& srcs, int n_src, const float* proj,
                    const float* hyp, int hyp_mode, const float* interval, const float* weights, int B, int H, int W, int D,
                    int flags, float* sim_out, float* hyp_out, cudaStream_t st) {
    switch (G) {
        case 1: return launch_tile<C, 1>(maps, ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 2: return launch_tile<C, 2>(maps, ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 4: return launch_tile<C, 4>(maps, ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 8: return launch_tile<C, 8>(maps, ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
    }
    set_error("warp_corr_agg: G=%d not in {1,2,4,8}", G);
    return EFFIMVS_EUNSUPPORTED;
}

}  // namespace

int warp_flags_from_env(int H, int W) {
    int flags = 0;
    const char* e = getenv("EFFIMVS_WARP_FORCE_GATHER");
    if (e && e[0] == '1') flags |= FLAG_FORCE_GATHER;
    if (ray_unfused_for(H, W)) flags |= FLAG_RAY_UNFUSED;
    return flags;
}

// channels-last fast path of effimvs_warp_corr_agg_f32 (arguments already validated by the caller)
int warp_corr_agg_tile(const float* ref_fea, const SrcPtrs& srcs, int n_src, const float* proj, const float* hyp, int hyp_mode,
                       const float* interval, const float* weights, int B, int C, int H, int W, int D, int G, float* sim_out,
                       float* hyp_out, cudaStream_t st) {
    TileMaps maps;
    for (int v = 0; v < n_src; ++v) {
        int rc = encode_map(&maps.m[v], srcs.p[v], B, C, H, W);
        if (rc) return rc;
    }
    for (int v = n_src; v < EFFIMVS_MAX_SRC_VIEWS; ++v) maps.m[v] = maps.m[0];
    const int flags = warp_flags_from_env(H, W);
    switch (C) {
        case 8: return tile_dispatch_g<8>(G, maps, ref_fea, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 16: return tile_dispatch_g<16>(G, maps, ref_fea, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 32: return tile_dispatch_g<32>(G, maps, ref_fea, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
    }
    set_error("warp_corr_agg: C=%d not in {8,16,32}", C);
    return EFFIMVS_EUNSUPPORTED;
}

}  // namespace effimvs
