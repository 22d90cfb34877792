// Tiled fused homography warp + group-wise correlation (+ view aggregation / softmax entropy) for
// channels-last feature maps: the source footprint of a block's reference tile is staged in shared
// memory by TMA and the bilinear taps are read from there.
//
// Replaces the same upstream code as warp_corr.cu (models/module.py:303-344, 554-570;
// models/Effi_MVS_plus.py:39-53, 65-67, 222-244); this file is the fast path for EFFIMVS_FEA_NHWC.
//
// Why a tile: the 4 taps x C channels of every (pixel, plane, view) sample are 16*C bytes of gather
// for 4*(C + 1) bytes of HBM traffic, so the kernels are bound by the on-chip gather rate, not HBM.
// Measured on B200: a warp-wide global load pays ~2 cycles per 128-byte line it touches (~64 B/clk/SM
// through L1), shared memory delivers 128 B/clk/SM when conflict free.  So per round (a source view,
// and in the stage-1 kernel a group of planes) the block
//   1. computes the sample positions of its 32x4 reference pixels x DPT planes exactly as upstream
//      does on a CUDA device (same op order, no contraction), keeps only (ix, iy) per sample in
//      registers and reduces the bounding box of the live 2x2 cells,
//   2. has one thread issue the TMA box loads of ALL channels at once (C/8 sub-boxes of 8 channels x BW
//      x BH source pixels, 32-byte swizzle, zero fill outside the image = grid_sample's zeros
//      padding) on one mbarrier,
//   3. re-derives cell and weights from (ix, iy) and samples from shared memory with 128-bit loads at
//      immediate offsets (the swizzle spreads the 32-byte pixels of neighbouring lanes over all
//      banks) and packed fp32 FMAs (FFMA2).
// Several blocks are resident per SM, so the TMA round trip of one block hides behind the sampling
// of the others.  A round whose footprint does not fit the box (depth discontinuities, wild
// hypotheses, strong in-plane rotation) gathers from global memory with 256-bit loads instead --
// same arithmetic, same results.
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
constexpr int FLAG_FORCE_GATHER = 1;      // debugging / tests: never stage, always gather from global
constexpr int FLAG_RAY_UNFUSED = 2;
// profiling only (EFFIMVS_WARP_DEBUG bit mask; results are wrong): no prefetch / no footprint loads / no per-plane lookups
constexpr int FLAG_DBG_NO_PREFETCH = 16, FLAG_DBG_NO_LOADS = 32, FLAG_DBG_NO_SAMPLES = 64;       // ray = (r0*x + r1*y) + r2 without contraction (see warp_coords.cuh)

// staged source box in pixels, per channel count (one sub-box holds 8 channels = 32 bytes per pixel).
// BW * 32 is a multiple of 256 so that the swizzle bit (address bit 7) does not depend on the row.
#ifndef EFFI_BOX32_W
#define EFFI_BOX32_W 64
#endif
#ifndef EFFI_BOX32_H
#define EFFI_BOX32_H 8
#endif
template <int C> struct Box { static constexpr int W = 64, H = 12; };
template <> struct Box<32> { static constexpr int W = EFFI_BOX32_W, H = EFFI_BOX32_H; };
template <int C> struct BoxBytes {
    static constexpr int ROW = Box<C>::W * 32;
    static constexpr int SUB = Box<C>::W * Box<C>::H * 32;
    static constexpr int ALL = SUB * (C / 8);
};
#ifndef EFFI_TILE_BPS8
#define EFFI_TILE_BPS8 5
#endif
#ifndef EFFI_TILE_BPS16
#define EFFI_TILE_BPS16 4
#endif
#ifndef EFFI_TILE_BPS32
#define EFFI_TILE_BPS32 3
#endif
template <int C> struct BlocksPerSM { static constexpr int value = C == 8 ? EFFI_TILE_BPS8 : (C == 16 ? EFFI_TILE_BPS16 : EFFI_TILE_BPS32); };

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
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
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

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct O2 { u64 a, b, c, d; };           // eight consecutive channels
__device__ __forceinline__ O2 ldg_o2(const float* p) {
    O2 o;
    asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(o.a), "=l"(o.b), "=l"(o.c), "=l"(o.d) : "l"(p));
    return o;
}

// interpolate 8 channels (4 packed pairs) of one sample and multiply-accumulate with the reference
template <int NACC, int ACC0, int PPA>
__device__ __forceinline__ void accumulate8(const u64 (&t00)[4], const u64 (&t01)[4], const u64 (&t10)[4], const u64 (&t11)[4],
                                            u64 W00, u64 W01, u64 W10, u64 W11, const u64* __restrict__ ref2, u64 (&acc)[NACC]) {
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

#ifndef EFFI_TILE_DPT1
#define EFFI_TILE_DPT1 8      // planes per thread of the aggregated kernel with one group (tuning builds: 4)
#endif
template <int G> struct TilePlanes { static constexpr int value = G >= 4 ? 4 : EFFI_TILE_DPT1; };

// constants every thread would otherwise derive with an fp64 / IEEE division of its own (87 + 50 warp instructions per warp
// for the two normalisation factors alone): computed once on the host, same roundings
struct alignas(8) Pair { float lo, hi; };
struct SegConsts {
    Pair ihw2, ihh2, hw2, hh2;     // {inv_half_w} x 2, {inv_half_h} x 2, {(W - 1) / 2} x 2, {(H - 1) / 2} x 2: operands of the packed chain
    Pair one2;                     // {1, 1} as a run-time value: see add2_after_mul()
    float inv_half_w, inv_half_h, inv_dm1;
};
__device__ __forceinline__ u64 pair_bits(const Pair& p) { return *reinterpret_cast<const u64*>(&p); }
// a * b (rounded) + c (rounded), as upstream's separate multiplication and addition kernels round it.  ptxas contracts
// mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 (a single rounding) even with explicit rounding modifiers and
// -fmad=false (checked in the SASS), so the addition is issued as fma(product, 1, c) with a one the compiler cannot see
// through: the same IEEE sum, one issue slot, no contraction.
__device__ __forceinline__ u64 mul_then_add2(u64 a, u64 b, u64 c, u64 one_rt) { return fma2(mul2(a, b), one_rt, c); }

// The coordinate chain of sample_coords() for TWO depth planes of one (pixel, view) at once on the packed fp32 pipe
// (add / mul / fma.rn.f32x2: the same IEEE operations in the same order, so both results are bit-identical to the scalar
// chain) -- 17 instead of 29 issue slots per sample in a kernel whose limit is instruction issue.  Every component of the
// ray is kept duplicated in both halves of a 64-bit register.
struct Ray2 { u64 rx, ry, rz, tx, ty, tz; };
__device__ __forceinline__ Ray2 make_ray2(const Ray& r) {
    Ray2 q;
    q.rx = pack2(r.rx, r.rx); q.ry = pack2(r.ry, r.ry); q.rz = pack2(r.rz, r.rz);
    q.tx = pack2(r.tx, r.tx); q.ty = pack2(r.ty, r.ty); q.tz = pack2(r.tz, r.tz);
    return q;
}
__device__ __forceinline__ void sample_coords2(const Ray2& r, float da, float db, const SegConsts& kc, float& ixa, float& iya,
                                               float& ixb, float& iyb) {
    const u64 d = pack2(da, db), one_rt = pair_bits(kc.one2);
    const u64 px = mul_then_add2(r.rx, d, r.tx, one_rt), py = mul_then_add2(r.ry, d, r.ty, one_rt);
    float za, zb;
    unpack2(mul_then_add2(r.rz, d, r.tz, one_rt), za, zb);
    if (za == 0.0f) za = __fadd_rn(za, 1e-8f);
    if (zb == 0.0f) zb = __fadd_rn(zb, 1e-8f);
    float ra, rb;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(za));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(zb));
    const u64 nz = pack2(-za, -zb), one = pack2(1.0f, 1.0f), mone = pack2(-1.0f, -1.0f);
    u64 rr = pack2(ra, rb);
    rr = fma2(rr, fma2(nz, rr, one), rr);                 // div2_rn's sequence, two lanes wide
    u64 u = mul2(px, rr), v = mul2(py, rr);
    u = fma2(fma2(nz, u, px), rr, u);
    v = fma2(fma2(nz, v, py), rr, v);
    u = fma2(fma2(nz, u, px), rr, u);
    v = fma2(fma2(nz, v, py), rr, v);
    const u64 gx = mul_then_add2(u, pair_bits(kc.ihw2), mone, one_rt), gy = mul_then_add2(v, pair_bits(kc.ihh2), mone, one_rt);   // x - 1 == x + (-1)
    unpack2(mul2(add2(gx, one), pair_bits(kc.hw2)), ixa, ixb);
    unpack2(mul2(add2(gy, one), pair_bits(kc.hh2)), iya, iyb);
}

// Shared state of a block: projection rows, the double-buffered bounding box and the TMA barrier.
struct TileShared {
    float P[EFFIMVS_MAX_SRC_VIEWS * 12];
    int box[2][4];                         // min x, min y, max x, max y of the live cells
    alignas(8) uint64_t bar;
};

// One sample re-derived from its position: the 2x2 cell and the four axis weights (products formed at
// use, in upstream's operand order).  Non-live samples (no corner inside the image, or NaN/inf
// coordinates: ATen's CUDA kernel gives zero) contribute a similarity of exactly zero.
struct Cell {
    int x0, y0;
    float ex, dx, ey, dy;
    bool live;
};
__device__ __forceinline__ Cell cell_at(float ix, float iy, int H, int W) {
    const float fx = floorf(ix), fy = floorf(iy);
    Cell c;
    c.live = (fx >= -1.0f) && (fx <= (float)(W - 1)) && (fy >= -1.0f) && (fy <= (float)(H - 1));   // false for NaN/inf
    c.x0 = (int)fx;
    c.y0 = (int)fy;
    c.ex = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
    c.dx = __fsub_rn(ix, fx);
    c.ey = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
    c.dy = __fsub_rn(iy, fy);
    return c;
}

template <int C, int G> struct AccShape {
    static constexpr int CG = C / G;
    static constexpr int NACC = CG == 1 ? C / 2 : G;       // packed accumulators per plane
    static constexpr int PPA = CG == 1 ? 1 : CG / 2;       // channel pairs per accumulator
};

template <int C, int G, class F>
__device__ __forceinline__ void emit_sims(const u64 (&acc)[AccShape<C, G>::NACC], F&& consume) {
    constexpr int CG = AccShape<C, G>::CG;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float lo, hi, sim;
        if (CG == 1) {
            unpack2(acc[g / 2], lo, hi);
            sim = (g & 1) ? hi : lo;
        } else {
            unpack2(acc[g], lo, hi);
            sim = __fmul_rn(__fadd_rn(lo, hi), 1.0f / CG);
        }
        consume(g, sim);
    }
}

// A live sample whose cell is not inside the staged box (depth discontinuities, wild hypotheses, a
// footprint larger than the box): the same arithmetic on 256-bit global loads.  Rare, hence out of line;
// the reference features are re-read from global memory so that no register array crosses the call.
template <int C, int G>
__device__ __noinline__ void gather_cell(const float* __restrict__ src, const float* __restrict__ refp, int x0, int y0, float ex,
                                         float dx, float ey, float dy, int H, int W, float* __restrict__ sims) {
    constexpr int NACC = AccShape<C, G>::NACC, PPA = AccShape<C, G>::PPA, CG = AccShape<C, G>::CG;
    // 2x2 block clamped into the image; the axis weights move with it (a corner outside the image
    // gets weight zero, the inside one keeps upstream's weight)
    const int xc = min(max(x0, 0), W - 2), yc = min(max(y0, 0), H - 2);
    const float wl = x0 == xc ? ex : (x0 + 1 == xc ? dx : 0.0f);
    const float wr = x0 == xc ? dx : (x0 == xc + 1 ? ex : 0.0f);
    const float wt = y0 == yc ? ey : (y0 + 1 == yc ? dy : 0.0f);
    const float wb = y0 == yc ? dy : (y0 == yc + 1 ? ey : 0.0f);
    const float w00 = __fmul_rn(wl, wt), w01 = __fmul_rn(wr, wt), w10 = __fmul_rn(wl, wb), w11 = __fmul_rn(wr, wb);
    const u64 W00 = pack2(w00, w00), W01 = pack2(w01, w01), W10 = pack2(w10, w10), W11 = pack2(w11, w11);
    const float* p = src + ((size_t)yc * W + xc) * C;
    u64 acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0ull;
    static_for<0, C / 8>([&](auto cb_c) {
        constexpr int cb = decltype(cb_c)::value;
        u64 t00[4], t01[4], t10[4], t11[4], r2[4];
        O2 o;
        o = ldg_o2(p + cb * 8);                       t00[0] = o.a; t00[1] = o.b; t00[2] = o.c; t00[3] = o.d;
        o = ldg_o2(p + C + cb * 8);                   t01[0] = o.a; t01[1] = o.b; t01[2] = o.c; t01[3] = o.d;
        o = ldg_o2(p + (size_t)W * C + cb * 8);       t10[0] = o.a; t10[1] = o.b; t10[2] = o.c; t10[3] = o.d;
        o = ldg_o2(p + (size_t)(W + 1) * C + cb * 8); t11[0] = o.a; t11[1] = o.b; t11[2] = o.c; t11[3] = o.d;
        o = ldg_o2(refp + cb * 8);                    r2[0] = o.a; r2[1] = o.b; r2[2] = o.c; r2[3] = o.d;
        accumulate8<NACC, (CG == 1 ? cb * 4 : (cb * 8) / CG), PPA>(t00, t01, t10, t11, W00, W01, W10, W11, r2, acc);
    });
    emit_sims<C, G>(acc, [&](int g, float sim) { sims[g] = sim; });
}

// One round (a source view, or a plane group of one) of a block:
//   1. the two end planes of every pixel give its part of the block's source bounding box (the sample
//      position is a Moebius function of the depth, monotonic along the epipolar line between planes on
//      the same side of the camera); warp reduction + shared atomics, one block barrier;
//   2. one thread issues the TMA loads of the box (anchored at its top-left corner; whatever extends past
//      BW x BH is simply not staged);
//   3. in the shadow of the copy the threads compute the exact sample positions of all their planes;
//   4. per sample: cell inside the staged box -> shared memory, else -> gather_cell().  The box is an
//      optimisation only: correctness never depends on step 1 having predicted it right.
// consume(k, g, sim) receives the correlation of plane k and group g.
template <int C, int G, int DPT, class F>
__device__ __forceinline__ void run_round(TileShared& sh, const uint8_t* __restrict__ tile, const CUtensorMap* map,
                                          const float* __restrict__ src, const float* __restrict__ refp, int round, uint32_t& phase,
                                          int flags, int b, const Ray& ray, const float (&depth)[DPT], unsigned valid, int H, int W,
                                          const SegConsts& kc, const u64* __restrict__ ref2, F&& consume) {
    constexpr int BW = Box<C>::W, BH = Box<C>::H, ROW = BoxBytes<C>::ROW, SUB = BoxBytes<C>::SUB;
    constexpr int NACC = AccShape<C, G>::NACC, PPA = AccShape<C, G>::PPA, CG = AccShape<C, G>::CG;
    const int lane = threadIdx.x & 31;
    float ix[DPT], iy[DPT];
    const Ray2 ray2 = make_ray2(ray);
    float jx, jy;                          // position of the last existing plane

    // ---- 1. bounding box from the end planes
    {
        float dl = depth[0];
#pragma unroll
        for (int k = 1; k < DPT; ++k) dl = ((valid >> k) & 1u) ? depth[k] : dl;
        sample_coords2(ray2, depth[0], dl, kc, ix[0], iy[0], jx, jy);
        const float lx = floorf(fminf(ix[0], jx)), hx = floorf(fmaxf(ix[0], jx));
        const float ly = floorf(fminf(iy[0], jy)), hy = floorf(fmaxf(iy[0], jy));
        // the curve touches the image extended by the -1 border (false for NaN/inf and for absent pixels)
        const bool on = (valid & 1u) && (lx <= (float)(W - 1)) && (hx >= -1.0f) && (ly <= (float)(H - 1)) && (hy >= -1.0f) &&
                        (fabsf(lx) < 1e9f) && (fabsf(hx) < 1e9f) && (fabsf(ly) < 1e9f) && (fabsf(hy) < 1e9f);
        int mnx = on ? (int)fmaxf(lx, -1.0f) : INT_MAX, mxx = on ? (int)fminf(hx, (float)(W - 1)) : INT_MIN;
        int mny = on ? (int)fmaxf(ly, -1.0f) : INT_MAX, mxy = on ? (int)fminf(hy, (float)(H - 1)) : INT_MIN;
        mnx = __reduce_min_sync(0xffffffffu, mnx);
        mny = __reduce_min_sync(0xffffffffu, mny);
        mxx = __reduce_max_sync(0xffffffffu, mxx);
        mxy = __reduce_max_sync(0xffffffffu, mxy);
        int* box = sh.box[round & 1];
        if (lane == 0 && mnx <= mxx) {
            atomicMin(&box[0], mnx);
            atomicMin(&box[1], mny);
            atomicMax(&box[2], mxx);
            atomicMax(&box[3], mxy);
        }
    }
    __syncthreads();   // box complete; every thread is also done with the tile of the previous round
    const int bx = sh.box[round & 1][0], by = sh.box[round & 1][1];
    const bool any = sh.box[round & 1][0] <= sh.box[round & 1][2] && !(flags & FLAG_FORCE_GATHER);

    // ---- 2. stage the box
    if (threadIdx.x == 0) {
        int* nb = sh.box[(round + 1) & 1];   // last read in the previous round: reset it for the next one
        nb[0] = INT_MAX; nb[1] = INT_MAX; nb[2] = INT_MIN; nb[3] = INT_MIN;
        if (any) {
            mbar_expect_tx(&sh.bar, BoxBytes<C>::ALL);
#pragma unroll
            for (int cb = 0; cb < C / 8; ++cb) tma_load_box(smem_u32(tile) + cb * SUB, map, &sh.bar, cb * 8, bx, by, b);
        }
    }

    // ---- 3. exact positions of all planes (NaN for planes / pixels that do not exist)
    //         planes 1 .. DPT-2 in pairs; plane DPT-1, where it exists, is the last existing plane of step 1
    static_assert(DPT % 2 == 0, "planes are paired");
    if (!(valid & 1u)) ix[0] = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 1; k + 1 < DPT; k += 2) {
        sample_coords2(ray2, depth[k], depth[k + 1], kc, ix[k], iy[k], ix[k + 1], iy[k + 1]);
        if (!((valid >> k) & 1u)) ix[k] = __int_as_float(0x7fc00000);
        if (!((valid >> (k + 1)) & 1u)) ix[k + 1] = __int_as_float(0x7fc00000);
    }
    ix[DPT - 1] = ((valid >> (DPT - 1)) & 1u) ? jx : __int_as_float(0x7fc00000);
    iy[DPT - 1] = jy;
    if (any) {
        mbar_wait(&sh.bar, phase);
        phase ^= 1u;
    }
    __syncwarp();   // lanes leave the wait loop one by one: reconverge before the long sampling code

    // ---- 4. sample
    const int org = by * BW + bx;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const Cell c = cell_at(ix[k], iy[k], H, W);
        u64 acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0ull;
        const bool staged = any && (unsigned)(c.x0 - bx) <= (unsigned)(BW - 2) && (unsigned)(c.y0 - by) <= (unsigned)(BH - 2);
        if (c.live && staged) {
            const uint32_t a0 = (uint32_t)(c.y0 * BW + c.x0 - org) * 32u, a1 = a0 + 32u;
            // 32-byte swizzle: address bit 4 ^= bit 7 (the two 16-byte halves of a pixel swap in every other 128-byte line)
            const uint32_t s0 = a0 ^ ((a0 >> 3) & 16u), s1 = a1 ^ ((a1 >> 3) & 16u);
            const uint8_t* p00 = tile + s0;
            const uint8_t* p01 = tile + s1;
            const uint8_t* q00 = tile + (s0 ^ 16u);
            const uint8_t* q01 = tile + (s1 ^ 16u);
            const float w00 = __fmul_rn(c.ex, c.ey), w01 = __fmul_rn(c.dx, c.ey), w10 = __fmul_rn(c.ex, c.dy), w11 = __fmul_rn(c.dx, c.dy);
            const u64 W00 = pack2(w00, w00), W01 = pack2(w01, w01), W10 = pack2(w10, w10), W11 = pack2(w11, w11);
            static_for<0, C / 8>([&](auto cb_c) {
                constexpr int cb = decltype(cb_c)::value;
                u64 t00[4], t01[4], t10[4], t11[4];
                ulonglong2 q;
                q = *reinterpret_cast<const ulonglong2*>(p00 + cb * SUB);       t00[0] = q.x; t00[1] = q.y;
                q = *reinterpret_cast<const ulonglong2*>(q00 + cb * SUB);       t00[2] = q.x; t00[3] = q.y;
                q = *reinterpret_cast<const ulonglong2*>(p01 + cb * SUB);       t01[0] = q.x; t01[1] = q.y;
                q = *reinterpret_cast<const ulonglong2*>(q01 + cb * SUB);       t01[2] = q.x; t01[3] = q.y;
                q = *reinterpret_cast<const ulonglong2*>(p00 + cb * SUB + ROW); t10[0] = q.x; t10[1] = q.y;
                q = *reinterpret_cast<const ulonglong2*>(q00 + cb * SUB + ROW); t10[2] = q.x; t10[3] = q.y;
                q = *reinterpret_cast<const ulonglong2*>(p01 + cb * SUB + ROW); t11[0] = q.x; t11[1] = q.y;
                q = *reinterpret_cast<const ulonglong2*>(q01 + cb * SUB + ROW); t11[2] = q.x; t11[3] = q.y;
                accumulate8<NACC, (CG == 1 ? cb * 4 : (cb * 8) / CG), PPA>(t00, t01, t10, t11, W00, W01, W10, W11, ref2 + cb * 4, acc);
            });
            emit_sims<C, G>(acc, [&](int g, float sim) { consume(k, g, sim); });
        } else if (c.live) {
            float sims[G];
            gather_cell<C, G>(src, refp, c.x0, c.y0, c.ex, c.dx, c.ey, c.dy, H, W, sims);
#pragma unroll
            for (int g = 0; g < G; ++g) consume(k, g, sims[g]);
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) consume(k, g, 0.0f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Segment form.  The D planes of one (pixel, source view) sample ONE epipolar segment, and in the cascade's
// local volumes (and in a stage-1 plane group) neighbouring planes are a fraction of a source pixel apart: the
// D x 4 bilinear taps keep hitting the same few source pixels.  Bilinear interpolation commutes with the channel
// dot product,
//     sim(p, d) = sum_taps w_t <ref_p, src_{q_t}> / C,
// so a thread first forms g(q) = <ref_p, src_q> ONCE for every source pixel q of the segment's footprint (the
// bounding box of its cells: FX x FY pixels, a handful), parks the scalars in a thread-private strip of shared
// memory (slot-major, so bank = lane whatever the slot: conflict free), and then every plane costs four 4-byte
// reads and three blends instead of 4 x C channel reads and a C-wide interpolation: ~3x fewer instructions and
// ~F / (4 D) of the tap traffic (C = 8, D = 8, F = 8: a quarter).
// With that little traffic the taps no longer need staging: the footprint pixels are read straight from global
// memory with 256-bit loads (neighbouring lanes read neighbouring 32-byte pixels: full sectors, L1 hits for the
// overlap between lanes and planes).  No TMA box, no block-wide bounding-box negotiation, no barrier per source
// view -- measured, the staged variant of this kernel was latency-bound on exactly that round trip (0.128 ms
// whatever the instruction count) -- and far less L2 -> SM traffic than a 64 x 12 box per (block, view).
// A thread whose footprint does not fit SEG_SLOTS (planes far apart: plane sweeps over the full depth range)
// samples plane by plane with gather_cell().  Pixels outside the image contribute zero (grid_sample's padding).
// ------------------------------------------------------------------------------------------------
constexpr int SEG_FX = 6, SEG_FY = 5;      // footprint capacity of a thread: columns (unrolled, predicated) x rows (a loop)
constexpr int SEG_SLOTS = SEG_FX * SEG_FY;
#ifndef EFFI_SEG_BPS8
#define EFFI_SEG_BPS8 6
#endif
#ifndef EFFI_SEG_BPS16
#define EFFI_SEG_BPS16 5
#endif
#ifndef EFFI_SEG_BPS32
#define EFFI_SEG_BPS32 4
#endif
template <int C> struct SegBlocksPerSM { static constexpr int value = C == 8 ? EFFI_SEG_BPS8 : (C == 16 ? EFFI_SEG_BPS16 : EFFI_SEG_BPS32); };

// Sample position without upstream's normalise / un-normalise round trip and with one Newton-refined reciprocal
// instead of two IEEE divisions: 8 instructions instead of ~36, ~2 ulp of a coordinate away from upstream's value
// (1e-4 px at x ~ 1000).  That is the rounding noise of upstream's own chain, but on white-noise features it moves the
// similarity by up to 2e-4 of its range -- outside the 1e-4 parity bar -- so it is opt-in (EFFIMVS_WARP_FAST_COORDS=1).
__device__ __forceinline__ void sample_coords_fast(const Ray& r, float depth, float& ix, float& iy) {
    const float px = fmaf(r.rx, depth, r.tx), py = fmaf(r.ry, depth, r.ty);
    float pz = fmaf(r.rz, depth, r.tz);
    if (pz == 0.0f) pz = 1e-8f;
    float q;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(pz));
    q = fmaf(q, fmaf(-pz, q, 1.0f), q);
    ix = px * q;
    iy = py * q;
}
template <bool EXACT>
__device__ __forceinline__ void seg_coords(const Ray& r, float depth, int H, int W, float ihw, float ihh, float& ix, float& iy) {
    if (EXACT) sample_coords(r, depth, H, W, ihw, ihh, ix, iy);
    else sample_coords_fast(r, depth, ix, iy);
}

// One source view of one thread in segment form; G = 1.  nk = number of planes of this thread.
// consume(k, sim) receives the correlation of plane k (mean over the C channels: ref2 holds the reference features
// already divided by C -- a power of two, so every product and sum scales exactly).  strip = this thread's column of the
// block's [SEG_SLOTS][STRIDE] scratch; slot = row * SEG_FX + column (constant strides: immediate offsets).
template <int C, int DPT, bool EXACT, int STRIDE, class F>
__device__ __forceinline__ void run_view_seg(float* __restrict__ strip, const float* __restrict__ src, const float* __restrict__ refp,
                                             const Ray& ray, const float (&depth)[DPT], int nk, int H, int W, float inv_half_w,
                                             float inv_half_h, const u64* __restrict__ ref2, int flags, F&& consume) {
    constexpr int PB = C == 32 ? 1 : 2;      // pixels whose loads are in flight together (16 / 32 / 32 registers)
    float ix[DPT], iy[DPT];
    // end planes first: they fix the footprint, whose lines are then requested (prefetch: no destination registers) while
    // the coordinates of the planes in between are computed -- otherwise every row of the footprint is a dependent
    // L2 / DRAM round trip of its own and the kernel idles on long-scoreboard stalls
    seg_coords<EXACT>(ray, depth[0], H, W, inv_half_w, inv_half_h, ix[0], iy[0]);
    seg_coords<EXACT>(ray, depth[DPT - 1], H, W, inv_half_w, inv_half_h, ix[DPT - 1], iy[DPT - 1]);
    float jx = ix[DPT - 1], jy = iy[DPT - 1];
    if (nk < DPT) {       // a partial plane group (block-uniform): its last plane
        jx = ix[0]; jy = iy[0];
#pragma unroll
        for (int k = 1; k < DPT - 1; ++k)
            if (k == nk - 1) seg_coords<EXACT>(ray, depth[k], H, W, inv_half_w, inv_half_h, jx, jy);
    }

    // ---- 1. footprint of the segment from its end planes (the position is a Moebius function of the depth: monotonic in
    //         x and in y between two planes on the same side of the camera; a plane that falls outside it anyway -- a
    //         hypothesis list that is not monotonic, a pole between the planes -- is sampled on its own below)
    const float lxf = floorf(fminf(ix[0], jx)), hxf = floorf(fmaxf(ix[0], jx));
    const float lyf = floorf(fminf(iy[0], jy)), hyf = floorf(fmaxf(iy[0], jy));
    bool on = (lxf <= (float)(W - 1)) && (hxf >= -1.0f) && (lyf <= (float)(H - 1)) && (hyf >= -1.0f) &&
              (fabsf(lxf) < 1e9f) && (fabsf(hxf) < 1e9f) && (fabsf(lyf) < 1e9f) && (fabsf(hyf) < 1e9f);   // false for NaN/inf
    // cells lx..hx x ly..hy, clamped to the image extended by its zero border: every cell inside is a live sample
    const int lx = on ? (int)fmaxf(lxf, -1.0f) : 0, ly = on ? (int)fmaxf(lyf, -1.0f) : 0;
    const int FX = on ? (int)fminf(hxf, (float)(W - 1)) - lx + 2 : 0, FY = on ? (int)fminf(hyf, (float)(H - 1)) - ly + 2 : 0;
    on = on && FX <= SEG_FX && FY <= SEG_FY;
    if (on && !(flags & FLAG_DBG_NO_PREFETCH)) {
        // first and last pixel of every footprint row: with 32-byte pixels (C = 8) a row of up to 6 pixels spans at most two
        // 128-byte lines; wider pixels get one request per pixel pair
        const int xa = min(max(lx, 0), W - 1), xb = min(max(lx + FX - 1, 0), W - 1);
        for (int j = 0; j < FY; ++j) {
            const float* rowp = src + min(max(ly + j, 0), H - 1) * (W * C);
            if (C == 8) {
                prefetch_l1(rowp + xa * C);
                prefetch_l1(rowp + xb * C + C - 1);
            } else {
                for (int x = xa; x <= xb; x += 128 / (C * 4) > 0 ? 128 / (C * 4) : 1) prefetch_l1(rowp + x * C);
                prefetch_l1(rowp + xb * C + C - 1);
            }
        }
    }
#pragma unroll
    for (int k = 1; k < DPT - 1; ++k) seg_coords<EXACT>(ray, depth[k], H, W, inv_half_w, inv_half_h, ix[k], iy[k]);

    // ---- 2. g(q) = <ref, src_q> / C for every source pixel of the footprint -> strip[row * SEG_FX + column].  Pixels
    //         outside the image are read at the clamped address and replaced by zero (grid_sample's zeros padding).
    if (on) {
        int coff[SEG_FX];
        unsigned colv = 0;
#pragma unroll
        for (int i = 0; i < SEG_FX; ++i) {
            const int x = lx + i;
            coff[i] = min(max(x, 0), W - 1) * C;
            colv |= ((unsigned)x < (unsigned)W) ? (1u << i) : 0u;
        }
        float* out = strip;
        for (int j = 0; j < FY; ++j, out += SEG_FX * STRIDE) {
            const int y = ly + j;
            const bool rowv = (unsigned)y < (unsigned)H;
            const float* rowp = src + min(max(y, 0), H - 1) * (W * C);
#pragma unroll
            for (int i0 = 0; i0 < SEG_FX; i0 += PB) {
                if (i0 < FX && !(flags & FLAG_DBG_NO_LOADS)) {
                    O2 px[PB][C / 8];
#pragma unroll
                    for (int i = 0; i < PB; ++i)
#pragma unroll
                        for (int cb = 0; cb < C / 8; ++cb) px[i][cb] = ldg_o2(rowp + coff[i0 + i] + cb * 8);
#pragma unroll
                    for (int i = 0; i < PB; ++i) {
                        u64 m0 = mul2(px[i][0].a, ref2[0]), m1 = mul2(px[i][0].b, ref2[1]);
                        m0 = fma2(px[i][0].c, ref2[2], m0);
                        m1 = fma2(px[i][0].d, ref2[3], m1);
#pragma unroll
                        for (int cb = 1; cb < C / 8; ++cb) {
                            m0 = fma2(px[i][cb].a, ref2[cb * 4], m0);
                            m1 = fma2(px[i][cb].b, ref2[cb * 4 + 1], m1);
                            m0 = fma2(px[i][cb].c, ref2[cb * 4 + 2], m0);
                            m1 = fma2(px[i][cb].d, ref2[cb * 4 + 3], m1);
                        }
                        float a0, a1, b0, b1;
                        unpack2(m0, a0, a1);
                        unpack2(m1, b0, b1);
                        const float g = (a0 + a1) + (b0 + b1);
                        out[(i0 + i) * STRIDE] = (rowv && ((colv >> (i0 + i)) & 1u)) ? g : 0.0f;
                    }
                }
            }
        }
    }

    // ---- 3. the planes: four scalars and three blends each
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        float sim = 0.0f;
        if (k < nk && !(flags & FLAG_DBG_NO_SAMPLES)) {
            const float fx = floorf(ix[k]), fy = floorf(iy[k]);
            const int x0 = (int)fx, y0 = (int)fy;
            const float dx = __fsub_rn(ix[k], fx), dy = __fsub_rn(iy[k], fy);
            const float ex = __fsub_rn(__fadd_rn(fx, 1.0f), ix[k]), ey = __fsub_rn(__fadd_rn(fy, 1.0f), iy[k]);
            const unsigned ox = (unsigned)(x0 - lx), oy = (unsigned)(y0 - ly);
            if (on && ox <= (unsigned)(FX - 2) && oy <= (unsigned)(FY - 2) && fabsf(fx) < 1e9f && fabsf(fy) < 1e9f) {
                const float* g = strip + (oy * SEG_FX + ox) * STRIDE;
                const float g00 = g[0], g01 = g[STRIDE], g10 = g[SEG_FX * STRIDE], g11 = g[(SEG_FX + 1) * STRIDE];
                sim = fmaf(fmaf(g11, dx, g10 * ex), dy, fmaf(g01, dx, g00 * ex) * ey);
            } else if ((fx >= -1.0f) && (fx <= (float)(W - 1)) && (fy >= -1.0f) && (fy <= (float)(H - 1))) {   // live (false for NaN/inf)
                float s1[1];
                gather_cell<C, 1>(src, refp, x0, y0, ex, dx, ey, dy, H, W, s1);
                sim = s1[0];
            }
        }
        consume(k, sim);
    }
}

// num[k] / div for all k, IEEE round-to-nearest, one shared reciprocal (see div2_rn in common.cuh)
template <int N>
__device__ __forceinline__ void div_all_rn(float (&num)[N], float div) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(div));
    r = fmaf(r, fmaf(-div, r, 1.0f), r);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        float q = __fmul_rn(num[k], r);
        q = fmaf(fmaf(-div, q, num[k]), r, q);
        num[k] = fmaf(fmaf(-div, q, num[k]), r, q);
    }
}

// reference features of a pixel, divided by C (exact: C is a power of two), as packed pairs
template <int C>
__device__ __forceinline__ void load_ref2_scaled(const float* __restrict__ p, u64 (&ref2)[C / 2]) {
    const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(p);
    const u64 sc = pack2(1.0f / C, 1.0f / C);
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
        const ulonglong2 v = __ldg(rp + q);
        ref2[2 * q] = mul2(v.x, sc);
        ref2[2 * q + 1] = mul2(v.y, sc);
    }
}

__device__ __forceinline__ const uint8_t* block_prologue(TileShared& sh, uint8_t* smem_raw, const float* __restrict__ proj, int n_proj) {
    const int tid = threadIdx.x;
    for (int i = tid; i < n_proj; i += TILE_THREADS) sh.P[i] = proj[i];
    if (tid < 8) sh.box[tid >> 2][tid & 3] = (tid & 2) ? INT_MIN : INT_MAX;
    if (tid == 0) {
        mbar_init(&sh.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    return smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
}

template <int C>
__device__ __forceinline__ void load_ref2(const float* __restrict__ p, u64 (&ref2)[C / 2]) {
    const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(p);
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
        const ulonglong2 v = __ldg(rp + q);
        ref2[2 * q] = v.x;
        ref2[2 * q + 1] = v.y;
    }
}

// ------------------------------------------------------------------------------------------------
// aggregated form (stages 2 / 3, config-2 microbench): a round per source view, all views in one block
// ------------------------------------------------------------------------------------------------
template <int C, int G>
__global__ void __launch_bounds__(TILE_THREADS, BlocksPerSM<C>::value)
warp_corr_tile_kernel(const __grid_constant__ TileMaps maps, const float* __restrict__ ref_fea,
                      const __grid_constant__ SrcPtrs srcs, int n_src, const float* __restrict__ proj, const float* __restrict__ hyp, int hyp_mode,
                      const float* __restrict__ interval, const float* __restrict__ weights, int H, int W, int D, int tiles_x,
                      int flags, const SegConsts kc, float* __restrict__ sim_out, float* __restrict__ hyp_out) {
    pdl_enter();
    constexpr int DPT = TilePlanes<G>::value;
    extern __shared__ uint8_t smem_raw[];
    __shared__ TileShared sh;

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z;
    const int HW = H * W;
    const uint8_t* tile = block_prologue(sh, smem_raw, proj + (size_t)b * n_src * 12, n_src * 12);

    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int xi = tx * TW + lane, yi = ty * TH + (tid >> 5);
    const bool inimg = xi < W && yi < H;
    const int pix = inimg ? yi * W + xi : 0;
    const int d0 = blockIdx.y * DPT;
    const float x = (float)xi, y = (float)yi;

    const float* refp = ref_fea + ((size_t)b * HW + pix) * C;
    u64 ref2[C / 2];
    load_ref2<C>(refp, ref2);
    float depth[DPT], num[DPT][G];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < DPT; ++k) valid |= (inimg && d0 + k < D) ? (1u << k) : 0u;
    if (hyp_mode == EFFIMVS_HYP_LOCAL) {          // the per-pixel part of the hypotheses once, not once per plane
        const LocalHyp lh = local_hypothesis_prepare(__ldg(hyp + (size_t)b * HW + pix), __ldg(interval + b), D, kc.inv_dm1);
#pragma unroll
        for (int k = 0; k < DPT; ++k) depth[k] = ((valid >> k) & 1u) ? local_hypothesis_at(lh, d0 + k) : 1.0f;
    } else {
#pragma unroll
        for (int k = 0; k < DPT; ++k) depth[k] = ((valid >> k) & 1u) ? fetch_hypothesis(hyp, hyp_mode, interval, b, d0 + k, D, pix, HW) : 1.0f;
    }
    if (hyp_out) {
        float* ho = hyp_out + ((size_t)b * D + d0) * HW + pix;
#pragma unroll
        for (int k = 0; k < DPT; ++k, ho += HW)
            if ((valid >> k) & 1u) *ho = depth[k];
    }
#pragma unroll
    for (int k = 0; k < DPT; ++k)
#pragma unroll
        for (int g = 0; g < G; ++g) num[k][g] = 0.0f;
    float den = 0.0f;
    uint32_t phase = 0;
    const float* wp = weights ? weights + (size_t)b * n_src * HW + pix : nullptr;

    for (int v = 0; v < n_src; ++v) {
        const Ray ray = make_ray(sh.P + v * 12, x, y, (flags & FLAG_RAY_UNFUSED) != 0);
        const float w = (wp && inimg) ? __ldg(wp + (size_t)v * HW) : 1.0f;
        run_round<C, G, DPT>(sh, tile, &maps.m[v], srcs.p[v] + (size_t)b * C * HW, refp, v, phase, flags, b, ray, depth, valid, H, W,
                             kc, ref2, [&](int k, int g, float sim) {
                                 num[k][g] = wp ? __fadd_rn(num[k][g], __fmul_rn(sim, w)) : __fadd_rn(num[k][g], sim);
                             });
        den = __fadd_rn(den, w);
    }
    if (!inimg) return;
    // num / div for every plane and group: IEEE round-to-nearest quotients from one shared reciprocal (div2_rn's sequence)
    const float div = wp ? __fadd_rn(den, 1e-6f) : (float)n_src;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(div));
    r = fmaf(r, fmaf(-div, r, 1.0f), r);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float* so = sim_out + (((size_t)b * G + g) * D + d0) * HW + pix;
#pragma unroll
        for (int k = 0; k < DPT; ++k, so += HW)
            if ((valid >> k) & 1u) {
                float q = __fmul_rn(num[k][g], r);
                q = fmaf(fmaf(-div, q, num[k][g]), r, q);
                *so = fmaf(fmaf(-div, q, num[k][g]), r, q);
            }
    }
}

// ------------------------------------------------------------------------------------------------
// stage-1 form: per-view similarities (no aggregation); a block = (tile, group of 8 planes, source
// view), one round.  The softmax entropy over D is a second, streaming kernel over the similarities.
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(TILE_THREADS, BlocksPerSM<C>::value)
warp_views_tile_kernel(const __grid_constant__ TileMaps maps, const float* __restrict__ ref_fea, const __grid_constant__ SrcPtrs srcs,
                       int n_src, const float* __restrict__ proj, const float* __restrict__ hyp, int hyp_mode, int H, int W, int D,
                       int tiles_x, int flags, const SegConsts kc, float* __restrict__ sims_out) {
    pdl_enter();
    constexpr int DPT = 8;
    extern __shared__ uint8_t smem_raw[];
    __shared__ TileShared sh;

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z, v = blockIdx.y % n_src, d0 = (blockIdx.y / n_src) * DPT;
    const int HW = H * W;
    const uint8_t* tile = block_prologue(sh, smem_raw, proj + ((size_t)b * n_src + v) * 12, 12);

    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int xi = tx * TW + lane, yi = ty * TH + (tid >> 5);
    const bool inimg = xi < W && yi < H;
    const int pix = inimg ? yi * W + xi : 0;

    const float* refp = ref_fea + ((size_t)b * HW + pix) * C;
    u64 ref2[C / 2];
    load_ref2<C>(refp, ref2);
    const Ray ray = make_ray(sh.P, (float)xi, (float)yi, (flags & FLAG_RAY_UNFUSED) != 0);
    float* out = sims_out + (((size_t)b * n_src + v) * D) * HW + pix;
    uint32_t phase = 0;

    float depth[DPT];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const bool on = inimg && d0 + k < D;
        valid |= on ? (1u << k) : 0u;
        depth[k] = on ? fetch_hypothesis(hyp, hyp_mode, nullptr, b, d0 + k, D, pix, HW) : 1.0f;
    }
    run_round<C, 1, DPT>(sh, tile, &maps.m[v], srcs.p[v] + (size_t)b * C * HW, refp, 0, phase, flags, b, ray, depth, valid, H, W,
                         kc, ref2, [&](int k, int, float sim) {
                             if ((valid >> k) & 1u) out[(size_t)(d0 + k) * HW] = sim;
                         });
}

// segment-form twins of the two kernels above (G = 1): no staging, no barriers.

// Aggregated form: one THREAD per (reference pixel, source view).  A block is a 32 x SEGV_TY pixel tile times the n_src
// views (threadIdx = lane / row / view), so a pixel's views run in parallel warps instead of one after the other in one
// thread: n_src times the loads in flight, a third fewer registers per thread, ~60 % occupancy -- the version with a
// view loop per thread was latency-bound (half of the issue slots idle on L1 / L2 loads at 24 warps per SM).  The
// per-pixel work is split over the view-threads through shared memory: each computes the hypotheses of planes
// view, view + n_src, ... (and writes hyp_out), and after the per-view similarities are parked in shared memory each
// forms the weighted aggregate of those planes in upstream's view order.
constexpr int SEGV_TY = 2, SEGV_PX = TW * SEGV_TY;
template <int NT> struct SegvBlocksPerSM { static constexpr int value = NT <= 256 ? 4 : (NT <= 512 ? 2 : 1); };
constexpr size_t segv_smem(int NT, int n_src) {
    return sizeof(float) * ((size_t)SEG_SLOTS * NT + (size_t)SEGV_PX * (8 + n_src + 8 * n_src));
}

template <int C, bool EXACT, int NT>
__global__ void __launch_bounds__(NT, SegvBlocksPerSM<NT>::value)
warp_corr_seg_kernel(const float* __restrict__ ref_fea, const __grid_constant__ SrcPtrs srcs, int n_src, const float* __restrict__ proj,
                     const float* __restrict__ hyp, int hyp_mode, const float* __restrict__ interval, const float* __restrict__ weights,
                     int H, int W, int D, int tiles_x, int flags, const SegConsts kc, float* __restrict__ sim_out,
                     float* __restrict__ hyp_out) {
    pdl_enter();
    constexpr int DPT = 8;
    extern __shared__ float seg_smem[];
    float* const scratch = seg_smem;                                   // [SEG_SLOTS][NT]
    float* const depth_s = scratch + SEG_SLOTS * NT;                   // [DPT][SEGV_PX]
    float* const wts_s = depth_s + DPT * SEGV_PX;                      // [n_src][SEGV_PX]
    float* const sims_s = wts_s + n_src * SEGV_PX;                     // [n_src][DPT][SEGV_PX]
    const int lane = threadIdx.x, row = threadIdx.y, view = threadIdx.z;
    const int p = row * TW + lane;                                     // pixel of the tile
    const int tid = (view * SEGV_TY + row) * TW + lane;
    const int b = blockIdx.z;
    const int HW = H * W;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int xi = tx * TW + lane, yi = ty * SEGV_TY + row;
    const bool inimg = xi < W && yi < H;
    const int pix = inimg ? yi * W + xi : 0;
    const int d0 = blockIdx.y * DPT;
    const int nk = min(DPT, D - d0);

    // ---- A. hypotheses of planes view, view + n_src, ... -> shared memory (+ hyp_out)
    if (inimg) {
        LocalHyp lh;
        if (hyp_mode == EFFIMVS_HYP_LOCAL) lh = local_hypothesis_prepare(__ldg(hyp + (size_t)b * HW + pix), __ldg(interval + b), D, kc.inv_dm1);
        for (int k = view; k < nk; k += n_src) {
            const float d = hyp_mode == EFFIMVS_HYP_LOCAL ? local_hypothesis_at(lh, d0 + k)
                                                          : fetch_hypothesis(hyp, hyp_mode, interval, b, d0 + k, D, pix, HW);
            depth_s[k * SEGV_PX + p] = d;
            if (hyp_out) hyp_out[((size_t)b * D + d0 + k) * HW + pix] = d;
        }
    }
    __syncthreads();

    // ---- B. this thread's view: similarities of its planes -> shared memory
    if (inimg) {
        const float* P = proj + ((size_t)b * n_src + view) * 12;
        float Pr[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) Pr[i] = __ldg(P + i);
        const Ray ray = make_ray(Pr, (float)xi, (float)yi, (flags & FLAG_RAY_UNFUSED) != 0);
        const float* refp = ref_fea + ((size_t)b * HW + pix) * C;
        u64 ref2[C / 2];
        load_ref2_scaled<C>(refp, ref2);
        float depth[DPT];
#pragma unroll
        for (int k = 0; k < DPT; ++k) depth[k] = k < nk ? depth_s[k * SEGV_PX + p] : 1.0f;
        wts_s[view * SEGV_PX + p] = weights ? __ldg(weights + ((size_t)b * n_src + view) * HW + pix) : 1.0f;
        float* mine = sims_s + view * DPT * SEGV_PX + p;
        run_view_seg<C, DPT, EXACT, NT>(scratch + tid, srcs.p[view] + (size_t)b * C * HW, refp, ray, depth, nk, H, W, kc.inv_half_w,
                                        kc.inv_half_h, ref2, flags, [&](int k, float sim) { mine[k * SEGV_PX] = sim; });
    }
    __syncthreads();

    // ---- C. weighted aggregate over the views (upstream's order: view 0 first) of planes view, view + n_src, ...
    if (inimg) {
        for (int k = view; k < nk; k += n_src) {
            float num = 0.0f, den = 0.0f;
            for (int u = 0; u < n_src; ++u) {
                const float sim = sims_s[(u * DPT + k) * SEGV_PX + p], w = wts_s[u * SEGV_PX + p];
                num = weights ? __fadd_rn(num, __fmul_rn(sim, w)) : __fadd_rn(num, sim);
                den = __fadd_rn(den, w);
            }
            sim_out[((size_t)b * D + d0 + k) * HW + pix] = __fdiv_rn(num, weights ? __fadd_rn(den, 1e-6f) : (float)n_src);
        }
    }
}

template <int C, bool EXACT>
__global__ void __launch_bounds__(TILE_THREADS, SegBlocksPerSM<C>::value)
warp_views_seg_kernel(const float* __restrict__ ref_fea, const __grid_constant__ SrcPtrs srcs, int n_src, const float* __restrict__ proj,
                      const float* __restrict__ hyp, int hyp_mode, int H, int W, int D, int tiles_x, int flags, const SegConsts kc,
                      float* __restrict__ sims_out) {
    pdl_enter();
    constexpr int DPT = 8;
    __shared__ float sP[12];
    __shared__ float scratch[SEG_SLOTS * TILE_THREADS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z, v = blockIdx.y % n_src, d0 = (blockIdx.y / n_src) * DPT;
    const int HW = H * W;
    if (tid < 12) sP[tid] = proj[((size_t)b * n_src + v) * 12 + tid];
    __syncthreads();
    float* strip = scratch + tid;

    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int xi = tx * TW + lane, yi = ty * TH + (tid >> 5);
    if (xi >= W || yi >= H) return;
    const int pix = yi * W + xi;
    const int nk = min(DPT, D - d0);

    const float* refp = ref_fea + ((size_t)b * HW + pix) * C;
    u64 ref2[C / 2];
    load_ref2_scaled<C>(refp, ref2);
    const Ray ray = make_ray(sP, (float)xi, (float)yi, (flags & FLAG_RAY_UNFUSED) != 0);
    float* out = sims_out + (((size_t)b * n_src + v) * D + d0) * HW + pix;
    float depth[DPT];
#pragma unroll
    for (int k = 0; k < DPT; ++k) depth[k] = k < nk ? fetch_hypothesis(hyp, hyp_mode, nullptr, b, d0 + k, D, pix, HW) : 1.0f;
    run_view_seg<C, DPT, EXACT, TILE_THREADS>(strip, srcs.p[v] + (size_t)b * C * HW, refp, ray, depth, nk, H, W, kc.inv_half_w,
                                              kc.inv_half_h, ref2, flags, [&](int k, float sim) {
                                                  if (k < nk) out[(size_t)k * HW] = sim;
                                              });
}

// softmax entropy over the D similarities of a (pixel, view) (models/Effi_MVS_plus.py:43-44); sims (N, D, HW)
__global__ void __launch_bounds__(256)
softmax_entropy_kernel(const float* __restrict__ sims, int D, int HW, float* __restrict__ entropy_out) {
    pdl_enter();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const float* mine = sims + (size_t)blockIdx.y * D * HW + pix;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(mine + (size_t)d * HW));
    float z = 0.0f;
    for (int d = 0; d < D; ++d) z += expf(__ldg(mine + (size_t)d * HW) - m);
    float ent = 0.0f;
    for (int d = 0; d < D; ++d) {
        const float p = __fdiv_rn(expf(__ldg(mine + (size_t)d * HW) - m), z);
        ent -= p * logf(p + 1e-7f);
    }
    entropy_out[(size_t)blockIdx.y * HW + pix] = ent;
}

// The same with the D similarities of a pixel held in registers (D <= DMAX, unrolled): one read of the volume instead
// of three and each exponential evaluated once; operations and their order are those of the kernel above.
template <int DMAX>
__global__ void __launch_bounds__(256)
softmax_entropy_reg_kernel(const float* __restrict__ sims, int D, int HW, float* __restrict__ entropy_out) {
    pdl_enter();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const float* mine = sims + (size_t)blockIdx.y * D * HW + pix;
    float e[DMAX];
    float m = -INFINITY;
#pragma unroll
    for (int d = 0; d < DMAX; ++d) {
        e[d] = d < D ? __ldg(mine + (size_t)d * HW) : -INFINITY;
        m = fmaxf(m, e[d]);
    }
    float z = 0.0f;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
        if (d < D) {
            e[d] = expf(e[d] - m);
            z += e[d];
        }
    float ent = 0.0f;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
        if (d < D) {
            const float p = __fdiv_rn(e[d], z);
            ent -= p * logf(p + 1e-7f);
        }
    entropy_out[(size_t)blockIdx.y * HW + pix] = ent;
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
int encode_map(CUtensorMap* m, const float* base, int B, int C, int H, int W, int bw, int bh) {
    EncodeTiledFn enc = tensor_map_encoder();
    EFFI_REQUIRE(enc, EFFIMVS_ECUDA, "warp_corr: cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {8, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EFFI_REQUIRE(r == CUDA_SUCCESS, EFFIMVS_ECUDA, "warp_corr: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EFFIMVS_OK;
}

template <int C>
int encode_maps(TileMaps& maps, const SrcPtrs& srcs, int n_src, int B, int H, int W) {
    for (int v = 0; v < n_src; ++v) {
        int rc = encode_map(&maps.m[v], srcs.p[v], B, C, H, W, Box<C>::W, Box<C>::H);
        if (rc) return rc;
    }
    for (int v = n_src; v < EFFIMVS_MAX_SRC_VIEWS; ++v) maps.m[v] = maps.m[0];
    return EFFIMVS_OK;
}

// which launches take the segment form: 0 (default) none, 1 local hypotheses (the cascade's stage-2/3 volumes), 2 every
// G = 1 launch including the stage-1 per-view kernel.  Off by default: measured on B200 at the DTU stage-3 shape the segment
// form executes about as many warp instructions as the staged plane-by-plane kernel (77-94 M against 87 M: the dot-first
// saving is eaten by the footprint bookkeeping and the upstream-exact coordinate chain, which is half of either kernel)
// and, reading its footprint through L1 in dependent batches, idles longer on long-scoreboard stalls: 0.146-0.163 ms
// against 0.150 ms, 0.135 against 0.118 ms on a smooth surface (profiles/r2_warp_segment_form.md).
int seg_mode() {
    const char* e = getenv("EFFIMVS_WARP_SEG");
    return e ? atoi(e) : 0;
}
bool seg_fast_coords() {
    const char* e = getenv("EFFIMVS_WARP_FAST_COORDS");
    return e && e[0] == '1';
}
SegConsts seg_consts(int H, int W, int D) {
    SegConsts k;
    k.inv_half_w = 1.0f / (float)((double)(W - 1) / 2.0);        // IEEE single division, as __fdiv_rn in the other kernels
    k.inv_half_h = 1.0f / (float)((double)(H - 1) / 2.0);
    k.inv_dm1 = D > 1 ? 1.0f / (float)(D - 1) : 0.0f;
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);   // exact (sample_coords folds ATen's halving the same way)
    k.ihw2 = {k.inv_half_w, k.inv_half_w};
    k.ihh2 = {k.inv_half_h, k.inv_half_h};
    k.hw2 = {hw, hw};
    k.hh2 = {hh, hh};
    k.one2 = {1.0f, 1.0f};
    return k;
}

template <int C, int G>
int launch_tile(const float* ref, const SrcPtrs& srcs, int n_src, const float* proj, const float* hyp, int hyp_mode,
                const float* interval, const float* weights, int B, int H, int W, int D, int flags, float* sim_out, float* hyp_out,
                cudaStream_t st) {
    const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
    if constexpr (G == 1) {
        const int mode = seg_mode();
        if (mode >= 2 || (mode == 1 && hyp_mode == EFFIMVS_HYP_LOCAL)) {
            const int tiles_y2 = ceil_div(H, SEGV_TY);
            const dim3 block(TW, SEGV_TY, n_src), grid(tiles_x * tiles_y2, ceil_div(D, 8), B);
            const SegConsts kc = seg_consts(H, W, D);
            const bool fast = seg_fast_coords();
#define EFFI_SEGV_LAUNCH(NT, EX)                                                                                                       \
    {                                                                                                                                  \
        const size_t smem = segv_smem(NT, n_src);                                                                                      \
        cudaFuncSetAttribute(warp_corr_seg_kernel<C, EX, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        launch_kernel(warp_corr_seg_kernel<C, EX, NT>, grid, block, smem, st, ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, H, W, D, tiles_x, \
                                                                   flags, kc, sim_out, hyp_out);                                      \
    }
            const int nt = TW * SEGV_TY * n_src;
            if (nt <= 256) { if (fast) EFFI_SEGV_LAUNCH(256, false) else EFFI_SEGV_LAUNCH(256, true) }
            else if (nt <= 512) { if (fast) EFFI_SEGV_LAUNCH(512, false) else EFFI_SEGV_LAUNCH(512, true) }
            else { if (fast) EFFI_SEGV_LAUNCH(1024, false) else EFFI_SEGV_LAUNCH(1024, true) }
#undef EFFI_SEGV_LAUNCH
            return check_launch("warp_corr_seg_kernel");
        }
    }
    TileMaps maps;
    int rc = encode_maps<C>(maps, srcs, n_src, B, H, W);
    if (rc) return rc;
    dim3 block(TILE_THREADS), grid(tiles_x * tiles_y, ceil_div(D, TilePlanes<G>::value), B);
    const size_t smem = BoxBytes<C>::ALL + 1024;
    cudaFuncSetAttribute(warp_corr_tile_kernel<C, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_kernel(warp_corr_tile_kernel<C, G>, grid, block, smem, st, maps, ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, H, W, D,
                                                            tiles_x, flags, seg_consts(H, W, D), sim_out, hyp_out);
    return check_launch("warp_corr_tile_kernel");
}

template <int C>
int tile_dispatch_g(int G, const float* ref, const SrcPtrs& srcs, int n_src, const float* proj, const float* hyp, int hyp_mode,
                    const float* interval, const float* weights, int B, int H, int W, int D, int flags, float* sim_out,
                    float* hyp_out, cudaStream_t st) {
    switch (G) {
        case 1: return launch_tile<C, 1>(ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 2: return launch_tile<C, 2>(ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 4: return launch_tile<C, 4>(ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 8: return launch_tile<C, 8>(ref, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
    }
    set_error("warp_corr_agg: G=%d not in {1,2,4,8}", G);
    return EFFIMVS_EUNSUPPORTED;
}

template <int C>
int launch_views(const float* ref, const SrcPtrs& srcs, int n_src, const float* proj, const float* hyp, int hyp_mode, int B, int H,
                 int W, int D, int flags, float* sims_out, float* entropy_out, cudaStream_t st) {
    int rc;
    const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
    dim3 block(TILE_THREADS), grid(tiles_x * tiles_y, n_src * ceil_div(D, 8), B);
    // Stage 1 keeps the TMA-staged plane-by-plane kernel by default: with C = 32 a source pixel is a 128-byte line of its own,
    // and the L1 path pays per line touched (measured at the DTU stage-1 shape: 0.38 ms in segment form from global memory
    // against 0.17 ms from the staged box).  EFFIMVS_WARP_SEG=2 forces the segment form (tests).
    if (seg_mode() >= 2) {
        if (seg_fast_coords())
            launch_kernel(warp_views_seg_kernel<C, false>, grid, block, 0, st, ref, srcs, n_src, proj, hyp, hyp_mode, H, W, D, tiles_x, flags,
                                                                    seg_consts(H, W, D), sims_out);
        else
            launch_kernel(warp_views_seg_kernel<C, true>, grid, block, 0, st, ref, srcs, n_src, proj, hyp, hyp_mode, H, W, D, tiles_x, flags,
                                                                   seg_consts(H, W, D), sims_out);
        if ((rc = check_launch("warp_views_seg_kernel"))) return rc;
    } else {
        TileMaps maps;
        if ((rc = encode_maps<C>(maps, srcs, n_src, B, H, W))) return rc;
        const size_t smem = BoxBytes<C>::ALL + 1024;
        cudaFuncSetAttribute(warp_views_tile_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_kernel(warp_views_tile_kernel<C>, grid, block, smem, st, maps, ref, srcs, n_src, proj, hyp, hyp_mode, H, W, D, tiles_x, flags,
                                                             seg_consts(H, W, D), sims_out);
        if ((rc = check_launch("warp_views_tile_kernel"))) return rc;
    }
    const dim3 egrid(ceil_div(H * W, 256), B * n_src);
    if (D <= 48) launch_kernel(softmax_entropy_reg_kernel<48>, dim3(egrid), dim3(256), 0, st, sims_out, D, H * W, entropy_out);
    else if (D <= 96) launch_kernel(softmax_entropy_reg_kernel<96>, dim3(egrid), dim3(256), 0, st, sims_out, D, H * W, entropy_out);
    else launch_kernel(softmax_entropy_kernel, dim3(egrid), dim3(256), 0, st, sims_out, D, H * W, entropy_out);
    return check_launch("softmax_entropy_kernel");
}

}  // namespace

int warp_flags_from_env(int H, int W) {
    int flags = 0;
    const char* e = getenv("EFFIMVS_WARP_FORCE_GATHER");
    if (e && e[0] == '1') flags |= FLAG_FORCE_GATHER;
    if (ray_unfused_for(H, W)) flags |= FLAG_RAY_UNFUSED;
    if (const char* d = getenv("EFFIMVS_WARP_DEBUG")) flags |= atoi(d) & (FLAG_DBG_NO_PREFETCH | FLAG_DBG_NO_LOADS | FLAG_DBG_NO_SAMPLES);
    return flags;
}

// channels-last fast path of effimvs_warp_corr_agg_f32 (arguments already validated by the caller)
int warp_corr_agg_tile(const float* ref_fea, const SrcPtrs& srcs, int n_src, const float* proj, const float* hyp, int hyp_mode,
                       const float* interval, const float* weights, int B, int C, int H, int W, int D, int G, float* sim_out,
                       float* hyp_out, cudaStream_t st) {
    const int flags = warp_flags_from_env(H, W);
    switch (C) {
        case 8: return tile_dispatch_g<8>(G, ref_fea, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 16: return tile_dispatch_g<16>(G, ref_fea, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
        case 32: return tile_dispatch_g<32>(G, ref_fea, srcs, n_src, proj, hyp, hyp_mode, interval, weights, B, H, W, D, flags, sim_out, hyp_out, st);
    }
    set_error("warp_corr_agg: C=%d not in {8,16,32}", C);
    return EFFIMVS_EUNSUPPORTED;
}

// channels-last fast path of effimvs_warp_corr_views_f32
int warp_corr_views_tile(const float* ref_fea, const SrcPtrs& srcs, int n_src, const float* proj, const float* hyp, int hyp_mode,
                         int B, int C, int H, int W, int D, float* sims_out, float* entropy_out, cudaStream_t st) {
    const int flags = warp_flags_from_env(H, W);
    switch (C) {
        case 8: return launch_views<8>(ref_fea, srcs, n_src, proj, hyp, hyp_mode, B, H, W, D, flags, sims_out, entropy_out, st);
        case 16: return launch_views<16>(ref_fea, srcs, n_src, proj, hyp, hyp_mode, B, H, W, D, flags, sims_out, entropy_out, st);
        case 32: return launch_views<32>(ref_fea, srcs, n_src, proj, hyp, hyp_mode, B, H, W, D, flags, sims_out, entropy_out, st);
    }
    set_error("warp_corr_views: C=%d not in {8,16,32}", C);
    return EFFIMVS_EUNSUPPORTED;
}

}  // namespace effimvs
