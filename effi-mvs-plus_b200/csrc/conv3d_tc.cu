// bf16 implicit-GEMM 3x3x3 (de)convolution on the 5th-generation tensor cores (tcgen05), fp32
// accumulation in tensor memory (TMEM), operands staged in shared memory by 1-D bulk TMA copies.
//
//   Conv3d / Deconv3d (+BN+ReLU)             upstream models/module.py:124-209
//   CostRegNet_2_sample_FPN3D_Fast.forward   upstream models/module.py:453-463
//   cost_up_small.forward                    upstream models/module.py:509-516
//
// Data layout ("c8 planar, padded"): an activation tensor with C channels on a (D,H,W) grid is
// stored as C/8 channel planes; a plane is a zero-padded (D+2,H+2,W+2) volume of 16-byte voxels
// (8 bf16 channels), preceded and followed by a guard.  A tile is 128 consecutive padded positions
// q = yp*Px + xp of one z plane.  For a tap (a,b,c) the 128 input voxels of the tile are again 128
// consecutive voxels, i.e. exactly the canonical K-major / no-swizzle UMMA operand (8 rows x 16
// bytes per core matrix, SBO = 128 B) at a shifted shared-memory address: the im2col matrix is
// never built, the MMA descriptors just point into the staged rows, and the x-taps (c = 0,1,2)
// share one staged row.  K = 16 per MMA is two 8-channel chunks whose distance is the descriptor's
// LBO: two taps for an 8-channel input, two channel planes otherwise.
// Stride-2 convolutions read a parity-split copy of their input (8 sub-volumes, each a padded
// half-resolution volume), which the producing layer's epilogue writes directly; transposed
// convolutions accumulate the 8 (or 4) output parity classes in separate TMEM column ranges.
// The positions of a tile that fall on the halo compute garbage that is never stored.
//
// Precision modes.  EFFIMVS_PREC_BF16: one bf16 MMA per K chunk pair.  EFFIMVS_PREC_BF16X3:
// activations and weights are carried as hi + lo bf16 pairs (x = hi + lo to 16 mantissa bits) and
// every product is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation -- fp32-grade results
// (needed for the 1e-3 depth tolerance) on the tensor cores.  The layers are bound by the A-operand
// fetch of the tensor pipe, so x_hi is streamed once against the stacked B block [w_hi ; w_lo]
// (one MMA of width 2N, the epilogue adds the two column halves) and x_lo once against w_hi: two
// A reads per product instead of three.  32-channel inputs are processed in two K phases
// (channel-plane pairs) through the same shared-memory slab.
//
// Persistent CTAs (one or two per SM) loop over tiles with warp-specialised roles: a TMA producer
// warp, an MMA issuer warp (one elected lane each) and four epilogue warps that own TMEM lanes
// 32w..32w+31 (= tile rows).  The operand slab is double buffered (full/empty mbarriers, the
// empty arrive comes from tcgen05.commit) and so is the TMEM accumulator (tmem_full / tmem_empty),
// so the loads of tile k+2, the MMAs of tile k+1 and the epilogue of tile k overlap; the packed
// weights are loaded once per CTA.
//
// Stride-1 layers with 8 or 16 input channels run as a z-sweep instead (build_conv_s1_zsweep): a unit is one
// tile in 4 consecutive output planes, a load phase stages one input plane, and every staged row is multiplied
// once against the stacked weights of the three z-taps (N = 3 x b_rows) -- a third of the MMAs and of the
// staging.  What limits all of these layers is the ~80 clk a CTA waits per MMA of this shape (operand-fetch
// latency; tools/tc_debug.py), which is why two resident CTAs are preferred over double buffering.
// The 8 -> 1 transposed convolution at the end of cost_up_small runs on CUDA cores (deconv_cout1_kernel).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace effimvs {
namespace {

constexpr int TILE_M = 128;
constexpr int SEG_VOX = 136;  // 128 + 2 (x halo) + 1 (dummy chunk) rounded up to a multiple of 8
constexpr int SEG_BYTES = SEG_VOX * 16;
constexpr int MAX_SEGS = 96;
constexpr int MAX_OPS = 160;
constexpr int MAX_BLOCKS = 56;   // packed weight blocks (one K chunk pair each)
constexpr int MAX_PHASES = 8;
constexpr int ZSWEEP_R = 4;     // output planes per unit of the z-sweep programs

enum { L_REG = 0, L_SPLIT = 1 };

struct ActLayout {
    int kind, planes, D, H, W;  // logical dims of the tensor; planes counts hi and lo halves
    int lo_off;                 // plane offset of the lo half (0: plain bf16 tensor)
    int Py, Px, guard;          // padded dims of one (sub-)volume
    long long zstride;          // Py * Px
    long long vs;               // voxels per (sub-)volume including both guards
    long long batch_stride;     // voxels per batch item
};

ActLayout make_layout(int kind, int C, int D, int H, int W, bool hilo) {
    ActLayout L;
    L.kind = kind; L.D = D; L.H = H; L.W = W;
    L.lo_off = hilo ? C / 8 : 0;
    L.planes = (C / 8) * (hilo ? 2 : 1);
    int d = D, h = H, w = W;
    if (kind == L_SPLIT) { d /= 2; h /= 2; w /= 2; }
    L.Py = h + 2; L.Px = w + 2; L.guard = L.Px + 136;  // a tile may over-read up to 129 voxels past the last plane
    L.zstride = (long long)L.Py * L.Px;
    L.vs = (long long)(d + 2) * L.zstride + 2 * L.guard;
    L.batch_stride = L.vs * L.planes * (kind == L_SPLIT ? 8 : 1);
    return L;
}
size_t layout_bytes(const ActLayout& L, int B) { return ((size_t)L.batch_stride * B * 16 + 255) & ~(size_t)255; }

__host__ __device__ inline long long act_index(const ActLayout& L, int b, int plane, int z, int y, int x) {
    if (L.kind == L_REG)
        return (long long)b * L.batch_stride + (long long)plane * L.vs + L.guard + (long long)(z + 1) * L.zstride +
               (long long)(y + 1) * L.Px + (x + 1);
    int sub = ((z & 1) << 2) | ((y & 1) << 1) | (x & 1);
    return (long long)b * L.batch_stride + (long long)(plane * 8 + sub) * L.vs + L.guard +
           (long long)((z >> 1) + 1) * L.zstride + (long long)((y >> 1) + 1) * L.Px + ((x >> 1) + 1);
}

struct Seg { long long src_off; int copy_vox; int slot; };          // slot: shared-memory segment index within its phase
struct Op { uint32_t a_off, a_lbo, b_off; uint16_t d_col; uint8_t accum, n8; };   // n8: MMA N / 8
struct Phase { int seg_begin, seg_end, op_begin, op_end, w_off, w_bytes; };
struct Block { short tap0, cb0, tap1, cb1; };                         // weight source of one packed K chunk pair

struct ConvProgram {
    int n_phases;
    int N, n_classes, cout, tmem_cols;
    int b_rows;                   // rows of a packed weight block: N, or 2N in hi/lo mode ([w_hi ; w_lo])
    int gD, gH, gW, gPx;          // tile grid (output grid for conv, input grid for transposed conv)
    long long zstride;            // voxels per padded z plane of the input (sub-)volumes
    int up_z, up_y, up_x;         // output coordinate = grid coordinate * up + class bit
    int relu;
    int w_smem_bytes;             // packed weights of all phases (resident in shared memory, loaded once per CTA)
    int slab_bytes, n_stages;     // operand slab of one load unit; number of slab stages (1 or 2)
    int tile_cols;                // TMEM columns of one accumulator stage
    int tmem_stages;              // accumulator stages: 2 (MMAs of tile k+1 overlap the epilogue of tile k inside one CTA), or 1 when two
                                  // stages would take more than 256 columns and the CTA is small enough for a second one per SM --
                                  // the neighbour CTA then provides the overlap and doubles the epilogue warps (transposed layers:
                                  // 8 output classes make the epilogue the long pole)
    int cls_z;                    // 1: the accumulator classes are consecutive output planes (z-sweep), output z = grid z * up_z + class
    int b_lbo_rows;               // rows between the two K chunks of a packed weight block (b_rows, or 3 * b_rows when stacked)
    int out_f32_pair;             // output (8 channels, hi/lo layout geometry) stored as raw fp32: channels 0-3 in the "hi" voxel, 4-7 in the "lo" voxel
                                  // (16 bytes each) -- for a consumer on CUDA cores (deconv_cout1_kernel), which then needs no hi + lo unpacking
    int debug;                    // EFFIMVS_TC_DEBUG bits (profiling only): 1 no MMA, 2 no operand copies, 4 no epilogue body, 8 no stores
    Phase ph[MAX_PHASES];
    Seg segs[MAX_SEGS];
    Op ops[MAX_OPS];
};
struct PackTable { int n_blocks, N, rows, cout, cin, transposed; int stack, brows; Block blk[MAX_BLOCKS]; };   // stack 3: rows = [a=2 ; a=1 ; a=0] blocks of brows rows, tap = a*9 + Block tap

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends in hardware for a bounded time; a barrier that never completes (a broken
    // transaction count) must surface as a launch failure, not as a hung GPU.
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor: core matrix = 8 rows x 16 bytes contiguous;
// LBO = byte distance between the two 8-element K chunks, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// two 8-column reads (the x_hi * w_hi and x_hi * w_lo halves of a hi/lo accumulator) in flight together, one wait
__device__ __forceinline__ void tmem_ld8x2(uint32_t taddr0, uint32_t taddr1, float (&v)[8], float (&w)[8]) {
    uint32_t r[8], q[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                 : "r"(taddr1));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = __uint_as_float(r[i]); w[i] = __uint_as_float(q[i]); }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack_bf16x8(const float (&v)[8]) {
    uint4 o;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    return o;
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& r, float (&v)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(p[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
    }
}
// store 8 channels of one voxel; hi/lo layouts get the bf16 value and the bf16 of the remainder
__device__ __forceinline__ void store_voxel(uint4* __restrict__ out, const ActLayout& L, int b, int g, int z, int y, int x,
                                            const float (&v)[8]) {
    const uint4 hi = pack_bf16x8(v);
    out[act_index(L, b, g, z, y, x)] = hi;
    if (L.lo_off) {
        float h[8], r[8];
        unpack_bf16x8(hi, h);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = v[j] - h[j];
        out[act_index(L, b, g + L.lo_off, z, y, x)] = pack_bf16x8(r);
    }
}
__device__ __forceinline__ void load_voxel(const uint4* __restrict__ in, const ActLayout& L, int b, int g, int z, int y, int x,
                                           float (&v)[8]) {
    unpack_bf16x8(__ldg(in + act_index(L, b, g, z, y, x)), v);
    if (L.lo_off) {
        float l[8];
        unpack_bf16x8(__ldg(in + act_index(L, b, g + L.lo_off, z, y, x)), l);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += l[j];
    }
}

// ------------------------------------------------------------------------------------------------
// the tile kernel: persistent, warp-specialised, software-pipelined over tiles
//   warp 4 (one lane)  TMA producer : bulk copies of the next load unit (tile x K phase) into slab stage
//   warp 5 (one lane)  MMA issuer   : tcgen05.mma sequence of a unit into TMEM accumulator stage k % 2
//   warps 0-3          epilogue     : TMEM -> registers -> bias / ReLU / residual -> bf16 -> global
// mbarriers: full[s] / empty[s] per slab stage (TMA <-> MMA), tmem_full[a] / tmem_empty[a] per
// accumulator stage (MMA <-> epilogue), wbar for the weights (loaded once per CTA).
// ------------------------------------------------------------------------------------------------
constexpr int MAX_STAGES = 2;
constexpr int CTA_THREADS = 192;

__global__ void __launch_bounds__(CTA_THREADS)
conv_tc_kernel(const __grid_constant__ ConvProgram P, const uint4* __restrict__ in, long long in_batch_stride,
               const uint8_t* __restrict__ wpk, const float* __restrict__ bias, const ActLayout OL, uint4* __restrict__ out,
               const ActLayout RL, const uint4* __restrict__ res, float* __restrict__ out_f32, int n_tiles, int tiles_per_plane) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2], bar_w;
    __shared__ uint32_t tmem_base_s;
    // warp index through a shuffle: the compiler then knows the role branches are warp-uniform (uniform-datapath operands for tcgen05.mma)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    uint8_t* wsm = smem;                              // all phases' weights, resident for the CTA's lifetime
    uint8_t* slab = smem + P.w_smem_bytes;            // n_stages slabs of slab_bytes
    const int S = P.n_stages;

    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)P.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 128) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
        mbar_init(&bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // programmatic dependent launch: TMEM allocation and barrier set-up above overlap the predecessor's tail; the packed
    // weights may have been written by the kernel right before this one (one-call entry points), so they wait as well
    pdl_trigger();
    pdl_wait();

    if (warp == 4) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            mbar_expect_tx(&bar_w, (uint32_t)P.w_smem_bytes);
            bulk_g2s(wsm, wpk, (uint32_t)P.w_smem_bytes, &bar_w);
            uint32_t u = 0;                                          // load unit counter
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int t = tile % tiles_per_plane, z = (tile / tiles_per_plane) % P.gD, b = tile / (tiles_per_plane * P.gD);
                const long long q0 = (long long)P.gPx + (long long)t * TILE_M;
                const uint4* base = in + (long long)b * in_batch_stride + (long long)z * P.zstride + q0;
                for (int p = 0; p < P.n_phases; ++p, ++u) {
                    const Phase ph = P.ph[p];
                    const uint32_t s = u % S, n = u / S;
                    mbar_wait(&bar_empty[s], (n & 1) ^ 1);          // slab stage free (first use passes)
                    uint32_t bytes = 0;
                    if (!(P.debug & 2))
                        for (int i = ph.seg_begin; i < ph.seg_end; ++i) bytes += (uint32_t)P.segs[i].copy_vox * 16u;
                    mbar_expect_tx(&bar_full[s], bytes);
                    uint8_t* dst = slab + (size_t)s * P.slab_bytes;
                    if (!(P.debug & 2))
                        for (int i = ph.seg_begin; i < ph.seg_end; ++i)
                            bulk_g2s(dst + (size_t)P.segs[i].slot * SEG_BYTES, base + P.segs[i].src_off, (uint32_t)P.segs[i].copy_vox * 16u, &bar_full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t w0 = smem_u32(wsm), b_lbo = (uint32_t)P.b_lbo_rows * 16u;
            mbar_wait(&bar_w, 0);
            uint32_t u = 0, k = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
                const uint32_t acc = P.tmem_stages == 2 ? (k & 1) : 0u, round = P.tmem_stages == 2 ? (k >> 1) : k;
                mbar_wait(&bar_tempty[acc], (round & 1) ^ 1);       // epilogue drained this accumulator stage
                tc_fence_after();
                const uint32_t d0 = tmem + acc * (uint32_t)P.tile_cols;
                for (int p = 0; p < P.n_phases; ++p, ++u) {
                    const Phase ph = P.ph[p];
                    const uint32_t s = u % S, n = u / S;
                    mbar_wait(&bar_full[s], n & 1);                  // operands landed
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(slab + (size_t)s * P.slab_bytes), b0 = w0 + (uint32_t)ph.w_off;
                    for (int i = ph.op_begin; i < ph.op_end && !(P.debug & 1); ++i) {
                        const Op op = P.ops[i];
                        umma_bf16(d0 + op.d_col, umma_desc(a0 + op.a_off, op.a_lbo, 128), umma_desc(b0 + op.b_off, b_lbo, 128),
                                  umma_idesc(op.n8 * 8), op.accum);
                    }
                    umma_commit(&bar_empty[s]);                      // slab stage reusable once these MMAs retire
                }
                umma_commit(&bar_tfull[acc]);                        // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: thread <-> tile row <-> TMEM lane ----------------
        const int groups = (P.cout + 7) / 8;
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int t = tile % tiles_per_plane, z = (tile / tiles_per_plane) % P.gD, b = tile / (tiles_per_plane * P.gD);
            const int q = P.gPx + t * TILE_M + tid;
            const int yp = q / P.gPx, xp = q - yp * P.gPx;
            const bool interior = yp >= 1 && yp <= P.gH && xp >= 1 && xp <= P.gW;
            const int gy = yp - 1, gx = xp - 1;
            const uint32_t acc = P.tmem_stages == 2 ? (k & 1) : 0u, round = P.tmem_stages == 2 ? (k >> 1) : k;
            mbar_wait(&bar_tfull[acc], round & 1);
            tc_fence_after();
            const uint32_t lane_base = tmem + acc * (uint32_t)P.tile_cols + ((uint32_t)(warp * 32) << 16);
            for (int g = 0; g < groups; ++g) {
                // residual voxels of all output classes of this channel group: issued up front so that their
                // global-memory latency overlaps the TMEM reads instead of serialising class by class
                uint4 rhi[8], rlo[8];
                if (res && interior && !(P.debug & 4)) {
#pragma unroll
                    for (int cls = 0; cls < 8; ++cls) {
                        if (cls < P.n_classes) {
                            int pz = 0, py = 0, px = 0;
                            if (P.cls_z) pz = cls;
                            else if (P.n_classes == 8) { pz = cls >> 2; py = (cls >> 1) & 1; px = cls & 1; }
                            else if (P.n_classes == 4) { py = cls >> 1; px = cls & 1; }
                            const int oz = z * P.up_z + pz, oy = gy * P.up_y + py, ox = gx * P.up_x + px;
                            rhi[cls] = __ldg(res + act_index(RL, b, g, oz, oy, ox));
                            if (RL.lo_off) rlo[cls] = __ldg(res + act_index(RL, b, g + RL.lo_off, oz, oy, ox));
                        }
                    }
                }
#pragma unroll
                for (int cls = 0; cls < 8; ++cls) {
                    if (cls >= P.n_classes) break;
                    int pz = 0, py = 0, px = 0;
                    if (P.cls_z) pz = cls;
                    else if (P.n_classes == 8) { pz = cls >> 2; py = (cls >> 1) & 1; px = cls & 1; }
                    else if (P.n_classes == 4) { py = cls >> 1; px = cls & 1; }
                    const int oz = z * P.up_z + pz, oy = gy * P.up_y + py, ox = gx * P.up_x + px;
                    float v[8];
                    if (P.b_rows != P.N) {   // hi/lo mode: columns [N, 2N) hold x_hi * w_lo
                        float w2[8];
                        tmem_ld8x2(lane_base + (uint32_t)(cls * P.b_rows + g * 8), lane_base + (uint32_t)(cls * P.b_rows + P.N + g * 8), v, w2);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += w2[j];
                    } else {
                        tmem_ld8(lane_base + (uint32_t)(cls * P.b_rows + g * 8), v);
                    }
                    if (cls == P.n_classes - 1 && g == groups - 1) {
                        // last TMEM read of this tile: hand the accumulator stage back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_tempty[acc])) : "memory");
                    }
                    if (!interior || (P.debug & 4)) continue;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (bias && g * 8 + j < P.cout) v[j] += __ldg(bias + g * 8 + j);
                        if (P.relu) v[j] = fmaxf(v[j], 0.0f);
                    }
                    if (out_f32) {  // single-channel fp32 NCDHW output
                        if (!(P.debug & 8)) out_f32[(((size_t)b * OL.D + oz) * OL.H + oy) * OL.W + ox] = v[0];
                        continue;
                    }
                    if (res) {
                        float r[8];
                        unpack_bf16x8(rhi[cls], r);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += r[j];
                        if (RL.lo_off) {
                            unpack_bf16x8(rlo[cls], r);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] += r[j];
                        }
                    }
                    if (P.out_f32_pair) {
                        out[act_index(OL, b, g, oz, oy, ox)] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                        out[act_index(OL, b, g + OL.lo_off, oz, oy, ox)] = make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
                        continue;
                    }
                    if (!(P.debug & 8)) store_voxel(out, OL, b, g, oz, oy, ox, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)P.tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------
// helper kernels: weight packing, single-input-channel convolution (CUDA cores), layout conversion
// ------------------------------------------------------------------------------------------------
// packed block: [chunk j][row][8 k] bf16; rows [0,N) = output channels (zero beyond cout) of bf16(w);
// in hi/lo mode rows [N,2N) hold bf16(w - bf16(w)).
__global__ void pack_weights_kernel(const __grid_constant__ PackTable T, const float* __restrict__ w, __nv_bfloat16* __restrict__ dst) {
    const int total = T.n_blocks * 2 * T.rows * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kk = i & 7, row_all = (i >> 3) % T.rows, j = (i / (8 * T.rows)) & 1, blk = i / (16 * T.rows);
        const int row = T.stack > 1 ? row_all % T.brows : row_all;
        const bool lo = row >= T.N;
        const int n = lo ? row - T.N : row;
        const Block c = T.blk[blk];
        const int tap_bc = j ? c.tap1 : c.tap0, cb = j ? c.cb1 : c.cb0;
        const int tap = (T.stack > 1 && tap_bc >= 0) ? (T.stack - 1 - row_all / T.brows) * 9 + tap_bc : tap_bc;
        float v = 0.0f;
        if (tap >= 0 && n < T.cout) {
            const int ci = cb + kk;
            v = T.transposed ? w[((size_t)ci * T.cout + n) * 27 + tap] : w[((size_t)n * T.cin + ci) * 27 + tap];
        }
        __nv_bfloat16 hi = __float2bfloat16(v);
        dst[i] = lo ? __float2bfloat16(v - __bfloat162float(hi)) : hi;
    }
}

struct PackJobs { int n; PackTable T[8]; const float* w[8]; __nv_bfloat16* dst[8]; };
__global__ void pack_weights_multi_kernel(const __grid_constant__ PackJobs J) {
    const PackTable& T = J.T[blockIdx.y];
    const float* __restrict__ w = J.w[blockIdx.y];
    __nv_bfloat16* __restrict__ dst = J.dst[blockIdx.y];
    const int total = T.n_blocks * 2 * T.rows * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kk = i & 7, row_all = (i >> 3) % T.rows, j = (i / (8 * T.rows)) & 1, blk = i / (16 * T.rows);
        const int row = T.stack > 1 ? row_all % T.brows : row_all;
        const bool lo = row >= T.N;
        const int n = lo ? row - T.N : row;
        const Block c = T.blk[blk];
        const int tap_bc = j ? c.tap1 : c.tap0, cb = j ? c.cb1 : c.cb0;
        const int tap = (T.stack > 1 && tap_bc >= 0) ? (T.stack - 1 - row_all / T.brows) * 9 + tap_bc : tap_bc;
        float v = 0.0f;
        if (tap >= 0 && n < T.cout) {
            const int ci = cb + kk;
            v = T.transposed ? w[((size_t)ci * T.cout + n) * 27 + tap] : w[((size_t)n * T.cin + ci) * 27 + tap];
        }
        __nv_bfloat16 hi = __float2bfloat16(v);
        dst[i] = lo ? __float2bfloat16(v - __bfloat162float(hi)) : hi;
    }
}

// Zeroes guards and halos of up to 8 activation buffers (the interiors are fully overwritten by the
// layers; tiles read halos as the zero padding of the convolution).  One block per padded z plane of
// a (sub-)volume: halo planes and the two guards are cleared whole, interior planes only get their
// first / last row and first / last column cleared.
struct HaloJobs { int n; uint4* ptr[8]; ActLayout L[8]; int planes_before[9]; };
__global__ void zero_halo_kernel(const __grid_constant__ HaloJobs J, int B) {
    int job = 0;
    while (job + 1 < J.n && (int)blockIdx.x >= J.planes_before[job + 1]) ++job;
    const ActLayout& L = J.L[job];
    const int d = L.kind == L_SPLIT ? L.D / 2 : L.D, Pz = d + 2;
    const int local = blockIdx.x - J.planes_before[job];       // (volume index, zp)
    const int vol = local / Pz, zp = local - vol * Pz;
    uint4* __restrict__ base = J.ptr[job] + (long long)vol * L.vs;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    // blockIdx.y splits a plane's work: the two halo planes of a volume are cleared whole (a fifth of an 8-plane
    // volume), which one 128-thread block per plane did at a fraction of the HBM rate
    const int tid = blockIdx.y * blockDim.x + threadIdx.x, nt = gridDim.y * blockDim.x;
    if (zp == 0) for (int i = tid; i < L.guard; i += nt) base[i] = zero;                                     // front guard
    if (zp == Pz - 1) for (int i = tid; i < L.guard; i += nt) base[L.vs - L.guard + i] = zero;               // back guard
    uint4* __restrict__ plane = base + L.guard + (long long)zp * L.zstride;
    if (zp == 0 || zp == Pz - 1) {
        for (int i = tid; i < (int)L.zstride; i += nt) plane[i] = zero;
    } else {
        for (int i = tid; i < L.Px; i += nt) { plane[i] = zero; plane[(long long)(L.Py - 1) * L.Px + i] = zero; }
        for (int y = 1 + tid; y < L.Py - 1; y += nt) { plane[(long long)y * L.Px] = zero; plane[(long long)y * L.Px + L.Px - 1] = zero; }
    }
}

// 1 -> 8 channel 3x3x3 convolution (BN folded, + bias + ReLU), fp32 NCDHW in, one c8 plane out.
// stride (1, s, s) with s in {1, 2}.  CUDA cores: K = 27 is too thin for an MMA tile.  A thread
// computes CIN1_X consecutive outputs of a row so that the input row segments and the broadcast
// weight reads are shared between them.
constexpr int CIN1_X = 4;
template <int S>
__global__ void __launch_bounds__(128)
conv_cin1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int D, int H, int W,
                 const ActLayout OL, int plane, uint4* __restrict__ out) {
    __shared__ __align__(16) float sw[27 * 8 + 8];
    pdl_trigger();
    for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) sw[(i % 27) * 8 + i / 27] = w[i];  // [tap][co]  (weights: constants, before the wait)
    if (threadIdx.x < 8) sw[216 + threadIdx.x] = bias[threadIdx.x];
    pdl_wait();
    __syncthreads();
    const int b = blockIdx.y;
    const int Ho = OL.H, Wo = OL.W;
    const int Wq = (Wo + CIN1_X - 1) / CIN1_X;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (size_t)D * Ho * Wq) return;
    const int ox0 = (int)(o % Wq) * CIN1_X, oy = (int)((o / Wq) % Ho), oz = (int)(o / ((size_t)Wq * Ho));
    constexpr int NIN = (CIN1_X - 1) * S + 3;   // input columns feeding CIN1_X outputs
    float acc[CIN1_X][8];
#pragma unroll
    for (int i = 0; i < CIN1_X; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = sw[216 + c];
    const float* xb = x + (size_t)b * D * H * W;
    const int ix0 = ox0 * S - 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int iz = oz + a - 1;
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const int iy = oy * S + bb - 1;
            const bool row_ok = iz >= 0 && iz < D && iy >= 0 && iy < H;
            const float* row = xb + ((size_t)(row_ok ? iz : 0) * H + (row_ok ? iy : 0)) * W;
            float in[NIN];
#pragma unroll
            for (int j = 0; j < NIN; ++j) {
                const int ix = ix0 + j;
                in[j] = (row_ok && ix >= 0 && ix < W) ? __ldg(row + ix) : 0.0f;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 w0 = *reinterpret_cast<const float4*>(sw + (a * 9 + bb * 3 + c) * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sw + (a * 9 + bb * 3 + c) * 8 + 4);
#pragma unroll
                for (int i = 0; i < CIN1_X; ++i) {
                    const float v = in[i * S + c];
                    acc[i][0] = fmaf(v, w0.x, acc[i][0]); acc[i][1] = fmaf(v, w0.y, acc[i][1]);
                    acc[i][2] = fmaf(v, w0.z, acc[i][2]); acc[i][3] = fmaf(v, w0.w, acc[i][3]);
                    acc[i][4] = fmaf(v, w1.x, acc[i][4]); acc[i][5] = fmaf(v, w1.y, acc[i][5]);
                    acc[i][6] = fmaf(v, w1.z, acc[i][6]); acc[i][7] = fmaf(v, w1.w, acc[i][7]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CIN1_X; ++i) {
        if (ox0 + i >= Wo) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaxf(acc[i][j], 0.0f);
        store_voxel(out, OL, b, plane, oz, oy, ox0 + i, acc[i]);
    }
}

// fp32 NCDHW <-> c8 layouts (single-layer entry point and tests)
__global__ void to_c8_kernel(const float* __restrict__ x, int C, const ActLayout L, uint4* __restrict__ out) {
    const int b = blockIdx.y;
    const int groups = (C + 7) / 8;
    const size_t vox = (size_t)L.D * L.H * L.W;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= vox * groups) return;
    const int g = (int)(o / vox);
    const size_t r = o % vox;
    const int xx = (int)(r % L.W), yy = (int)((r / L.W) % L.H), zz = (int)(r / ((size_t)L.W * L.H));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        v[j] = c < C ? x[((size_t)b * C + c) * vox + r] : 0.0f;
    }
    store_voxel(out, L, b, g, zz, yy, xx, v);
}
__global__ void from_c8_kernel(const uint4* __restrict__ in, int C, const ActLayout L, float* __restrict__ y) {
    const int b = blockIdx.y;
    const int groups = (C + 7) / 8;
    const size_t vox = (size_t)L.D * L.H * L.W;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= vox * groups) return;
    const int g = (int)(o / vox);
    const size_t r = o % vox;
    const int xx = (int)(r % L.W), yy = (int)((r / L.W) % L.H), zz = (int)(r / ((size_t)L.W * L.H));
    float v[8];
    load_voxel(in, L, b, g, zz, yy, xx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        if (c < C) y[((size_t)b * C + c) * vox + r] = v[j];
    }
}

// ------------------------------------------------------------------------------------------------
// host: program builders
// ------------------------------------------------------------------------------------------------
struct Term { int seg, byte_off, tap, cls; };      // seg: index within one plane's segments
struct PlaneSeg { long long off; int copy_vox; };  // per-plane segment: source offset relative to the plane base

int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

// Assembles segments, MMA ops and packed-weight blocks.
//   NP channel planes (Cin / 8), segments per plane `pseg`, `plane_stride` voxels between planes,
//   hilo: operands carried as hi + lo bf16 (three MMA products per K chunk pair).
// Shared-memory slot of (half h, local plane pl, plane segment k) inside a phase: (h * NPph + pl) * SPP + k.
bool assemble(ConvProgram& P, PackTable& T, const std::vector<Term>& terms, const std::vector<PlaneSeg>& pseg, int NP,
              long long plane_stride, long long lo_plane_off, bool hilo) {
    const int SPP = (int)pseg.size();
    const int halves = hilo ? 2 : 1;
    // K phases: hi/lo with 4 planes does not fit one slab -> one phase per plane pair
    const int n_phases = (hilo && NP == 4) ? 2 : 1;
    const int NPph = NP / n_phases;
    if (n_phases > MAX_PHASES) return false;
    P.n_phases = n_phases;
    P.b_rows = T.rows = hilo ? 2 * P.N : P.N;
    P.b_lbo_rows = P.b_rows; P.cls_z = 0; T.stack = 1; T.brows = T.rows;
    int n_seg = 0, n_op = 0, n_blk = 0;
    std::vector<bool> started(P.n_classes, false);
    size_t max_slab = 0;
    struct Base { uint32_t a_off, lbo; int cls; Block blk; };
    for (int ph = 0; ph < n_phases; ++ph) {
        Phase& F = P.ph[ph];
        F.seg_begin = n_seg; F.op_begin = n_op;
        const int blk_begin = n_blk;
        for (int h = 0; h < halves; ++h)
            for (int pl = 0; pl < NPph; ++pl)
                for (int k = 0; k < SPP; ++k) {
                    if (n_seg >= MAX_SEGS) return false;
                    Seg& s = P.segs[n_seg++];
                    s.src_off = (long long)(ph * NPph + pl) * plane_stride + (h ? lo_plane_off : 0) + pseg[k].off;
                    s.copy_vox = pseg[k].copy_vox;
                    s.slot = (h * NPph + pl) * SPP + k;
                }
        const uint32_t lo_smem = (uint32_t)(NPph * SPP * SEG_BYTES);  // hi -> lo distance in the slab
        const uint32_t blk_bytes = 2u * P.b_rows * 16u;
        std::vector<Base> bases;
        if (NP == 1) {
            // 8 input channels: pair taps of the same class in ascending shared-memory address (LBO > 0)
            for (int cls = 0; cls < P.n_classes; ++cls) {
                std::vector<Term> v;
                for (auto& t : terms) if (t.cls == cls) v.push_back(t);
                std::sort(v.begin(), v.end(), [](const Term& x, const Term& y) { return x.seg * SEG_BYTES + x.byte_off < y.seg * SEG_BYTES + y.byte_off; });
                for (size_t i = 0; i < v.size(); i += 2) {
                    const uint32_t a0 = (uint32_t)(v[i].seg * SEG_BYTES + v[i].byte_off);
                    if (i + 1 < v.size()) {
                        const uint32_t a1 = (uint32_t)(v[i + 1].seg * SEG_BYTES + v[i + 1].byte_off);
                        bases.push_back(Base{a0, a1 - a0, cls, Block{(short)v[i].tap, 0, (short)v[i + 1].tap, 0}});
                    } else {
                        bases.push_back(Base{a0, 16, cls, Block{(short)v[i].tap, 0, -1, 0}});  // zero-weight dummy chunk
                    }
                }
            }
        } else {
            for (auto& t : terms)
                for (int pl = 0; pl < NPph; pl += 2) {
                    const int cb = (ph * NPph + pl) * 8;
                    bases.push_back(Base{(uint32_t)((pl * SPP + t.seg) * SEG_BYTES + t.byte_off), (uint32_t)(SPP * SEG_BYTES), t.cls,
                                         Block{(short)t.tap, (short)cb, (short)t.tap, (short)(cb + 8)}});
                }
        }
        for (auto& bs : bases) {
            if (n_blk + 1 > MAX_BLOCKS || n_op + halves > MAX_OPS) return false;
            const uint32_t b_off = (uint32_t)(n_blk - blk_begin) * blk_bytes;
            T.blk[n_blk++] = bs.blk;
            auto push = [&](uint32_t a_off, int n) {
                Op& op = P.ops[n_op++];
                op.a_off = a_off; op.a_lbo = bs.lbo; op.b_off = b_off;
                op.d_col = (uint16_t)(bs.cls * P.b_rows); op.accum = started[bs.cls] ? 1 : 0; op.n8 = (uint8_t)(n / 8);
                started[bs.cls] = true;
            };
            push(bs.a_off, P.b_rows);                      // x_hi * [w_hi ; w_lo]  (or x * w)
            if (hilo) push(bs.a_off + lo_smem, P.N);        // x_lo * w_hi, into the first N columns
        }
        F.seg_end = n_seg; F.op_end = n_op;
        F.w_off = (int)((size_t)blk_begin * blk_bytes); F.w_bytes = (int)((size_t)(n_blk - blk_begin) * blk_bytes);
        max_slab = std::max(max_slab, (size_t)halves * NPph * SPP * SEG_BYTES);
    }
    T.n_blocks = n_blk;
    P.w_smem_bytes = (int)((size_t)n_blk * 2u * P.b_rows * 16u);
    P.slab_bytes = (int)max_slab;
    const size_t budget = 220 * 1024;
    if ((size_t)P.w_smem_bytes + P.slab_bytes > budget) return false;
    P.n_stages = ((size_t)P.w_smem_bytes + 2 * (size_t)P.slab_bytes <= budget) ? 2 : 1;
    // An MMA of this shape costs a CTA ~80 clk of mostly operand-fetch latency that a second resident CTA hides
    // (measured: 16->8 layer 150 -> 113 us), which is worth more than double-buffering the slab within one CTA:
    // when two stages would leave room for only one CTA per SM but a single stage fits twice, take the latter.
    const size_t two_cta = 110 * 1024;
    if (P.n_stages == 2 && (size_t)P.w_smem_bytes + 2 * (size_t)P.slab_bytes > two_cta &&
        (size_t)P.w_smem_bytes + (size_t)P.slab_bytes <= two_cta)
        P.n_stages = 1;
    P.tile_cols = P.b_rows * P.n_classes;
    P.tmem_stages = 2;
    P.tmem_cols = pow2_cols(2 * P.tile_cols);
    if (P.tmem_cols > 256 && pow2_cols(P.tile_cols) <= 256 && (size_t)P.w_smem_bytes + (size_t)P.slab_bytes <= two_cta &&
        !getenv("EFFIMVS_TC_TWO_TMEM_STAGES")) {
        P.tmem_stages = 1;
        P.n_stages = 1;
        P.tmem_cols = pow2_cols(P.tile_cols);
    }
    if (P.tmem_cols > 512) return false;
    return true;
}

size_t program_smem(const ConvProgram& P) { return (size_t)P.w_smem_bytes + (size_t)P.n_stages * P.slab_bytes; }
size_t program_weight_bytes(const PackTable& T) { return (size_t)T.n_blocks * 2 * T.rows * 16; }

void init_program(ConvProgram& P, PackTable& T, int Cin, int Cout, int n_classes, int gD, int gH, int gW, const ActLayout& IL,
                  int relu, int transposed) {
    memset(&P, 0, sizeof(P));
    memset(&T, 0, sizeof(T));
    P.N = std::max(16, (Cout + 15) / 16 * 16);
    P.n_classes = n_classes; P.cout = Cout; P.tmem_cols = pow2_cols(P.N * n_classes);
    P.gD = gD; P.gH = gH; P.gW = gW; P.gPx = IL.Px; P.zstride = IL.zstride;
    P.up_z = P.up_y = P.up_x = 1; P.relu = relu;
    T.N = P.N; T.cout = Cout; T.cin = Cin; T.transposed = transposed;
}

// stride-1 convolution, input REGULAR on the output grid
bool build_conv_s1_zsweep(ConvProgram& P, PackTable& T, int Cin, int Cout, const ActLayout& IL, int relu, bool hilo);

bool build_conv_s1(ConvProgram& P, PackTable& T, int Cin, int Cout, const ActLayout& IL, int relu) {
    if (build_conv_s1_zsweep(P, T, Cin, Cout, IL, relu, IL.lo_off != 0)) return true;
    init_program(P, T, Cin, Cout, 1, IL.D, IL.H, IL.W, IL, relu, 0);
    std::vector<PlaneSeg> pseg;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            pseg.push_back(PlaneSeg{IL.guard + (long long)a * IL.zstride + (long long)(b - 1) * IL.Px - 1, TILE_M + 3});
    std::vector<Term> terms;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c) terms.push_back(Term{a * 3 + b, c * 16, a * 9 + b * 3 + c, 0});
    return assemble(P, T, terms, pseg, Cin / 8, IL.vs, (long long)IL.lo_off * IL.vs, IL.lo_off != 0);
}

// Stride-1 convolution as a z-sweep: a unit is one 128-position tile in ZSWEEP_R consecutive output planes, whose
// accumulators sit side by side in TMEM.  A load phase stages ONE input plane (its three y rows per channel plane
// and hi/lo half); every staged row then feeds the up-to-three output planes it contributes to with a single MMA
// against the stacked weight block [w(a=2) ; w(a=1) ; w(a=0)] -- N = 3 * b_rows instead of three N = b_rows
// MMAs, i.e. one A-operand fetch (the binding cost of these layers, ~32 clk for 8-16 clk of math) serves three
// taps, and each input row is staged (R+2)/R times per output plane instead of three times.
// The x_lo operand is multiplied with the full [w_hi ; w_lo] block too (the extra x_lo * w_lo term is 2^-16 of
// 2^-16 and only makes the product more exact), so hi and lo MMAs share shape, weights and accumulator columns.
bool build_conv_s1_zsweep(ConvProgram& P, PackTable& T, int Cin, int Cout, const ActLayout& IL, int relu, bool hilo) {
    const int R = ZSWEEP_R, NP = Cin / 8, halves = hilo ? 2 : 1;
    if (IL.D % R != 0 || (NP != 1 && NP != 2) || getenv("EFFIMVS_TC_NO_ZSWEEP")) return false;
    init_program(P, T, Cin, Cout, R, IL.D / R, IL.H, IL.W, IL, relu, 0);
    P.up_z = R; P.cls_z = 1;
    P.zstride = (long long)R * IL.zstride;          // the tile enumeration advances by R planes
    P.b_rows = hilo ? 2 * P.N : P.N;
    P.b_lbo_rows = 3 * P.b_rows;
    T.rows = 3 * P.b_rows; T.stack = 3; T.brows = P.b_rows;
    const long long plane_stride = IL.vs, lo_plane_off = (long long)IL.lo_off * IL.vs;
    const uint32_t blk_bytes = 2u * (uint32_t)T.rows * 16u;
    const uint32_t lo_smem = (uint32_t)(NP * 3 * SEG_BYTES);
    // weight blocks: one per K-chunk pair of (b, c) taps -- shared by all phases
    struct Base { uint32_t a_off, lbo; int blk; };
    std::vector<Base> bases;
    int n_blk = 0;
    if (NP == 1) {   // 8 input channels: pair the nine (b, c) taps in ascending shared-memory address
        for (int t = 0; t < 9; t += 2) {
            const int b0 = t / 3, c0 = t % 3;
            const uint32_t a0 = (uint32_t)(b0 * SEG_BYTES + c0 * 16);
            if (t + 1 < 9) {
                const int b1 = (t + 1) / 3, c1 = (t + 1) % 3;
                const uint32_t a1 = (uint32_t)(b1 * SEG_BYTES + c1 * 16);
                T.blk[n_blk] = Block{(short)t, 0, (short)(t + 1), 0};
                bases.push_back(Base{a0, a1 - a0, n_blk++});
            } else {
                T.blk[n_blk] = Block{(short)t, 0, -1, 0};                 // zero-weight dummy chunk
                bases.push_back(Base{a0, 16, n_blk++});
            }
        }
    } else {         // 16 input channels: the two channel planes are the two K chunks
        for (int t = 0; t < 9; ++t) {
            const int b = t / 3, c = t % 3;
            T.blk[n_blk] = Block{(short)t, 0, (short)t, 8};
            bases.push_back(Base{(uint32_t)(b * SEG_BYTES + c * 16), (uint32_t)(3 * SEG_BYTES), n_blk++});
        }
    }
    T.n_blocks = n_blk;
    P.n_phases = R + 2;
    if (P.n_phases > MAX_PHASES) return false;
    int n_seg = 0, n_op = 0;
    std::vector<bool> started(R, false);
    for (int zi = 0; zi < R + 2; ++zi) {             // padded input plane z0 * R + zi
        Phase& F = P.ph[zi];
        F.seg_begin = n_seg; F.op_begin = n_op; F.w_off = 0; F.w_bytes = 0;
        for (int h = 0; h < halves; ++h)
            for (int pl = 0; pl < NP; ++pl)
                for (int b = 0; b < 3; ++b) {
                    if (n_seg >= MAX_SEGS) return false;
                    Seg& sg = P.segs[n_seg++];
                    sg.src_off = (long long)pl * plane_stride + (h ? lo_plane_off : 0) + IL.guard + (long long)zi * IL.zstride +
                                 (long long)(b - 1) * IL.Px - 1;
                    sg.copy_vox = TILE_M + 3;
                    sg.slot = (h * NP + pl) * 3 + b;
                }
        // output plane o (0..R-1) reads padded input planes o + a: this plane serves a in [a_min, a_max], o = zi - a
        const int a_max = std::min(2, zi), a_min = std::max(0, zi - R + 1);
        const int o_first = zi - a_max, o_last = zi - a_min;
        for (auto& bs : bases)
            for (int h = 0; h < halves; ++h) {
                // split into runs of output planes that are all started / all fresh (one accumulate flag per MMA)
                int o = o_first;
                while (o <= o_last) {
                    int e = o;
                    while (e + 1 <= o_last && started[e + 1] == started[o]) ++e;
                    if (n_op >= MAX_OPS) return false;
                    Op& op = P.ops[n_op++];
                    op.a_off = bs.a_off + (h ? lo_smem : 0u);
                    op.a_lbo = bs.lbo;
                    const int a_hi = zi - o;                                  // tap of the first covered output plane
                    op.b_off = (uint32_t)bs.blk * blk_bytes + (uint32_t)(2 - a_hi) * (uint32_t)P.b_rows * 16u;
                    op.d_col = (uint16_t)(o * P.b_rows);
                    op.n8 = (uint8_t)((e - o + 1) * P.b_rows / 8);
                    op.accum = started[o] ? 1 : 0;
                    for (int k = o; k <= e; ++k) started[k] = true;
                    o = e + 1;
                }
            }
        F.seg_end = n_seg; F.op_end = n_op;
    }
    P.w_smem_bytes = (int)((size_t)n_blk * blk_bytes);
    P.slab_bytes = halves * NP * 3 * SEG_BYTES;
    const size_t budget = 220 * 1024;
    if ((size_t)P.w_smem_bytes + P.slab_bytes > budget) return false;
    P.n_stages = ((size_t)P.w_smem_bytes + 2 * (size_t)P.slab_bytes <= budget) ? 2 : 1;
    P.tile_cols = R * P.b_rows;
    P.tmem_stages = 2;
    P.tmem_cols = pow2_cols(2 * P.tile_cols);
    return P.tmem_cols <= 512 && n_blk <= MAX_BLOCKS;
}

// stride-(2,2,2) convolution, input PARITY-SPLIT (its sub-volumes live on the output grid)
bool build_conv_s2(ConvProgram& P, PackTable& T, int Cin, int Cout, const ActLayout& IL, int relu) {
    init_program(P, T, Cin, Cout, 1, IL.D / 2, IL.H / 2, IL.W / 2, IL, relu, 0);
    std::vector<PlaneSeg> pseg;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int ex = 0; ex < 2; ++ex) {
                // input coordinate 2o + k - 1: k = 1 -> even sub-volume, index o; k = 0 / 2 -> odd sub-volume, index o - 1 / o
                const int sub = ((a != 1) << 2) | ((b != 1) << 1) | ex;
                pseg.push_back(PlaneSeg{(long long)sub * IL.vs + IL.guard + (long long)(a == 0 ? 0 : 1) * IL.zstride +
                                            (long long)(b == 0 ? -1 : 0) * IL.Px - ex, TILE_M + 2});
            }
    std::vector<Term> terms;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c) terms.push_back(Term{(a * 3 + b) * 2 + (c != 1), c == 2 ? 16 : 0, a * 9 + b * 3 + c, 0});
    return assemble(P, T, terms, pseg, Cin / 8, 8 * IL.vs, (long long)IL.lo_off * 8 * IL.vs, IL.lo_off != 0);
}

// transposed convolution, stride (sz,2,2) with sz in {1,2}; input REGULAR; tile grid = input grid
bool build_deconv(ConvProgram& P, PackTable& T, int Cin, int Cout, const ActLayout& IL, int sz, int relu) {
    init_program(P, T, Cin, Cout, sz == 2 ? 8 : 4, IL.D, IL.H, IL.W, IL, relu, 1);
    P.up_z = sz; P.up_y = 2; P.up_x = 2;
    const int NZ = sz == 2 ? 2 : 3;
    // z slot k holds padded input plane z + (sz == 2 ? 1 + k : k), i.e. input z + k (stride 2) or z - 1 + k (stride 1)
    std::vector<PlaneSeg> pseg;
    for (int k = 0; k < NZ; ++k)
        for (int dy = 0; dy < 2; ++dy)
            pseg.push_back(PlaneSeg{IL.guard + (long long)(sz == 2 ? 1 + k : k) * IL.zstride + (long long)dy * IL.Px, TILE_M + 2});
    std::vector<Term> terms;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c) {
                // output o = s*i + k - 1.  stride 2: class bit = (k != 1), input offset +1 when k == 0.  stride 1: i = o + 1 - k.
                const int kz = sz == 2 ? (a == 0) : (2 - a);
                const int cls = sz == 2 ? (((a != 1) << 2) | ((b != 1) << 1) | (c != 1)) : (((b != 1) << 1) | (c != 1));
                terms.push_back(Term{kz * 2 + (b == 0), (c == 0) * 16, a * 9 + b * 3 + c, cls});
            }
    return assemble(P, T, terms, pseg, Cin / 8, IL.vs, (long long)IL.lo_off * IL.vs, IL.lo_off != 0);
}

// ------------------------------------------------------------------------------------------------
// host: launches
// ------------------------------------------------------------------------------------------------
int run_pack(const PackTable& T, const float* w, void* dst, cudaStream_t st) {
    int total = T.n_blocks * 2 * T.rows * 8;
    pack_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(T, w, (__nv_bfloat16*)dst);
    return check_launch("pack_weights_kernel");
}

int run_pack_multi(int n, const PackTable* T, const float* const* w, void* const* dst, cudaStream_t st) {
    static thread_local PackJobs J;
    J.n = n;
    int most = 0;
    for (int i = 0; i < n; ++i) {
        J.T[i] = T[i]; J.w[i] = w[i]; J.dst[i] = (__nv_bfloat16*)dst[i];
        most = std::max(most, T[i].n_blocks * 2 * T[i].rows * 8);
    }
    pack_weights_multi_kernel<<<dim3(ceil_div(most, 256), n), 256, 0, st>>>(J);
    return check_launch("pack_weights_multi_kernel");
}

int run_zero_halo(int n, void* const* ptr, const ActLayout* L, int B, cudaStream_t st) {
    static thread_local HaloJobs J;
    J.n = n;
    int total = 0;
    for (int i = 0; i < n; ++i) {
        J.ptr[i] = (uint4*)ptr[i]; J.L[i] = L[i];
        J.planes_before[i] = total;
        const int d = L[i].kind == L_SPLIT ? L[i].D / 2 : L[i].D;
        total += B * L[i].planes * (L[i].kind == L_SPLIT ? 8 : 1) * (d + 2);
    }
    J.planes_before[n] = total;
    zero_halo_kernel<<<dim3(total, 16), 128, 0, st>>>(J, B);
    return check_launch("zero_halo_kernel");
}

int run_tile_kernel(const ConvProgram& P, int B, const void* in, const ActLayout& IL, const void* wpk, const float* bias,
                    const ActLayout& OL, void* out, const ActLayout* RL, const void* res, float* out_f32, cudaStream_t st) {
    size_t smem = program_smem(P);
    EFFI_REQUIRE(smem <= 220 * 1024, EFFIMVS_EUNSUPPORTED, "conv_tc: tile needs %zu bytes of shared memory", smem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return EFFIMVS_ECUDA; }
        attr_set = true;
    }
    const int tiles_per_plane = ceil_div(P.gH * P.gPx, TILE_M);
    const long long n_tiles = (long long)tiles_per_plane * P.gD * B;
    EFFI_REQUIRE(n_tiles < (1ll << 31) && (long long)P.gPx * (P.gH + 2) < (1ll << 30), EFFIMVS_EUNSUPPORTED, "conv_tc: volume too large");
    // persistent CTAs: as many per SM as shared memory and TMEM (512 columns) allow, at most 2
    int per_sm = (smem <= 110 * 1024 && P.tmem_cols <= 256) ? 2 : 1;
    if (const char* e = getenv("EFFIMVS_TC_MAXCTA")) {
        const int by_smem = (int)((227 * 1024) / (smem + 1024 + 4096)), by_tmem = 512 / P.tmem_cols;
        per_sm = std::max(1, std::min(atoi(e), std::min(by_smem, by_tmem)));
    }
    const int grid = (int)std::min<long long>(n_tiles, (long long)kNumSMs * per_sm);
    ActLayout rl = RL ? *RL : OL;
    if (const char* dbg = getenv("EFFIMVS_TC_DEBUG")) const_cast<ConvProgram&>(P).debug = atoi(dbg);
    launch_kernel(conv_tc_kernel, grid, dim3(CTA_THREADS), smem, st, P, (const uint4*)in, IL.batch_stride, (const uint8_t*)wpk, bias, OL, (uint4*)out, rl,
                                                    (const uint4*)res, out_f32, (int)n_tiles, tiles_per_plane);
    return check_launch("conv_tc_kernel");
}

int run_cin1(const float* x, const float* w, const float* bias, int B, int D, int H, int W, int s, const ActLayout& OL, int plane,
             void* out, cudaStream_t st) {
    size_t work = (size_t)D * OL.H * ((OL.W + CIN1_X - 1) / CIN1_X);
    dim3 grid((unsigned)((work + 127) / 128), B);
    if (s == 2) launch_kernel(conv_cin1_kernel<2>, grid, dim3(128), 0, st, x, w, bias, D, H, W, OL, plane, (uint4*)out);
    else launch_kernel(conv_cin1_kernel<1>, grid, dim3(128), 0, st, x, w, bias, D, H, W, OL, plane, (uint4*)out);
    return check_launch("conv_cin1_kernel");
}

// ------------------------------------------------------------------------------------------------
// 8 -> 1 channel layers (the `prob` convolution of the regularization FPN, the last transposed convolution of
// cost_up_small) on CUDA cores.  With one output channel an MMA tile would spend 15 of its 16 accumulator
// columns on padding (these layers took as long as their 8 -> 8 neighbours on the tensor pipe); as a dot
// product of 216 terms per voxel they are a few hundred FMAs per thread.  Inputs are read from the padded c8
// layout (halos are the zero padding, so no bounds checks), hi + lo halves are summed back to fp32 and the
// weights stay fp32 in shared memory.
// ------------------------------------------------------------------------------------------------
// RAW: the two 16-byte voxels hold fp32 channels 0-3 / 4-7 (ConvProgram::out_f32_pair) instead of bf16 hi / lo halves
template <bool RAW = false>
__device__ __forceinline__ void load_voxel_f32(const uint4* __restrict__ hi, long long lo_delta, long long idx, float (&v)[8]) {
    if (RAW) {
        const uint4 a = __ldg(hi + idx), c = __ldg(hi + idx + lo_delta);
        v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(c.x); v[5] = __uint_as_float(c.y); v[6] = __uint_as_float(c.z); v[7] = __uint_as_float(c.w);
        return;
    }
    unpack_bf16x8(__ldg(hi + idx), v);
    if (lo_delta) {
        float l[8];
        unpack_bf16x8(__ldg(hi + idx + lo_delta), l);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += l[j];
    }
}

constexpr int COUT1_X = 4;   // consecutive outputs of a row per thread (the convolution form)

// conv 8 -> 1, k3 p1 s1: w (1,8,3,3,3) fp32, out (B,1,D,H,W) fp32
template <bool RAW>
__global__ void __launch_bounds__(128)
conv_cout1_kernel(const uint4* __restrict__ in, const ActLayout IL, const float* __restrict__ w, const float* __restrict__ bias, int relu,
                  float* __restrict__ out) {
    __shared__ __align__(16) float sw[27 * 8];
    pdl_trigger();
    for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) sw[(i % 27) * 8 + i / 27] = w[i];   // [tap][ci]
    pdl_wait();
    __syncthreads();
    const int b = blockIdx.y, D = IL.D, H = IL.H, W = IL.W;
    const int Wq = (W + COUT1_X - 1) / COUT1_X;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (size_t)D * H * Wq) return;
    const int x0 = (int)(o % Wq) * COUT1_X, y = (int)((o / Wq) % H), z = (int)(o / ((size_t)Wq * H));
    const long long lo_delta = (long long)IL.lo_off * IL.vs;
    const long long base = (long long)b * IL.batch_stride + IL.guard;   // plane 0 (hi) of batch item b
    float acc[COUT1_X];
#pragma unroll
    for (int i = 0; i < COUT1_X; ++i) acc[i] = bias ? __ldg(bias) : 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            // padded coordinates: input voxel (z + a - 1, y + bb - 1, x + c - 1) sits at (z + a, y + bb, x + c)
            const long long row = base + (long long)(z + a) * IL.zstride + (long long)(y + bb) * IL.Px + x0;
            float v[COUT1_X + 2][8];
#pragma unroll
            for (int j = 0; j < COUT1_X + 2; ++j) {
                if (x0 + j <= W + 1) load_voxel_f32<RAW>(in, lo_delta, row + j, v[j]);   // padded x up to W + 1
                else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[j][c] = 0.0f;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 w0 = *reinterpret_cast<const float4*>(sw + (a * 9 + bb * 3 + c) * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sw + (a * 9 + bb * 3 + c) * 8 + 4);
#pragma unroll
                for (int i = 0; i < COUT1_X; ++i) {
                    const float (&u)[8] = v[i + c];
                    float t = acc[i];
                    t = fmaf(u[0], w0.x, t); t = fmaf(u[1], w0.y, t); t = fmaf(u[2], w0.z, t); t = fmaf(u[3], w0.w, t);
                    t = fmaf(u[4], w1.x, t); t = fmaf(u[5], w1.y, t); t = fmaf(u[6], w1.z, t); t = fmaf(u[7], w1.w, t);
                    acc[i] = t;
                }
            }
        }
    float* op = out + (((size_t)b * D + z) * H + y) * W + x0;
#pragma unroll
    for (int i = 0; i < COUT1_X; ++i)
        if (x0 + i < W) op[i] = relu ? fmaxf(acc[i], 0.0f) : acc[i];
}

// transposed conv 8 -> 1, k3, stride (1,2,2), padding 1, output_padding (0,1,1): w (8,1,3,3,3) fp32,
// out (B,1,D,2H,2W) fp32.  A thread owns one input position and writes its 2x2 output parities:
// out[z, 2y+py, 2x+px] = sum over a and the taps b, c of that parity (b = 1 <-> input y; b = 0 <-> y + 1; b = 2 <-> y)
template <bool RAW>
__global__ void __launch_bounds__(128)
deconv_cout1_kernel(const uint4* __restrict__ in, const ActLayout IL, const float* __restrict__ w, const float* __restrict__ bias, int relu,
                    float* __restrict__ out) {
    __shared__ __align__(16) float sw[27 * 8];
    pdl_trigger();
    for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) sw[(i % 27) * 8 + i / 27] = w[i];   // w[ci][0][tap] -> [tap][ci]
    pdl_wait();
    __syncthreads();
    const int b = blockIdx.y, D = IL.D, H = IL.H, W = IL.W;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (size_t)D * H * W) return;
    const int x = (int)(o % W), y = (int)((o / W) % H), z = (int)(o / ((size_t)W * H));
    const long long lo_delta = (long long)IL.lo_off * IL.vs;
    const long long base = (long long)b * IL.batch_stride + IL.guard;
    const float b0 = bias ? __ldg(bias) : 0.0f;
    float acc[2][2] = {{b0, b0}, {b0, b0}};
    // ncu: long scoreboard 13 warps per issue -- ptxas interleaves the voxel loads with the FMAs of earlier voxels to save registers
    // (whatever the source order), so a thread waits on one L2 round trip after the other.  The lines of all three input planes are
    // therefore requested up front with L1 prefetches (no destination registers); the x + 2 voxel of a thread is its neighbour's x + 1.
    {
        const bool last = (threadIdx.x & 31) == 31 || x == W - 1;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const uint4* p = in + base + (long long)(z + 2 - a) * IL.zstride + (long long)(y + 1 + dy) * IL.Px + (x + 1);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
                if (RAW || lo_delta) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + lo_delta));
                if (last) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 1));
                    if (RAW || lo_delta) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 1 + lo_delta));
                }
            }
    }
    float v[3][2][2][8];   // plane a, input (y + dy, x + dx)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // out z = in z - 1 + a  ->  input plane z + 1 - a, padded index z + 2 - a
        const long long plane = base + (long long)(z + 2 - a) * IL.zstride;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) load_voxel_f32<RAW>(in, lo_delta, plane + (long long)(y + 1 + dy) * IL.Px + (x + 1 + dx), v[a][dy][dx]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int py = bb != 1, dy = bb == 0, px = c != 1, dx = c == 0;
                const float4 w0 = *reinterpret_cast<const float4*>(sw + (a * 9 + bb * 3 + c) * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sw + (a * 9 + bb * 3 + c) * 8 + 4);
                const float (&u)[8] = v[a][dy][dx];
                float t = acc[py][px];
                t = fmaf(u[0], w0.x, t); t = fmaf(u[1], w0.y, t); t = fmaf(u[2], w0.z, t); t = fmaf(u[3], w0.w, t);
                t = fmaf(u[4], w1.x, t); t = fmaf(u[5], w1.y, t); t = fmaf(u[6], w1.z, t); t = fmaf(u[7], w1.w, t);
                acc[py][px] = t;
            }
    }
    const int Ho = 2 * H, Wo = 2 * W;
#pragma unroll
    for (int py = 0; py < 2; ++py) {
        float2 r = make_float2(acc[py][0], acc[py][1]);
        if (relu) { r.x = fmaxf(r.x, 0.0f); r.y = fmaxf(r.y, 0.0f); }
        *reinterpret_cast<float2*>(out + (((size_t)b * D + z) * Ho + 2 * y + py) * Wo + 2 * x) = r;
    }
}

int run_cout1(bool deconv, int B, const void* in, const ActLayout& IL, const float* w, const float* bias, int relu, float* out,
              cudaStream_t st, bool raw_f32 = false) {
    if (deconv) {
        const size_t work = (size_t)IL.D * IL.H * IL.W;
        const dim3 grid((unsigned)((work + 127) / 128), B);
        if (raw_f32) launch_kernel(deconv_cout1_kernel<true>, grid, dim3(128), 0, st, (const uint4*)in, IL, w, bias, relu, out);
        else launch_kernel(deconv_cout1_kernel<false>, grid, dim3(128), 0, st, (const uint4*)in, IL, w, bias, relu, out);
        return check_launch("deconv_cout1_kernel");
    }
    const size_t work = (size_t)IL.D * IL.H * ((IL.W + COUT1_X - 1) / COUT1_X);
    const dim3 grid((unsigned)((work + 127) / 128), B);
    if (raw_f32) launch_kernel(conv_cout1_kernel<true>, grid, dim3(128), 0, st, (const uint4*)in, IL, w, bias, relu, out);
    else launch_kernel(conv_cout1_kernel<false>, grid, dim3(128), 0, st, (const uint4*)in, IL, w, bias, relu, out);
    return check_launch("conv_cout1_kernel");
}

struct Carver {
    char* base; size_t off, cap;
    void* take(size_t bytes) { void* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; }
};

constexpr size_t W_SLOT = 128 * 1024;  // packed weights of one layer (<= 108 blocks * 2 * 32 * 16 B)

}  // namespace

// ------------------------------------------------------------------------------------------------
// a9 on tensor cores
// ------------------------------------------------------------------------------------------------
size_t costreg_bf16_workspace_bytes(int B, int D, int H, int W, bool hilo) {
    ActLayout c0 = make_layout(L_REG, 8, D, H, W, hilo), c1 = make_layout(L_SPLIT, 8, D, H, W, hilo);
    ActLayout c2 = make_layout(L_REG, 16, D / 2, H / 2, W / 2, hilo), c3 = make_layout(L_SPLIT, 16, D / 2, H / 2, W / 2, hilo);
    ActLayout c4 = make_layout(L_REG, 32, D / 4, H / 4, W / 4, hilo);
    return layout_bytes(c0, B) * 2 + layout_bytes(c1, B) + layout_bytes(c2, B) * 2 + layout_bytes(c3, B) + layout_bytes(c4, B) * 2 +
           8 * W_SLOT + 4096;
}

int costreg_bf16(const float* x, const float* const* weights, const float* const* biases, int B, int D, int H, int W, bool hilo,
                 int phases, void* ws, size_t ws_bytes, float* prob_out, cudaStream_t st) {
    EFFI_REQUIRE(D % 4 == 0 && H % 4 == 0 && W % 4 == 0, EFFIMVS_EUNSUPPORTED, "costreg bf16: D, H, W must be multiples of 4");
    EFFI_REQUIRE(D <= 65535 && B <= 65535, EFFIMVS_EUNSUPPORTED, "costreg bf16: D or B too large");
    const int D2 = D / 2, H2 = H / 2, W2 = W / 2, D4 = D / 4, H4 = H / 4, W4 = W / 4;
    ActLayout L0 = make_layout(L_REG, 8, D, H, W, hilo), L1 = make_layout(L_SPLIT, 8, D, H, W, hilo);
    ActLayout L2 = make_layout(L_REG, 16, D2, H2, W2, hilo), L3 = make_layout(L_SPLIT, 16, D2, H2, W2, hilo);
    ActLayout L4 = make_layout(L_REG, 32, D4, H4, W4, hilo), L5 = L4, L6 = L2, L7 = L0;
    Carver cv{(char*)ws, 0, ws_bytes};
    void* c0 = cv.take(layout_bytes(L0, B)); void* c7 = cv.take(layout_bytes(L7, B)); void* c1 = cv.take(layout_bytes(L1, B));
    void* c2 = cv.take(layout_bytes(L2, B)); void* c6 = cv.take(layout_bytes(L6, B)); void* c3 = cv.take(layout_bytes(L3, B));
    void* c4 = cv.take(layout_bytes(L4, B)); void* c5 = cv.take(layout_bytes(L5, B));
    const size_t act_bytes = cv.off;
    void* wp[8];
    for (int i = 0; i < 8; ++i) wp[i] = cv.take(W_SLOT);
    EFFI_REQUIRE(cv.off <= ws_bytes, EFFIMVS_EWORKSPACE, "costreg bf16: workspace %zu < %zu", ws_bytes, cv.off);
    (void)act_bytes;
    int rc;
    const bool prepare = (phases & EFFIMVS_WS_PREPARE) != 0, run = (phases & EFFIMVS_WS_RUN) != 0;
    if (prepare) {   // halos and guards must read as zero (the layers only ever store to interior positions)
        void* bufs[8] = {c0, c7, c1, c2, c6, c3, c4, c5};
        ActLayout lays[8] = {L0, L7, L1, L2, L6, L3, L4, L5};
        if ((rc = run_zero_halo(8, bufs, lays, B, st))) return rc;
    }

    static thread_local ConvProgram P[8];
    static thread_local PackTable T[8];
    bool ok = build_conv_s1(P[0], T[0], 8, 8, L0, 1) && build_conv_s2(P[1], T[1], 8, 16, L1, 1) &&
              build_conv_s1(P[2], T[2], 16, 16, L2, 1) && build_conv_s2(P[3], T[3], 16, 32, L3, 1) &&
              build_conv_s1(P[4], T[4], 32, 32, L4, 1) && build_deconv(P[5], T[5], 32, 16, L5, 2, 1) &&
              build_deconv(P[6], T[6], 16, 8, L6, 2, 1) && build_conv_s1(P[7], T[7], 8, 1, L7, 0);
    EFFI_REQUIRE(ok, EFFIMVS_EUNSUPPORTED, "costreg bf16: program does not fit");
    for (int i = 0; i < 8; ++i)
        EFFI_REQUIRE(program_weight_bytes(T[i]) <= W_SLOT, EFFIMVS_EUNSUPPORTED, "costreg bf16: packed weights of layer %d too large", i + 1);
    if (prepare && (rc = run_pack_multi(8, T, weights + 1, wp, st))) return rc;
    if (!run) return EFFIMVS_OK;
    if ((rc = run_cin1(x, weights[0], biases[0], B, D, H, W, 1, L0, 0, c0, st))) return rc;
    if ((rc = run_tile_kernel(P[0], B, c0, L0, wp[0], biases[1], L1, c1, nullptr, nullptr, nullptr, st))) return rc;
    if ((rc = run_tile_kernel(P[1], B, c1, L1, wp[1], biases[2], L2, c2, nullptr, nullptr, nullptr, st))) return rc;
    if ((rc = run_tile_kernel(P[2], B, c2, L2, wp[2], biases[3], L3, c3, nullptr, nullptr, nullptr, st))) return rc;
    if ((rc = run_tile_kernel(P[3], B, c3, L3, wp[3], biases[4], L4, c4, nullptr, nullptr, nullptr, st))) return rc;
    if ((rc = run_tile_kernel(P[4], B, c4, L4, wp[4], biases[5], L5, c5, nullptr, nullptr, nullptr, st))) return rc;
    if ((rc = run_tile_kernel(P[5], B, c5, L5, wp[5], biases[6], L6, c6, &L3, c3, nullptr, st))) return rc;
    // conv7's only consumer is the 8 -> 1 prob layer.  EFFIMVS_PROB_PATH=raw: conv7 stores its eight fp32 channels as they are
    // (two 16-byte voxels of the hi/lo geometry) and prob runs on CUDA cores without unpacking; default: prob as a z-sweep
    // tensor-core program on hi/lo bf16 (see the measurements in DESIGN.md)
    const char* prob_path = getenv("EFFIMVS_PROB_PATH");
    const bool raw_prob = hilo && prob_path && !strcmp(prob_path, "raw");
    P[6].out_f32_pair = raw_prob ? 1 : 0;
    if ((rc = run_tile_kernel(P[6], B, c6, L6, wp[6], biases[7], L7, c7, &L1, c1, nullptr, st))) return rc;
    if (raw_prob) return run_cout1(false, B, c7, L7, weights[8], nullptr, 0, prob_out, st, true);
    // measured: the z-sweep tensor-core program (0.49 ms for the net) beats the CUDA-core form on hi/lo input (0.52 ms) for this stride-1 layer
    if (getenv("EFFIMVS_CUDA_CORE_PROB")) return run_cout1(false, B, c7, L7, weights[8], nullptr, 0, prob_out, st);
    return run_tile_kernel(P[7], B, c7, L7, wp[7], nullptr, L7, nullptr, nullptr, nullptr, prob_out, st);  // fp32 out, logical dims of L7
}

// ------------------------------------------------------------------------------------------------
// a10 on tensor cores
// ------------------------------------------------------------------------------------------------
size_t cost_up_bf16_workspace_bytes(int B, int D, int H, int W, bool hilo) {
    ActLayout cat = make_layout(L_REG, 16, D, H / 2, W / 2, hilo), c1 = make_layout(L_REG, 8, D, H / 2, W / 2, hilo);
    return layout_bytes(cat, B) + layout_bytes(c1, B) + 2 * W_SLOT + 4096;
}

int cost_up_bf16(const float* x, const float* prev, const float* const* weights, const float* const* biases, int B, int D, int H,
                 int W, bool hilo, int phases, void* ws, size_t ws_bytes, float* out, cudaStream_t st) {
    EFFI_REQUIRE(D <= 65535 && B <= 65535, EFFIMVS_EUNSUPPORTED, "cost_up bf16: D or B too large");
    const int H2 = H / 2, W2 = W / 2;
    ActLayout Lcat = make_layout(L_REG, 16, D, H2, W2, hilo), L1 = make_layout(L_REG, 8, D, H2, W2, hilo);
    Carver cv{(char*)ws, 0, ws_bytes};
    void* cat = cv.take(layout_bytes(Lcat, B)); void* c1 = cv.take(layout_bytes(L1, B));
    const size_t act_bytes = cv.off;
    void* wp1 = cv.take(W_SLOT); void* wp2 = cv.take(W_SLOT);
    EFFI_REQUIRE(cv.off <= ws_bytes, EFFIMVS_EWORKSPACE, "cost_up bf16: workspace %zu < %zu", ws_bytes, cv.off);
    (void)act_bytes;
    int rc;
    const bool prepare = (phases & EFFIMVS_WS_PREPARE) != 0, run = (phases & EFFIMVS_WS_RUN) != 0;
    if (prepare) {
        void* bufs[2] = {cat, c1};
        ActLayout lays[2] = {Lcat, L1};
        if ((rc = run_zero_halo(2, bufs, lays, B, st))) return rc;
    }
    static thread_local ConvProgram P1, P2;
    static thread_local PackTable T12[2];
    PackTable &T1 = T12[0], &T2 = T12[1];
    bool ok = build_conv_s1(P1, T1, 16, 8, Lcat, 1) && build_deconv(P2, T2, 8, 1, L1, 1, 1);
    EFFI_REQUIRE(ok, EFFIMVS_EUNSUPPORTED, "cost_up bf16: program does not fit");
    if (prepare) {
        void* dsts[2] = {wp1, wp2};
        if ((rc = run_pack_multi(2, T12, weights + 2, dsts, st))) return rc;
    }
    if (!run) return EFFIMVS_OK;
    // in a hi/lo layout the concatenated tensor has planes [conv0 hi, conv_cost hi, conv0 lo, conv_cost lo]
    if ((rc = run_cin1(x, weights[0], biases[0], B, D, H, W, 2, Lcat, 0, cat, st))) return rc;
    if ((rc = run_cin1(prev, weights[1], biases[1], B, D, H2, W2, 1, Lcat, 1, cat, st))) return rc;
    // conv1's only consumer is the CUDA-core 8 -> 1 layer: in a hi/lo layout (two 16-byte voxels per position) it gets the
    // eight fp32 values as they are instead of bf16 hi + lo halves (same bytes, exact, no unpacking on the other side)
    const bool tc_cout1 = getenv("EFFIMVS_TC_COUT1") != nullptr;
    const bool raw = hilo && !tc_cout1 && !getenv("EFFIMVS_NO_RAW_F32");
    P1.out_f32_pair = raw ? 1 : 0;
    if ((rc = run_tile_kernel(P1, B, cat, Lcat, wp1, biases[2], L1, c1, nullptr, nullptr, nullptr, st))) return rc;
    ActLayout LO = make_layout(L_REG, 8, D, H, W, false);  // fp32 output indexed with the logical (full-resolution) dims
    if (!tc_cout1) return run_cout1(true, B, c1, L1, weights[3], biases[3], 1, out, st, raw);        // 8 -> 1: CUDA cores
    return run_tile_kernel(P2, B, c1, L1, wp2, biases[3], LO, nullptr, nullptr, nullptr, out, st);
}

// ------------------------------------------------------------------------------------------------
// single layer (tests, other callers): fp32 NCDHW in/out around one tensor-core layer
// ------------------------------------------------------------------------------------------------
size_t conv3d_bf16_workspace_bytes(int B, int Cin, int Cout, int D, int H, int W, int sd, int transposed, bool hilo) {
    int Do = D, Ho = H, Wo = W;
    if (transposed) { Do = D * sd; Ho = H * 2; Wo = W * 2; }
    else if (sd == 2) { Do = D / 2; Ho = H / 2; Wo = W / 2; }
    ActLayout LI = make_layout(!transposed && sd == 2 ? L_SPLIT : L_REG, (Cin + 7) / 8 * 8, D, H, W, hilo);
    ActLayout LO = make_layout(L_REG, (Cout + 7) / 8 * 8, Do, Ho, Wo, hilo);
    return layout_bytes(LI, B) + 2 * layout_bytes(LO, B) + W_SLOT + 4096;
}

int conv3d_bf16(const float* x, const float* weight, const float* bias, const float* residual, int B, int Cin, int Cout, int D,
                int H, int W, int sd, int transposed, int relu, bool hilo, void* ws, size_t ws_bytes, float* y, cudaStream_t st) {
    EFFI_REQUIRE((Cin == 8 || Cin == 16 || Cin == 32) && Cout >= 1 && Cout <= 32, EFFIMVS_EUNSUPPORTED,
                 "conv3d_bf16: Cin in {8,16,32}, Cout <= 32");
    int Do = D, Ho = H, Wo = W;
    if (transposed) { Do = D * sd; Ho = H * 2; Wo = W * 2; }
    else if (sd == 2) { EFFI_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, EFFIMVS_EUNSUPPORTED, "conv3d_bf16: stride 2 needs even dims"); Do = D / 2; Ho = H / 2; Wo = W / 2; }
    const int Cop = (Cout + 7) / 8 * 8;
    ActLayout LI = make_layout(!transposed && sd == 2 ? L_SPLIT : L_REG, Cin, D, H, W, hilo);
    ActLayout LO = make_layout(L_REG, Cop, Do, Ho, Wo, hilo);
    Carver cv{(char*)ws, 0, ws_bytes};
    void* xin = cv.take(layout_bytes(LI, B)); void* yout = cv.take(layout_bytes(LO, B)); void* rin = cv.take(layout_bytes(LO, B));
    const size_t act_bytes = cv.off;
    void* wp = cv.take(W_SLOT);
    EFFI_REQUIRE(cv.off <= ws_bytes, EFFIMVS_EWORKSPACE, "conv3d_bf16: workspace %zu < %zu", ws_bytes, cv.off);
    if (cudaMemsetAsync(ws, 0, act_bytes, st) != cudaSuccess) { set_error("conv3d_bf16: memset failed"); return EFFIMVS_ECUDA; }
    static thread_local ConvProgram P;
    static thread_local PackTable T;
    bool ok = transposed ? build_deconv(P, T, Cin, Cout, LI, sd, relu) : (sd == 2 ? build_conv_s2(P, T, Cin, Cout, LI, relu) : build_conv_s1(P, T, Cin, Cout, LI, relu));
    EFFI_REQUIRE(ok && program_weight_bytes(T) <= W_SLOT, EFFIMVS_EUNSUPPORTED, "conv3d_bf16: program does not fit");
    int rc;
    if ((rc = run_pack(T, weight, wp, st))) return rc;
    size_t n_in = (size_t)D * H * W * (Cin / 8), n_out = (size_t)Do * Ho * Wo * (Cop / 8);
    to_c8_kernel<<<dim3((unsigned)((n_in + 255) / 256), B), 256, 0, st>>>(x, Cin, LI, (uint4*)xin);
    if (residual) to_c8_kernel<<<dim3((unsigned)((n_out + 255) / 256), B), 256, 0, st>>>(residual, Cout, LO, (uint4*)rin);
    if ((rc = check_launch("to_c8_kernel"))) return rc;
    if ((rc = run_tile_kernel(P, B, xin, LI, wp, bias, LO, yout, residual ? &LO : nullptr, residual ? rin : nullptr, nullptr, st))) return rc;
    from_c8_kernel<<<dim3((unsigned)((n_out + 255) / 256), B), 256, 0, st>>>((const uint4*)yout, Cout, LO, y);
    return check_launch("from_c8_kernel");
}

}  // namespace effimvs

extern "C" size_t effimvs_conv3d_bf16_workspace_bytes(int B, int Cin, int Cout, int D, int H, int W, int sd, int transposed, int precision) {
    return effimvs::conv3d_bf16_workspace_bytes(B, Cin, Cout, D, H, W, sd, transposed, precision == EFFIMVS_PREC_BF16X3);
}
extern "C" int effimvs_conv3d_bf16(const float* x, const float* weight, const float* bias, const float* residual, int B, int Cin,
                                   int Cout, int D, int H, int W, int sd, int transposed, int relu, int precision, void* workspace,
                                   size_t workspace_bytes, float* y, void* stream) {
    EFFI_REQUIRE(x && weight && y && workspace, EFFIMVS_EINVAL, "conv3d_bf16: null pointer");
    EFFI_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && (sd == 1 || sd == 2), EFFIMVS_EINVAL, "conv3d_bf16: bad sizes");
    EFFI_REQUIRE(precision == EFFIMVS_PREC_BF16 || precision == EFFIMVS_PREC_BF16X3, EFFIMVS_EINVAL, "conv3d_bf16: precision=%d", precision);
    return effimvs::conv3d_bf16(x, weight, bias, residual, B, Cin, Cout, D, H, W, sd, transposed, relu, precision == EFFIMVS_PREC_BF16X3,
                                workspace, workspace_bytes, y, (cudaStream_t)stream);
}
