// fp32 CUDA-core 3x3x3 convolution / transposed convolution with the eval-mode BatchNorm folded
// into weight and bias, fused bias + ReLU + residual epilogue and channel-offset output (fuses
// the torch.cat of cost_up_small).
//
//   Conv3d / Deconv3d wrappers     upstream models/module.py:124-209
//
// This is the exact-parity (fp32) path of the regularization nets and the reference the bf16
// tcgen05 implicit-GEMM path (conv3d_tc.cu) is validated against on the device.  One thread per
// output voxel (x fastest -> coalesced NCDHW loads), all COUT accumulators in registers, weights
// staged once per block in shared memory as [ci][tap][co] and read as broadcast float4.
#include "common.cuh"

namespace effimvs {
namespace {

template <int COUT, bool TRANSPOSED>
__global__ void __launch_bounds__(128)
conv3d_f32_kernel(const float* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias,
                  const float* __restrict__ residual, int Cin, int D, int H, int W, int Do, int Ho, int Wo,
                  int sd, int sh, int sw, int relu, float* __restrict__ y, int y_coff, int y_ctot) {
    extern __shared__ __align__(16) float s_w[];  // [Cin][27][COUT]
    const int n_w = Cin * 27 * COUT;
    for (int i = threadIdx.x; i < n_w; i += blockDim.x) {
        int co = i % COUT, tap = (i / COUT) % 27, ci = i / (27 * COUT);
        s_w[i] = TRANSPOSED ? weight[((size_t)ci * COUT + co) * 27 + tap] : weight[((size_t)co * Cin + ci) * 27 + tap];
    }
    __syncthreads();
    const int b = blockIdx.y;
    const size_t vox_in = (size_t)D * H * W, vox_out = (size_t)Do * Ho * Wo;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= vox_out) return;
    const int ox = (int)(o % Wo), oy = (int)((o / Wo) % Ho), oz = (int)(o / ((size_t)Wo * Ho));

    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;

    // per-dimension input index and validity of the three taps
    int iz[3], iy[3], ix[3];
    bool vz[3], vy[3], vx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (TRANSPOSED) {
            int tz = oz + 1 - k, ty = oy + 1 - k, tx = ox + 1 - k;
            vz[k] = tz >= 0 && tz % sd == 0 && tz / sd < D; iz[k] = tz / sd;
            vy[k] = ty >= 0 && ty % sh == 0 && ty / sh < H; iy[k] = ty / sh;
            vx[k] = tx >= 0 && tx % sw == 0 && tx / sw < W; ix[k] = tx / sw;
        } else {
            iz[k] = oz * sd + k - 1; vz[k] = iz[k] >= 0 && iz[k] < D;
            iy[k] = oy * sh + k - 1; vy[k] = iy[k] >= 0 && iy[k] < H;
            ix[k] = ox * sw + k - 1; vx[k] = ix[k] >= 0 && ix[k] < W;
        }
    }
    const float* xb = x + (size_t)b * Cin * vox_in;
    for (int ci = 0; ci < Cin; ++ci) {
        const float* xc = xb + (size_t)ci * vox_in;
        const float* wc = s_w + (size_t)ci * 27 * COUT;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const bool ok = vz[a] && vy[bb] && vx[c];
                    const float v = ok ? __ldg(xc + ((size_t)iz[a] * H + iy[bb]) * W + ix[c]) : 0.0f;
                    const float* wt = wc + (a * 9 + bb * 3 + c) * COUT;
                    if (COUT % 4 == 0) {
#pragma unroll
                        for (int co = 0; co < COUT; co += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wt + co);
                            acc[co] = fmaf(v, w4.x, acc[co]);
                            acc[co + 1] = fmaf(v, w4.y, acc[co + 1]);
                            acc[co + 2] = fmaf(v, w4.z, acc[co + 2]);
                            acc[co + 3] = fmaf(v, w4.w, acc[co + 3]);
                        }
                    } else {
#pragma unroll
                        for (int co = 0; co < COUT; ++co) acc[co] = fmaf(v, wt[co], acc[co]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float r = acc[co] + (bias ? __ldg(bias + co) : 0.0f);
        if (relu) r = fmaxf(r, 0.0f);
        if (residual) r += __ldg(residual + ((size_t)b * COUT + co) * vox_out + o);
        y[((size_t)b * y_ctot + y_coff + co) * vox_out + o] = r;
    }
}

template <int COUT, bool T>
int launch(const float* x, const float* w, const float* bias, const float* res, int B, int Cin, int D, int H, int W,
           int Do, int Ho, int Wo, int sd, int sh, int sw, int relu, float* y, int y_coff, int y_ctot, cudaStream_t st) {
    size_t smem = (size_t)Cin * 27 * COUT * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv3d_f32_kernel<COUT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("conv3d_f32: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return EFFIMVS_ECUDA;
        }
    }
    size_t vox = (size_t)Do * Ho * Wo;
    dim3 block(128), grid((unsigned)((vox + 127) / 128), B);
    conv3d_f32_kernel<COUT, T><<<grid, block, smem, st>>>(x, w, bias, res, Cin, D, H, W, Do, Ho, Wo, sd, sh, sw, relu, y,
                                                         y_coff, y_ctot);
    return check_launch("conv3d_f32_kernel");
}

}  // namespace

int conv3d_f32(const float* x, const float* weight, const float* bias, const float* residual, int B, int Cin, int Cout,
               int D, int H, int W, int sd, int sh, int sw, int transposed, int relu, float* y, int y_coff, int y_ctot,
               cudaStream_t st) {
    EFFI_REQUIRE(x && weight && y, EFFIMVS_EINVAL, "conv3d_f32: null pointer");
    EFFI_REQUIRE(B > 0 && Cin > 0 && D > 0 && H > 0 && W > 0, EFFIMVS_EINVAL, "conv3d_f32: bad sizes");
    EFFI_REQUIRE((sd == 1 || sd == 2) && (sh == 1 || sh == 2) && (sw == 1 || sw == 2), EFFIMVS_EUNSUPPORTED,
                 "conv3d_f32: strides must be 1 or 2");
    EFFI_REQUIRE(y_coff >= 0 && y_coff + Cout <= y_ctot, EFFIMVS_EINVAL, "conv3d_f32: channel window outside output");
    EFFI_REQUIRE((size_t)Cin * 27 * Cout * 4 <= 200 * 1024, EFFIMVS_EUNSUPPORTED, "conv3d_f32: Cin*Cout too large");
    int Do, Ho, Wo;
    if (transposed) { Do = D * sd; Ho = H * sh; Wo = W * sw; }
    else { Do = (D - 1) / sd + 1; Ho = (H - 1) / sh + 1; Wo = (W - 1) / sw + 1; }
#define EFFI_CONV_CASE(CO)                                                                                          \
    case CO:                                                                                                        \
        return transposed ? launch<CO, true>(x, weight, bias, residual, B, Cin, D, H, W, Do, Ho, Wo, sd, sh, sw, relu, y, y_coff, y_ctot, st) \
                          : launch<CO, false>(x, weight, bias, residual, B, Cin, D, H, W, Do, Ho, Wo, sd, sh, sw, relu, y, y_coff, y_ctot, st);
    switch (Cout) {
        EFFI_CONV_CASE(1)
        EFFI_CONV_CASE(8)
        EFFI_CONV_CASE(16)
        EFFI_CONV_CASE(32)
    }
#undef EFFI_CONV_CASE
    set_error("conv3d_f32: Cout=%d not in {1,8,16,32}", Cout);
    return EFFIMVS_EUNSUPPORTED;
}

}  // namespace effimvs

extern "C" int effimvs_conv3d_f32(const float* x, const float* weight, const float* bias, const float* residual,
                                  int B, int Cin, int Cout, int D, int H, int W, int sd, int sh, int sw,
                                  int transposed, int relu, float* y, int y_coff, int y_ctot, void* stream) {
    return effimvs::conv3d_f32(x, weight, bias, residual, B, Cin, Cout, D, H, W, sd, sh, sw, transposed, relu, y, y_coff,
                               y_ctot, (cudaStream_t)stream);
}
