// Per-pixel preprocessing kernels (HBM-bound):
//   * GMA_float_conv2 (MIMC_module.c:2517-2585) with its in-place, stale-border semantics
//     (SURVEY.md H6): interior stencil with NaN-on-zero-pixel, global minimum over the WHOLE
//     output buffer (stale border values included, NaN skipped), shift so valid pixels are
//     >= 1 and NaN -> 0 over rows [ocwy, H-ocwy) x columns [ocwx, W).
//   * u8/u16 -> f32 ingest cast (GMA_float_load_tiff's per-pixel loops, GMA.c:288-310).
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kMinBlocks = 1024;   // partial minima written by the stencil pass
constexpr int kThreads = 256;

__device__ __forceinline__ float block_min(float v) {
    __shared__ float red[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 1e+37f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;   // valid in thread 0
}

// dn_in = (int32_t)(px + 0.5) ? px : NaN   (MIMC_module.c:2548; the +0.5 is in double)
__device__ __forceinline__ float null_to_nan(float px) {
    double t = (double)px + 0.5;
    return (t > -1.0 && t < 1.0) ? CUDART_NAN_F : px;
}

// Pass 1: stencil over the interior (written to out), every other pixel keeps its stale value;
// the block minimum over ALL pixels of `out` after this pass goes to partial[blockIdx.x].
// NOTE: `fminf` drops NaN operands, which matches the reference's `if (out < dn_min)` scan.
__global__ void __launch_bounds__(kThreads) conv2_stencil_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                                 int H, int W, int kh, int kw, float k0, float k1,
                                                                 float k2, float k3, float k4, float k5, float k6,
                                                                 float k7, float k8, float *__restrict__ partial) {
    const float kk[9] = {k0, k1, k2, k3, k4, k5, k6, k7, k8};
    const int ocwx = kw / 2, ocwy = kh / 2;
    const size_t total = (size_t)H * W;
    float vmin = 1e+37f;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
        int r = (int)(i / W), c = (int)(i - (size_t)r * W);
        float v;
        if (r >= ocwy && r < H - ocwy && c >= ocwx && c < W - ocwx) {
            float sum = 0.0f;
            for (int a = 0; a < kh; a++)
                for (int b = 0; b < kw; b++) {
                    float px = __ldg(&in[(size_t)(r + a - ocwy) * W + (c + b - ocwx)]);
                    sum = __fadd_rn(sum, __fmul_rn(null_to_nan(px), kk[a * kw + b]));
                }
            out[i] = sum;
            v = sum;
        } else {
            v = out[i];
        }
        vmin = fminf(vmin, v);
    }
    vmin = block_min(vmin);
    if (threadIdx.x == 0) partial[blockIdx.x] = vmin;
}

// Pass 2: out -= dn_min - 1 (NaN -> 0) over rows [ocwy, H-ocwy) and columns [ocwx, W).
__global__ void __launch_bounds__(kThreads) conv2_shift_kernel(float *__restrict__ out, int H, int W, int kh, int kw,
                                                               const float *__restrict__ partial, int npartial) {
    __shared__ float s_min;
    float v = 1e+37f;
    for (int i = threadIdx.x; i < npartial; i += kThreads) v = fminf(v, partial[i]);
    v = block_min(v);
    if (threadIdx.x == 0) s_min = v;
    __syncthreads();
    const float shift = __fsub_rn(s_min, 1.0f);
    const int ocwx = kw / 2, ocwy = kh / 2;
    const size_t total = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
        int r = (int)(i / W), c = (int)(i - (size_t)r * W);
        if (r >= ocwy && r < H - ocwy && c >= ocwx) {   // right border included, left excluded (:2570-2572)
            float o = out[i];
            out[i] = isnan(o) ? 0.0f : __fsub_rn(o, shift);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) cast_kernel(const T *__restrict__ src, float *__restrict__ dst, size_t count) {
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (size_t)gridDim.x * kThreads)
        dst[i] = (float)src[i];
}

}  // namespace

int launch_conv2(mimc3cu_ctx *ctx, const float *src, int32_t H, int32_t W, const float *kernel, int32_t kh, int32_t kw,
                 float *dst) {
    if (!((kh == 1 || kh == 3) && (kw == 1 || kw == 3))) return mimc3cu_fail(ctx, "conv2: kernel must be 1x3, 3x1 or 3x3");
    float k[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < kh * kw; i++) k[i] = kernel[i];
    size_t total = (size_t)H * W;
    int blocks = (int)((total + kThreads - 1) / kThreads);
    if (blocks > kMinBlocks) blocks = kMinBlocks;
    conv2_stencil_kernel<<<blocks, kThreads, 0, ctx->stream>>>(src, dst, H, W, kh, kw, k[0], k[1], k[2], k[3], k[4], k[5],
                                                             k[6], k[7], k[8], ctx->minbuf);
    CU_CHECK(ctx, cudaGetLastError());
    int blocks2 = (int)((total + kThreads - 1) / kThreads);
    int cap = ctx->num_sms * 8;
    if (blocks2 > cap) blocks2 = cap;
    conv2_shift_kernel<<<blocks2, kThreads, 0, ctx->stream>>>(dst, H, W, kh, kw, ctx->minbuf, blocks);
    CU_CHECK(ctx, cudaGetLastError());
    ctx->launches += 2;
    return 0;
}

int launch_cast_u8(mimc3cu_ctx *ctx, const uint8_t *src, float *dst, size_t count) {
    int blocks = (int)((count + kThreads - 1) / kThreads);
    int cap = ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    cast_kernel<uint8_t><<<blocks, kThreads, 0, ctx->stream>>>(src, dst, count);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

int launch_cast_u16(mimc3cu_ctx *ctx, const uint16_t *src, float *dst, size_t count) {
    int blocks = (int)((count + kThreads - 1) / kThreads);
    int cap = ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    cast_kernel<uint16_t><<<blocks, kThreads, 0, ctx->stream>>>(src, dst, count);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}
