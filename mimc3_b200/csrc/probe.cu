// Measurement helper: sustained FP32 FMA throughput of the CUDA cores, the denominator of
// the matcher's roofline ("measure with an FMA micro-benchmark on the box").
// MEASURED_PEAKS.json carries HBM and bf16 tensor numbers only.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) fma_probe_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll 8
        for (int k = 0; k < 8; k++) {
            x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

extern "C" int mimc3cu_fp32_peak(mimc3cu_ctx *ctx, double *tflops, double *ms_out) {
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->num_sms * 8, threads = 256, iters = 4096;
    if (int rc = ensure_scratch(ctx, (size_t)blocks * threads * sizeof(float))) return rc;
    cudaEvent_t e0, e1;
    CU_CHECK(ctx, cudaEventCreate(&e0));
    CU_CHECK(ctx, cudaEventCreate(&e1));
    double best = 0.0, best_ms = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CU_CHECK(ctx, cudaEventRecord(e0, ctx->stream));
        fma_probe_kernel<<<blocks, threads, 0, ctx->stream>>>((float *)ctx->scratch, iters, 1.0000001f, 1e-7f);
        CU_CHECK(ctx, cudaEventRecord(e1, ctx->stream));
        CU_CHECK(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        CU_CHECK(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * blocks * threads * 64.0 * iters / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) { best = tf; best_ms = ms; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (tflops) *tflops = best;
    if (ms_out) *ms_out = best_ms;
    return 0;
}
