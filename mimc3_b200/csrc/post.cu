// Postprocess kernels (cluster -> dpf0 -> dpf1 sweeps -> smoothing -> snap -> pseudosmoothing).
// Placeholder translation unit: filled in after the matcher is parity-green on the GPU.
#include "common.cuh"

struct Post { int dummy; };

int post_cluster(mimc3cu_ctx *ctx, const float *, int32_t, int32_t, float *, int32_t *) {
    return mimc3cu_fail(ctx, "cluster: not implemented yet");
}
int post_run(mimc3cu_ctx *ctx, const float *, const double *, const mimc3cu_params *, float *, int32_t *) {
    return mimc3cu_fail(ctx, "postprocess: not implemented yet");
}
int post_stage(mimc3cu_ctx *ctx, int32_t, void *) { return mimc3cu_fail(ctx, "postprocess_stage: not implemented yet"); }
int post_finalize(mimc3cu_ctx *ctx, float *, const mimc3cu_params *, float *, float *) {
    return mimc3cu_fail(ctx, "finalize: not implemented yet");
}
void post_free(mimc3cu_ctx *ctx) { delete ctx->post; ctx->post = nullptr; }
