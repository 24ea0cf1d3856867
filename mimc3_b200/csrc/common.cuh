// Internal declarations shared by the translation units of libmimc3cu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mimc3cu.h"

#define MIMC3CU_MAX_PIVOT_SLOTS 8

struct Image {
    float *d = nullptr;
    int32_t H = 0, W = 0;
    bool used = false;
    // sat.cu: statistics + summed-area table, valid until the payload changes
    bool stats_valid = false, sat_valid = false;
    bool exact_class = false;     // finite, >= 0, multiples of 2^-frac_bits, scaled max < 2^24
    float max_value = 0.0f;
    int frac_bits = 0;            // 0 (integer DN) or 3 (multiples of 1/8)
    void *sat = nullptr;          // ulonglong2[(H+1)*(W+1)]
};

struct PivotSet {
    int32_t *off = nullptr;    // device, n+1
    int32_t *piv = nullptr;    // device, total*2
    int32_t n = 0;
    int64_t total = 0;
    size_t off_cap = 0, piv_cap = 0;        // device capacities (elements), reused across set_pivots calls
    cudaEvent_t last_use = nullptr;         // recorded after every matcher launch that reads this slot: replacing the
                                            // slot only waits for those launches, not for the whole stream
    int32_t max_abs_u = 0, max_abs_v = 0;   // over the last pivot of every node
    int64_t max_cells = 0;                  // max (2|ul|+4)(2|vl|+4): reachable cmap region
    int64_t max_sarea_extra = 0;            // helper: max over nodes of (|ul|+2, |vl|+2) product terms
    std::vector<int32_t> last_u, last_v;    // host copy of |last pivot| per node (for smem sizing)
    // match2.cu: node lists per shared-memory bin, cached per chip half-width
    struct Bins {
        int32_t ocw = -1;
        int32_t *lists = nullptr;            // device: concatenated node indices
        size_t lists_cap = 0;
        int32_t count[6] = {0, 0, 0, 0, 0, 0};  // [0..4] v2 bins (match2.cu bin_table), [5] general kernel
        int32_t start[6] = {0, 0, 0, 0, 0, 0};
        int64_t grp_bytes[5] = {0, 0, 0, 0, 0}; // shared memory per node group in each v2 bin
        int64_t max_global_cells = 0;        // largest cmap (cells, padded) among the bin-2 nodes that keep it in global memory
    } bins[2];
};

// comm.cu: NCCL communicator of the node-row bands (one rank per band)
struct BandComm {
    void *comm = nullptr;      // ncclComm_t
    int rank = 0, world = 1;
    void *tmp = nullptr;       // receive buffers of the dirty-flag OR-reduction
    size_t tmp_bytes = 0;
    int64_t n_exchanges = 0, n_allreduce = 0;
    // device time of the collectives (CUDA events around them on the context stream), read by mimc3cu_comm_timing
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_exchange, ev_allreduce;
    std::vector<cudaEvent_t> ev_pool;
    double ms_exchange = 0.0, ms_allreduce = 0.0;
    bool timing = false;
};

struct mimc3cu_ctx {
    int device = 0;
    int num_sms = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t upload_stream = nullptr;   // host -> device copies of nodes / pivots / node lists (upload_sync)
    std::string err;
    int64_t launches = 0;

    std::vector<Image> images;

    // nodes
    int32_t n = 0;
    int2 *node_uv = nullptr;     // device (n)
    size_t node_cap = 0;

    PivotSet pivots[MIMC3CU_MAX_PIVOT_SLOTS];

    // matcher scratch
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    unsigned int *counter = nullptr;   // dynamic node fetch
    float *minbuf = nullptr;           // conv2 reduction
    unsigned int *statbuf = nullptr;   // image statistics (sat.cu)
    int *overflow_list = nullptr;      // nodes the v2 matcher hands to the general kernel
    size_t overflow_cap = 0;
    int matcher = 0;                   // 0 auto, 1 force the general FP64 kernel, 2 require v2
    int last_matcher = 0;              // which kernel family the last match call used (1 general, 2 exact-FP32)

    // device buffers of the control-point stage (cp.cu), grow-only
    void *cp_pool = nullptr;
    size_t cp_pool_bytes = 0;

    // postprocess state kept for mimc3cu_postprocess_stage
    struct Post *post = nullptr;
    BandComm *comm = nullptr;

    // optional per-kernel-family event timing (0 match, 1 conv2/ingest, 2 postprocess)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timers[3];
    std::vector<cudaEvent_t> event_pool;
};

// RAII event bracket: records start now and stop at scope exit when ctx->timing is on.
struct ScopedTimer {
    mimc3cu_ctx *ctx;
    int family;
    cudaEvent_t stop = nullptr;
    ScopedTimer(mimc3cu_ctx *c, int fam);
    ~ScopedTimer();
};

extern std::string g_mimc3cu_error;

int mimc3cu_fail(mimc3cu_ctx *ctx, const char *fmt, ...);

#define CU_CHECK(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return mimc3cu_fail(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,             \
                                cudaGetErrorString(e__));                                           \
    } while (0)

int ensure_scratch(mimc3cu_ctx *ctx, size_t bytes);
// Host -> device copy that has LANDED when the call returns, without waiting for the work queued on ctx->stream.
// (cudaMemcpy on the legacy stream is not ordered against the non-blocking context stream, and for pageable
// sources it may return once the data is staged; kernels launched afterwards could then read stale buffers.)
// The caller guarantees that no kernel in flight still reads `dev`.
int upload_sync(mimc3cu_ctx *ctx, void *dev, const void *host, size_t bytes);
Image *get_image(mimc3cu_ctx *ctx, int32_t handle);

// match.cu
struct MatchLaunch {
    // image mode
    const float *ref = nullptr, *srch = nullptr;
    int32_t H = 0, W = 0;
    const int2 *node_uv = nullptr;
    int32_t off_u = 0, off_v = 0;
    const int32_t *csr_off = nullptr;   // NULL => shared pivot list (explicit mode)
    const int32_t *piv = nullptr;
    int32_t sign = 1;
    // explicit mode (CP stage)
    const float *chips = nullptr, *sareas = nullptr;
    int32_t D = 0, P = 0;
    int64_t chip_stride = 0;   // 0 => S*S (dense chips)
    int32_t chip_pitch = 0;    // 0 => S
    // common
    int32_t n = 0, ocw = 0;
    float negate = 1.0f;
    float *dp = nullptr;
    int32_t *peak = nullptr, *ncell = nullptr;
    int64_t max_cells = 0;       // reachable cmap cells, max over nodes
    int64_t max_sarea = 0;       // Dx2*Dy2, max over nodes
    // optional indirection: process node_list[0 .. *list_count) instead of 0..n-1
    const int32_t *node_list = nullptr;
    const unsigned int *list_count = nullptr;   // device; NULL => list_n
    int32_t list_n = 0;
};
int launch_match(mimc3cu_ctx *ctx, const MatchLaunch &L);
// match2.cu: exact-FP32 matcher for exact-class image pairs; falls back to launch_match per node
int launch_match2(mimc3cu_ctx *ctx, const MatchLaunch &L, const Image *ref, const Image *srch, PivotSet *ps);
bool match2_supported(const MatchLaunch &L, const Image *ref, const Image *srch);

// cp.cu
int cp_get_offset_image(mimc3cu_ctx *ctx, Image *i0, Image *i1, const double *xyuvav, int32_t n, const mimc3cu_params *p,
                        const float *k1x3, const float *k3x1, const float *k3x3, uint32_t seed, int32_t *offset,
                        uint8_t *flag_cp, int32_t *result, int32_t *num_cp_found);

// sat.cu
void image_invalidate(Image *im);
int ensure_image_stats(mimc3cu_ctx *ctx, Image *im);
int ensure_image_sat(mimc3cu_ctx *ctx, Image *im);

// conv2.cu
int launch_conv2(mimc3cu_ctx *ctx, const float *src, int32_t H, int32_t W, const float *kernel, int32_t kh,
                 int32_t kw, float *dst);
int launch_cast_u8(mimc3cu_ctx *ctx, const uint8_t *src, float *dst, size_t count);
int launch_cast_u16(mimc3cu_ctx *ctx, const uint16_t *src, float *dst, size_t count);

// comm.cu (all asynchronous on ctx->stream; local rows [own0, own1) are owned, `halo` rows on either side are the neighbours')
int bandcomm_halo_exchange(mimc3cu_ctx *ctx, void *const *arrays, const int32_t *elem_bytes, int32_t count, int dimx, int rows,
                           int own0, int own1, int halo);
int bandcomm_halo_or_reduce(mimc3cu_ctx *ctx, uint8_t *flags, int dimx, int rows, int own0, int own1, int halo);
int bandcomm_allreduce_sum(mimc3cu_ctx *ctx, int32_t *dev_vals, int32_t count);

// post.cu
int post_cluster(mimc3cu_ctx *ctx, const float *dp, int32_t n, int32_t num_dp, float *mvn, int32_t *ncl);
int post_run(mimc3cu_ctx *ctx, const float *dp, const double *xyuvav_host, const mimc3cu_params *p, float *planes,
             int32_t *stats);
int post_run_band(mimc3cu_ctx *ctx, const float *dp, const double *xyuvav, const mimc3cu_params *p, int32_t own_row0,
                  int32_t own_rows, const mimc3cu_band_comm *comm, float *planes, int32_t *stats);
int post_band_halo(const mimc3cu_params *p);
int post_stage(mimc3cu_ctx *ctx, int32_t which, void *host);
int post_finalize(mimc3cu_ctx *ctx, float *planes, const mimc3cu_params *p, float *du_cp, float *dv_cp);
void post_free(mimc3cu_ctx *ctx);
int post_negate_uv(mimc3cu_ctx *ctx, float *dp, int32_t n);
