// C ABI of libmimc3cu.so (include/mimc3cu.h): context, images, nodes, pivots and the
// host-side orchestration of the multi-match (MIMC_main.c:261-350).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <thread>

#include "common.cuh"

std::string g_mimc3cu_error;

int mimc3cu_fail(mimc3cu_ctx *ctx, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_mimc3cu_error = buf;
    if (ctx) ctx->err = buf;
    return 1;
}

int ensure_scratch(mimc3cu_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return 0;
    // previously issued kernels may still be using the old buffer
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) CU_CHECK(ctx, cudaFree(ctx->scratch));
    ctx->scratch = nullptr; ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4;
    CU_CHECK(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return 0;
}

int upload_sync(mimc3cu_ctx *ctx, void *dev, const void *host, size_t bytes) {
    if (bytes == 0) return 0;
    CU_CHECK(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->upload_stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->upload_stream));
    return 0;
}

static cudaEvent_t take_event(mimc3cu_ctx *ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->event_pool.empty()) { e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

ScopedTimer::ScopedTimer(mimc3cu_ctx *c, int fam) : ctx(c), family(fam) {
    if (!ctx->timing) return;
    if (ctx->timers[fam].size() >= 65536) return;   // nobody reads them: stop recording instead of growing without bound
    cudaEvent_t start = take_event(ctx);
    stop = take_event(ctx);
    cudaEventRecord(start, ctx->stream);
    ctx->timers[family].push_back({start, stop});
}
ScopedTimer::~ScopedTimer() {
    if (stop) cudaEventRecord(stop, ctx->stream);
}

Image *get_image(mimc3cu_ctx *ctx, int32_t h) {
    if (!ctx || h < 0 || h >= (int32_t)ctx->images.size() || !ctx->images[h].used) return nullptr;
    return &ctx->images[h];
}

namespace {
struct PivotStep { float incr_u, incr_v; int32_t count; };

// Everything is computed with the reference's exact float/double mix (see the comments in
// the reference lines cited); -ffp-contract=off semantics are guaranteed by building this
// file with -Xcompiler -ffp-contract=off.
inline PivotStep pivot_step(const double *row, float dt, float mpp, float AW_SF, float AW_CRE, int32_t ocw, int32_t H,
                            int32_t W) {
    float u = 0.0f, v = 0.0f, incr_u, incr_v, norm_incr, theta;
    theta = (float)atan2(row[5], row[4]);                                   // :560
    incr_u = (float)cos((double)theta);
    incr_v = (float)sin((double)theta);
    if (fabs((double)incr_u) > fabs((double)incr_v)) {                      // :563-571 (keeps the quirk)
        incr_u = (float)((double)incr_u / fabs((double)incr_u));
        incr_v = (float)((double)incr_v / fabs((double)incr_u));
    } else {
        incr_u = (float)((double)incr_u / fabs((double)incr_v));
        incr_v = (float)((double)incr_v / fabs((double)incr_v));
    }
    norm_incr = (float)sqrt((double)(incr_u * incr_u + incr_v * incr_v));  // :572
    double length_pivot = sqrt(row[4] * row[4] + row[5] * row[5]) / (double)mpp / 365 * (double)dt * (double)AW_SF +
                          (double)AW_CRE + 1;                               // :573
    int32_t num = 0;
    const float fu = (float)row[2], fv = (float)row[3], focw = (float)ocw;
    while (u + fu - focw > 0 && u + fu + focw < (float)(W - 1) && v + fv - focw > 0 && v + fv + focw < (float)(H - 1) &&
           length_pivot > (double)((double)norm_incr * (double)num)) {      // :576-580
        num++;
        u += incr_u;
        v += incr_v;
    }
    return {incr_u, incr_v, num};
}

template <typename F>
void parallel_for(int32_t n, F f) {
    // MIMC3CU_HOST_THREADS caps the worker threads (one process per GPU shares the host cores with its siblings)
    static const int cap = [] { const char *e = getenv("MIMC3CU_HOST_THREADS"); int v = e ? atoi(e) : 0; return v > 0 ? v : 64; }();
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max(1u, std::min(hw ? hw : 4u, (unsigned)cap));
    if (n < 4096) nt = 1;
    if (nt == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    int32_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        int32_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=] { f(b, e); });
    }
    for (auto &t : th) t.join();
}
}  // namespace

extern "C" {

void mimc3cu_default_params(mimc3cu_params *p) {
    memset(p, 0, sizeof(*p));
    p->vec_ocw[0] = 7; p->vec_ocw[1] = 15; p->vec_ocw[2] = 30; p->vec_ocw[3] = 40;   // MIMC_main.c:134-137
    p->AW_CRE = 10.0f; p->AW_SF = 1.8f;                                                // :153-154
    p->mpp = 15.0f; p->meter_per_spacing = 300.0f;
    p->radius_neighbor_dpf1 = (float)(1000 / 300);                                     // :161 (integer division)
    p->radius_neighbor_ps = 5.0f;                                                      // :162
    p->dt = 16.0f;
    p->num_dp = 32;
    p->num_cp_max = 500; p->num_cp_min = 50; p->ratio_cp = 0.03f; p->thres_spd_cp = 10.0f;            // :165-168
}

int mimc3cu_version(void) { return MIMC3CU_VERSION; }

int mimc3cu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mimc3cu_create(int device, mimc3cu_ctx **out) {
    if (!out) return mimc3cu_fail(nullptr, "mimc3cu_create: null output pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return mimc3cu_fail(nullptr, "mimc3cu_create: no CUDA device available (%s); this library has no CPU fallback",
                            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return mimc3cu_fail(nullptr, "mimc3cu_create: device %d out of range (0..%d)", device, n - 1);
    CU_CHECK(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_CHECK(nullptr, cudaGetDeviceProperties(&prop, device));
    mimc3cu_ctx *ctx = new mimc3cu_ctx();
    ctx->device = device;
    if (prop.major < 10) {
        delete ctx;
        return mimc3cu_fail(nullptr, "mimc3cu_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                            prop.major, prop.minor);
    }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    cudaError_t ce = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc(&ctx->counter, 256);
    if (ce == cudaSuccess) ce = cudaMalloc(&ctx->minbuf, 4096 * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&ctx->statbuf, 64);
    if (ce != cudaSuccess) {
        mimc3cu_destroy(ctx);   // releases whatever was created
        return mimc3cu_fail(nullptr, "mimc3cu_create: %s", cudaGetErrorString(ce));
    }
    if (const char *m = getenv("MIMC3CU_MATCHER")) {
        if (!strcmp(m, "v1") || !strcmp(m, "general")) ctx->matcher = 1;
        else if (!strcmp(m, "v2")) ctx->matcher = 2;
    }
    *out = ctx;
    return 0;
}

void mimc3cu_destroy(mimc3cu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    mimc3cu_comm_destroy(ctx);
    post_free(ctx);
    for (auto &tv : ctx->timers) for (auto &pr : tv) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    for (auto &im : ctx->images) if (im.used) { if (im.d) cudaFree(im.d); if (im.sat) cudaFree(im.sat); }
    for (auto &p : ctx->pivots) {
        if (p.off) cudaFree(p.off);
        if (p.piv) cudaFree(p.piv);
        for (auto &b : p.bins) if (b.lists) cudaFree(b.lists);
        if (p.last_use) cudaEventDestroy(p.last_use);
    }
    if (ctx->cp_pool) cudaFree(ctx->cp_pool);
    if (ctx->statbuf) cudaFree(ctx->statbuf);
    if (ctx->overflow_list) cudaFree(ctx->overflow_list);
    if (ctx->node_uv) cudaFree(ctx->node_uv);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->counter) cudaFree(ctx->counter);
    if (ctx->minbuf) cudaFree(ctx->minbuf);
    if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *mimc3cu_last_error(const mimc3cu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_mimc3cu_error.c_str(); }
void *mimc3cu_stream(mimc3cu_ctx *ctx) { return (void *)ctx->stream; }
int mimc3cu_sync(mimc3cu_ctx *ctx) { CU_CHECK(ctx, cudaSetDevice(ctx->device)); CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); return 0; }
int64_t mimc3cu_launch_count(const mimc3cu_ctx *ctx) { return ctx->launches; }

/* ---- images ------------------------------------------------------------------------- */
int mimc3cu_image_create(mimc3cu_ctx *ctx, int32_t H, int32_t W, int32_t *handle) {
    if (H <= 0 || W <= 0) return mimc3cu_fail(ctx, "image_create: bad size %dx%d", H, W);
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    Image im;
    im.H = H; im.W = W; im.used = true;
    CU_CHECK(ctx, cudaMalloc(&im.d, (size_t)H * W * sizeof(float)));
    CU_CHECK(ctx, cudaMemsetAsync(im.d, 0, (size_t)H * W * sizeof(float), ctx->stream));
    for (size_t i = 0; i < ctx->images.size(); i++)
        if (!ctx->images[i].used) { ctx->images[i] = im; *handle = (int32_t)i; return 0; }
    ctx->images.push_back(im);
    *handle = (int32_t)ctx->images.size() - 1;
    return 0;
}

int mimc3cu_image_destroy(mimc3cu_ctx *ctx, int32_t handle) {
    cudaSetDevice(ctx->device);
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_destroy: bad handle %d", handle);
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    CU_CHECK(ctx, cudaFree(im->d));
    if (im->sat) CU_CHECK(ctx, cudaFree(im->sat));
    *im = Image();
    return 0;
}

int mimc3cu_image_upload(mimc3cu_ctx *ctx, int32_t handle, const float *host) {
    cudaSetDevice(ctx->device);
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_upload: bad handle %d", handle);
    image_invalidate(im);
    CU_CHECK(ctx, cudaMemcpyAsync(im->d, host, (size_t)im->H * im->W * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int upload_int(mimc3cu_ctx *ctx, int32_t handle, const void *host, int bytes_per_px) {
    cudaSetDevice(ctx->device);
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_upload: bad handle %d", handle);
    size_t count = (size_t)im->H * im->W;
    image_invalidate(im);
    if (int rc = ensure_scratch(ctx, count * bytes_per_px)) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->scratch, host, count * bytes_per_px, cudaMemcpyHostToDevice, ctx->stream));
    int rc = bytes_per_px == 1 ? launch_cast_u8(ctx, (const uint8_t *)ctx->scratch, im->d, count)
                               : launch_cast_u16(ctx, (const uint16_t *)ctx->scratch, im->d, count);
    if (rc) return rc;
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
int mimc3cu_image_upload_u8(mimc3cu_ctx *ctx, int32_t handle, const uint8_t *host) { return upload_int(ctx, handle, host, 1); }
int mimc3cu_image_upload_u16(mimc3cu_ctx *ctx, int32_t handle, const uint16_t *host) { return upload_int(ctx, handle, host, 2); }

int mimc3cu_image_copy_from_device(mimc3cu_ctx *ctx, int32_t handle, const float *dev) {
    cudaSetDevice(ctx->device);
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_copy_from_device: bad handle %d", handle);
    image_invalidate(im);
    CU_CHECK(ctx, cudaMemcpyAsync(im->d, dev, (size_t)im->H * im->W * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mimc3cu_image_download(mimc3cu_ctx *ctx, int32_t handle, float *host) {
    cudaSetDevice(ctx->device);
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_download: bad handle %d", handle);
    CU_CHECK(ctx, cudaMemcpyAsync(host, im->d, (size_t)im->H * im->W * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mimc3cu_image_fill_zero(mimc3cu_ctx *ctx, int32_t handle) {
    cudaSetDevice(ctx->device);
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_fill_zero: bad handle %d", handle);
    image_invalidate(im);
    CU_CHECK(ctx, cudaMemsetAsync(im->d, 0, (size_t)im->H * im->W * sizeof(float), ctx->stream));
    return 0;
}

const float *mimc3cu_image_ptr(mimc3cu_ctx *ctx, int32_t handle) {
    Image *im = get_image(ctx, handle);
    return im ? im->d : nullptr;
}

int mimc3cu_image_invalidate(mimc3cu_ctx *ctx, int32_t handle) {
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_invalidate: bad handle %d", handle);
    image_invalidate(im);
    return 0;
}

int mimc3cu_conv2(mimc3cu_ctx *ctx, int32_t src, const float *kernel, int32_t kh, int32_t kw, int32_t dst) {
    cudaSetDevice(ctx->device);
    Image *s = get_image(ctx, src), *d = get_image(ctx, dst);
    if (!s || !d) return mimc3cu_fail(ctx, "conv2: bad image handle");
    if (s == d) return mimc3cu_fail(ctx, "conv2: src and dst must differ");
    if (s->H != d->H || s->W != d->W) return mimc3cu_fail(ctx, "conv2: size mismatch");
    ScopedTimer tm(ctx, 1);
    image_invalidate(d);
    return launch_conv2(ctx, s->d, s->H, s->W, kernel, kh, kw, d->d);
}

/* ---- pivots (host; bit-identical restatement of get_uv_pivot, MIMC_module.c:543-602) ---- */

int64_t mimc3cu_get_uv_pivot(const double *xyuvav, int32_t n, float dt, float mpp, float AW_SF, float AW_CRE, int32_t ocw,
                             int32_t H, int32_t W, int32_t *off, int32_t *piv) {
    if (!xyuvav || !off || n < 0) { mimc3cu_fail(nullptr, "get_uv_pivot: bad arguments"); return -1; }
    // The two-call protocol (sizes first, then the pivots) would evaluate the trigonometry three
    // times per node; the steps of the sizing call are kept for the immediately following fill call.
    struct Cache {
        const double *xy = nullptr; int32_t n = 0, ocw = 0, H = 0, W = 0; float dt = 0, mpp = 0, sf = 0, cre = 0;
        double first_row[6] = {0, 0, 0, 0, 0, 0}, mid_row[6] = {0, 0, 0, 0, 0, 0}, last_row[6] = {0, 0, 0, 0, 0, 0};
        std::vector<PivotStep> steps;
    };
    static thread_local Cache cache;
    // the fill call must follow the sizing call for the same array: first, middle and last row are compared as well
    // (an array edited in place between the two calls falls back to recomputing the steps)
    auto rows_match = [&] {
        return memcmp(cache.first_row, xyuvav, sizeof(cache.first_row)) == 0 &&
               memcmp(cache.mid_row, xyuvav + 6 * (size_t)(n / 2), sizeof(cache.mid_row)) == 0 &&
               memcmp(cache.last_row, xyuvav + 6 * (size_t)(n - 1), sizeof(cache.last_row)) == 0;
    };
    const bool hit = piv && cache.xy == xyuvav && cache.n == n && cache.ocw == ocw && cache.H == H && cache.W == W &&
                     cache.dt == dt && cache.mpp == mpp && cache.sf == AW_SF && cache.cre == AW_CRE && n > 0 &&
                     cache.steps.size() == (size_t)n && rows_match();
    if (!hit) {
        cache.steps.resize((size_t)n);
        PivotStep *out = cache.steps.data();   // NOT `cache` inside the workers: it is thread_local
        parallel_for(n, [&, out](int32_t b, int32_t e) {
            for (int32_t g = b; g < e; g++) out[g] = pivot_step(xyuvav + 6 * (size_t)g, dt, mpp, AW_SF, AW_CRE, ocw, H, W);
        });
        cache.xy = xyuvav; cache.n = n; cache.ocw = ocw; cache.H = H; cache.W = W;
        cache.dt = dt; cache.mpp = mpp; cache.sf = AW_SF; cache.cre = AW_CRE;
        if (n > 0) {
            memcpy(cache.first_row, xyuvav, sizeof(cache.first_row));
            memcpy(cache.mid_row, xyuvav + 6 * (size_t)(n / 2), sizeof(cache.mid_row));
            memcpy(cache.last_row, xyuvav + 6 * (size_t)(n - 1), sizeof(cache.last_row));
        }
    }
    const PivotStep *steps = cache.steps.data();
    int64_t tot = 0;
    for (int32_t g = 0; g < n; g++) { off[g] = (int32_t)tot; tot += steps[g].count; }
    if (tot > 0x7fffffffLL) { mimc3cu_fail(nullptr, "get_uv_pivot: more than 2^31 pivots"); return -1; }
    off[n] = (int32_t)tot;
    if (!piv) return tot;
    parallel_for(n, [&, steps](int32_t b, int32_t e) {
        for (int32_t g = b; g < e; g++) {
            const PivotStep s = steps[g];
            int32_t *dst = piv + 2 * (size_t)off[g];
            float u = 0.0f, v = 0.0f;
            if (s.count > 0) { dst[0] = 0; dst[1] = 0; }
            for (int32_t k = 1; k < s.count; k++) {
                u += s.incr_u; v += s.incr_v;
                dst[2 * k] = (int32_t)((double)u + 0.5);        // :596
                dst[2 * k + 1] = -(int32_t)((double)v + 0.5);   // :597
            }
        }
    });
    cache.xy = nullptr;   // one-shot: the host array may be rewritten before the next call
    return tot;
}

/* ---- nodes / pivot upload ------------------------------------------------------------- */
int mimc3cu_set_nodes(mimc3cu_ctx *ctx, const double *xyuvav, int32_t n) {
    if (n <= 0) return mimc3cu_fail(ctx, "set_nodes: n must be positive");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if ((size_t)n > ctx->node_cap) {   // device arrays are kept and reused: cudaFree/cudaMalloc churn costs more than the copies
        if (ctx->node_uv) { CU_CHECK(ctx, cudaFree(ctx->node_uv)); ctx->node_uv = nullptr; }
        CU_CHECK(ctx, cudaMalloc(&ctx->node_uv, sizeof(int2) * (size_t)n));
        ctx->node_cap = (size_t)n;
    }
    std::vector<int2> uv((size_t)n);
    parallel_for(n, [&](int32_t b, int32_t e) {
        for (int32_t g = b; g < e; g++) {
            uv[g].x = (int32_t)xyuvav[6 * (size_t)g + 2];   // truncation, MIMC_module.c:822-823
            uv[g].y = (int32_t)xyuvav[6 * (size_t)g + 3];
        }
    });
    if (int rc = upload_sync(ctx, ctx->node_uv, uv.data(), sizeof(int2) * (size_t)n)) return rc;
    ctx->n = n;
    return 0;
}

int mimc3cu_set_pivots(mimc3cu_ctx *ctx, int32_t slot, const int32_t *off, const int32_t *piv, int32_t n) {
    if (slot < 0 || slot >= MIMC3CU_MAX_PIVOT_SLOTS) return mimc3cu_fail(ctx, "set_pivots: bad slot %d", slot);
    if (n <= 0) return mimc3cu_fail(ctx, "set_pivots: n must be positive");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    PivotSet &ps = ctx->pivots[slot];
    // Only launches that read THIS slot have to finish: the host can prepare the pivots of the next chip
    // size while the GPU is still matching with the previous one (mimc3_b200/pipeline.py does).
    if (ps.last_use) CU_CHECK(ctx, cudaEventSynchronize(ps.last_use));
    for (auto &b : ps.bins) b.ocw = -1;   // node lists are rebuilt lazily (their device buffers are reused)
    ps.n = n; ps.total = off[n];
    ps.last_u.assign((size_t)n, 0); ps.last_v.assign((size_t)n, 0);
    parallel_for(n, [&](int32_t b, int32_t e) {
        for (int32_t g = b; g < e; g++) {
            const int32_t P = off[g + 1] - off[g];
            if (P <= 0) continue;
            ps.last_u[g] = abs(piv[2 * ((size_t)off[g] + P - 1)]);
            ps.last_v[g] = abs(piv[2 * ((size_t)off[g] + P - 1) + 1]);
        }
    });
    ps.max_abs_u = 0; ps.max_abs_v = 0; ps.max_cells = 16; ps.max_sarea_extra = 0;
    for (int32_t g = 0; g < n; g++) {
        const int32_t lu = ps.last_u[g], lv = ps.last_v[g];
        ps.max_abs_u = std::max(ps.max_abs_u, lu); ps.max_abs_v = std::max(ps.max_abs_v, lv);
        ps.max_cells = std::max<int64_t>(ps.max_cells, (int64_t)(2 * lu + 4) * (2 * lv + 4));
    }
    if ((size_t)n + 1 > ps.off_cap) {
        if (ps.off) { CU_CHECK(ctx, cudaFree(ps.off)); ps.off = nullptr; }
        CU_CHECK(ctx, cudaMalloc(&ps.off, sizeof(int32_t) * ((size_t)n + 1)));
        ps.off_cap = (size_t)n + 1;
    }
    const size_t tot = (size_t)std::max<int64_t>(ps.total, 1);
    if (tot > ps.piv_cap) {
        if (ps.piv) { CU_CHECK(ctx, cudaFree(ps.piv)); ps.piv = nullptr; }
        CU_CHECK(ctx, cudaMalloc(&ps.piv, sizeof(int32_t) * 2 * (tot + tot / 8)));
        ps.piv_cap = tot + tot / 8;
    }
    // both copies are queued before the wait (the caller's arrays are pageable as a rule: staged chunk by chunk)
    CU_CHECK(ctx, cudaMemcpyAsync(ps.off, off, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->upload_stream));
    if (int rc = upload_sync(ctx, ps.piv, piv, sizeof(int32_t) * 2 * (size_t)std::max<int64_t>(ps.total, 0))) return rc;
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->upload_stream));
    return 0;
}

/* ---- matcher ---------------------------------------------------------------------------- */
int mimc3cu_match_async(mimc3cu_ctx *ctx, int32_t ref_img, int32_t search_img, const int32_t *offset, int32_t pivot_slot,
                        int32_t sign, int32_t ocw, int32_t negate_duv, float *dp_dev, int32_t *peak_dev,
                        int32_t *ncell_dev) {
    Image *r = get_image(ctx, ref_img), *s = get_image(ctx, search_img);
    if (!r || !s) return mimc3cu_fail(ctx, "match: bad image handle");
    if (r->H != s->H || r->W != s->W) return mimc3cu_fail(ctx, "match: the two images must have the same size");
    if (pivot_slot < 0 || pivot_slot >= MIMC3CU_MAX_PIVOT_SLOTS || !ctx->pivots[pivot_slot].off)
        return mimc3cu_fail(ctx, "match: pivot slot %d is empty", pivot_slot);
    PivotSet &ps = ctx->pivots[pivot_slot];
    if (!ctx->node_uv || ps.n != ctx->n) return mimc3cu_fail(ctx, "match: nodes not set or pivot/node count mismatch");
    if (!dp_dev) return mimc3cu_fail(ctx, "match: dp output is required");
    if (sign != 1 && sign != -1) return mimc3cu_fail(ctx, "match: sign must be +1 or -1");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    MatchLaunch L;
    L.ref = r->d; L.srch = s->d; L.H = r->H; L.W = r->W;
    L.node_uv = ctx->node_uv; L.off_u = offset ? offset[0] : 0; L.off_v = offset ? offset[1] : 0;
    L.csr_off = ps.off; L.piv = ps.piv; L.sign = sign;
    L.n = ctx->n; L.ocw = ocw; L.negate = negate_duv ? -1.0f : 1.0f;
    L.dp = dp_dev; L.peak = peak_dev; L.ncell = ncell_dev;
    L.max_cells = ps.max_cells;
    L.max_sarea = (int64_t)(2 * (ps.max_abs_u + ocw + 2) + 1) * (2 * (ps.max_abs_v + ocw + 2) + 1);
    bool v2 = false;
    if (ctx->matcher != 1) {
        {   // image statistics + summed-area tables (cached until the image changes): preprocessing family
            ScopedTimer tp(ctx, 1);
            if (int rc = ensure_image_sat(ctx, r)) return rc;
            if (int rc = ensure_image_sat(ctx, s)) return rc;
        }
        v2 = match2_supported(L, r, s);
        if (!v2 && ctx->matcher == 2)
            return mimc3cu_fail(ctx, "match: the exact-FP32 matcher was required (MIMC3CU_MATCHER=v2) but this image pair / "
                                     "chip size is outside its class");
    }
    ctx->last_matcher = v2 ? 2 : 1;
    int rc;
    {
        ScopedTimer tm(ctx, 0);
        rc = v2 ? launch_match2(ctx, L, r, s, &ps) : launch_match(ctx, L);
    }
    // recorded even after a failed launch sequence: some of its kernels may already be reading the slot
    if (!ps.last_use && cudaEventCreateWithFlags(&ps.last_use, cudaEventDisableTiming) != cudaSuccess) ps.last_use = nullptr;
    if (ps.last_use) cudaEventRecord(ps.last_use, ctx->stream);
    else if (!rc) CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // no event: fall back to a full wait
    return rc;
}

int mimc3cu_set_matcher(mimc3cu_ctx *ctx, int32_t mode) {
    if (mode < 0 || mode > 2) return mimc3cu_fail(ctx, "set_matcher: mode must be 0 (auto), 1 (general FP64) or 2 (require exact-FP32)");
    ctx->matcher = mode;
    return 0;
}
int mimc3cu_last_matcher(const mimc3cu_ctx *ctx) { return ctx->last_matcher; }

int mimc3cu_image_class(mimc3cu_ctx *ctx, int32_t handle, int32_t *exact_class, int32_t *frac_bits, float *max_value) {
    Image *im = get_image(ctx, handle);
    if (!im) return mimc3cu_fail(ctx, "image_class: bad handle %d", handle);
    if (int rc = ensure_image_stats(ctx, im)) return rc;
    if (exact_class) *exact_class = im->exact_class ? 1 : 0;
    if (frac_bits) *frac_bits = im->frac_bits;
    if (max_value) *max_value = im->max_value;
    return 0;
}

int mimc3cu_match(mimc3cu_ctx *ctx, int32_t ref_img, int32_t search_img, const int32_t *offset, int32_t pivot_slot,
                  int32_t sign, int32_t ocw, int32_t negate_duv, float *dp_host, int32_t *peak_host, int32_t *ncell_host) {
    if (!dp_host) return mimc3cu_fail(ctx, "match: dp output is required");
    const size_t n = (size_t)ctx->n;
    float *dp = nullptr; int32_t *peak = nullptr, *ncell = nullptr;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaMalloc(&dp, n * 3 * sizeof(float)));
    if (peak_host) CU_CHECK(ctx, cudaMalloc(&peak, n * 2 * sizeof(int32_t)));
    if (ncell_host) CU_CHECK(ctx, cudaMalloc(&ncell, n * sizeof(int32_t)));
    int rc = mimc3cu_match_async(ctx, ref_img, search_img, offset, pivot_slot, sign, ocw, negate_duv, dp, peak, ncell);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(dp_host, dp, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && peak) e = cudaMemcpyAsync(peak_host, peak, n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && ncell) e = cudaMemcpyAsync(ncell_host, ncell, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = mimc3cu_fail(ctx, "match: %s", cudaGetErrorString(e));
    }
    cudaFree(dp); if (peak) cudaFree(peak); if (ncell) cudaFree(ncell);
    return rc;
}

int mimc3cu_find_ncc_peak_batch(mimc3cu_ctx *ctx, const float *refchips, int32_t S, const float *sareas, int32_t D,
                                int32_t count, const int32_t *piv, int32_t P, float *uvncc_host, int32_t *peak_host,
                                int32_t *ncell_host) {
    if (count <= 0) return 0;
    if (!(S & 1) || S < 3 || D < S + 4) return mimc3cu_fail(ctx, "find_ncc_peak_batch: bad chip/search sizes %d/%d", S, D);
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    const size_t cs = (size_t)count * S * S, ss = (size_t)count * D * D;
    float *d_chips = nullptr, *d_sa = nullptr, *d_dp = nullptr;
    int32_t *d_piv = nullptr, *d_peak = nullptr, *d_ncell = nullptr;
    CU_CHECK(ctx, cudaMalloc(&d_chips, cs * sizeof(float)));
    CU_CHECK(ctx, cudaMalloc(&d_sa, ss * sizeof(float)));
    CU_CHECK(ctx, cudaMalloc(&d_dp, (size_t)count * 3 * sizeof(float)));
    CU_CHECK(ctx, cudaMalloc(&d_piv, (size_t)P * 2 * sizeof(int32_t)));
    CU_CHECK(ctx, cudaMalloc(&d_peak, (size_t)count * 2 * sizeof(int32_t)));
    CU_CHECK(ctx, cudaMalloc(&d_ncell, (size_t)count * sizeof(int32_t)));
    CU_CHECK(ctx, cudaMemcpyAsync(d_chips, refchips, cs * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(d_sa, sareas, ss * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(d_piv, piv, (size_t)P * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    MatchLaunch L;
    L.chips = d_chips; L.sareas = d_sa; L.D = D; L.P = P; L.piv = d_piv; L.sign = 1;
    L.n = count; L.ocw = S / 2; L.dp = d_dp; L.peak = d_peak; L.ncell = d_ncell;
    L.max_cells = (int64_t)(D - 2 * (S / 2) - 1) * (D - 2 * (S / 2) - 1);
    L.max_sarea = (int64_t)D * D;
    int rc = launch_match(ctx, L);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(uvncc_host, d_dp, (size_t)count * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && peak_host) e = cudaMemcpyAsync(peak_host, d_peak, (size_t)count * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && ncell_host) e = cudaMemcpyAsync(ncell_host, d_ncell, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = mimc3cu_fail(ctx, "find_ncc_peak_batch: %s", cudaGetErrorString(e));
    }
    cudaFree(d_chips); cudaFree(d_sa); cudaFree(d_dp); cudaFree(d_piv); cudaFree(d_peak); cudaFree(d_ncell);
    return rc;
}

// The three filter kernels of MIMC_main.c:175-196.
static const float kFilter[3][9] = {
    {-1.f, 0.f, 1.f},
    {-1.f, 0.f, 1.f},
    {-0.125f, -0.125f, -0.125f, -0.125f, 1.f, -0.125f, -0.125f, -0.125f, -0.125f},
};
static const int kFilterH[3] = {1, 3, 3}, kFilterW[3] = {3, 1, 3};

int mimc3cu_multimatch_async(mimc3cu_ctx *ctx, int32_t i0, int32_t i1, int32_t i0c, int32_t i1c, const int32_t *offset,
                             const mimc3cu_params *p, float *dp_dev, int32_t *ncell_dev) {
    return mimc3cu_multimatch_diag_async(ctx, i0, i1, i0c, i1c, offset, p, dp_dev, ncell_dev, nullptr);
}

int mimc3cu_multimatch_diag_async(mimc3cu_ctx *ctx, int32_t i0, int32_t i1, int32_t i0c, int32_t i1c, const int32_t *offset,
                                  const mimc3cu_params *p, float *dp_dev, int32_t *ncell_dev, int32_t *peak_dev) {
    if (!p || !dp_dev) return mimc3cu_fail(ctx, "multimatch: params and dp output are required");
    const size_t n = (size_t)ctx->n;
    int32_t off_rev[2] = {offset ? -offset[0] : 0, offset ? -offset[1] : 0};
    int32_t off_fwd[2] = {offset ? offset[0] : 0, offset ? offset[1] : 0};
    // main allocates i0c/i1c once per run (MIMC_main.c:304-305); under the zero-initialised-
    // allocation semantics (H1) they start as zeros, and the three conv2 calls then reuse them.
    for (int32_t h : {i0c, i1c}) {
        Image *im = get_image(ctx, h);
        if (!im) return mimc3cu_fail(ctx, "multimatch: bad scratch image handle");
        CU_CHECK(ctx, cudaMemsetAsync(im->d, 0, (size_t)im->H * im->W * sizeof(float), ctx->stream));
    }
    for (int variant = 0; variant < 4; variant++) {   // raw, then the three filtered pairs (:261-350)
        int32_t a = i0, b = i1;
        if (variant > 0) {
            const int k = variant - 1;
            if (int rc = mimc3cu_conv2(ctx, i0, kFilter[k], kFilterH[k], kFilterW[k], i0c)) return rc;
            if (int rc = mimc3cu_conv2(ctx, i1, kFilter[k], kFilterH[k], kFilterW[k], i1c)) return rc;
            a = i0c; b = i1c;
        }
        for (int c = 0; c < 4; c++) {
            const int idx = variant * 8 + c * 2;   // dp[cnt*2] / dp[cnt*8+cntc*2+8]
            if (int rc = mimc3cu_match_async(ctx, a, b, off_fwd, c, +1, p->vec_ocw[c], 0, dp_dev + (size_t)idx * n * 3,
                                             peak_dev ? peak_dev + (size_t)idx * n * 2 : nullptr,
                                             ncell_dev ? ncell_dev + (size_t)idx * n : nullptr)) return rc;
            if (int rc = mimc3cu_match_async(ctx, b, a, off_rev, c, -1, p->vec_ocw[c], 1, dp_dev + (size_t)(idx + 1) * n * 3,
                                             peak_dev ? peak_dev + (size_t)(idx + 1) * n * 2 : nullptr,
                                             ncell_dev ? ncell_dev + (size_t)(idx + 1) * n : nullptr)) return rc;
        }
    }
    return 0;
}

/* ---- control points (cp.cu) ---------------------------------------------------------------- */
int mimc3cu_get_offset_image(mimc3cu_ctx *ctx, int32_t i0, int32_t i1, const double *xyuvav, int32_t n, const mimc3cu_params *p,
                             const float *k1x3, const float *k3x1, const float *k3x3, uint32_t seed, int32_t *offset,
                             uint8_t *flag_cp, int32_t *result, int32_t *num_cp_found) {
    Image *a = get_image(ctx, i0), *b = get_image(ctx, i1);
    if (!a || !b) return mimc3cu_fail(ctx, "get_offset_image: bad image handle");
    if (a->H != b->H || a->W != b->W) return mimc3cu_fail(ctx, "get_offset_image: the two images must have the same size");
    if (!xyuvav || !p || !offset || !flag_cp || !result || n <= 0) return mimc3cu_fail(ctx, "get_offset_image: bad arguments");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    return cp_get_offset_image(ctx, a, b, xyuvav, n, p, k1x3 ? k1x3 : kFilter[0], k3x1 ? k3x1 : kFilter[1], k3x3 ? k3x3 : kFilter[2],
                               seed, offset, flag_cp, result, num_cp_found);
}

/* ---- postprocess (post.cu) ---------------------------------------------------------------- */
int mimc3cu_cluster_async(mimc3cu_ctx *ctx, const float *dp_dev, int32_t n, int32_t num_dp, float *mvn_dev, int32_t *ncl_dev) {
    return post_cluster(ctx, dp_dev, n, num_dp, mvn_dev, ncl_dev);
}
int mimc3cu_postprocess(mimc3cu_ctx *ctx, const float *dp_dev, const double *xyuvav, const mimc3cu_params *p,
                        float *planes_dev, int32_t *stats) {
    ScopedTimer tm(ctx, 2);
    return post_run(ctx, dp_dev, xyuvav, p, planes_dev, stats);
}
int mimc3cu_band_halo(const mimc3cu_params *p) { return p ? post_band_halo(p) : 0; }
int mimc3cu_postprocess_band(mimc3cu_ctx *ctx, const float *dp_dev, const double *xyuvav, const mimc3cu_params *p, int32_t own_row0,
                             int32_t own_rows, const mimc3cu_band_comm *comm, float *planes_dev, int32_t *stats) {
    ScopedTimer tm(ctx, 2);
    return post_run_band(ctx, dp_dev, xyuvav, p, own_row0, own_rows, comm, planes_dev, stats);
}
int mimc3cu_postprocess_stage(mimc3cu_ctx *ctx, int32_t which, void *host) { return post_stage(ctx, which, host); }
int mimc3cu_dp_negate_uv_async(mimc3cu_ctx *ctx, float *dp_dev, int32_t n) { return post_negate_uv(ctx, dp_dev, n); }
int mimc3cu_finalize(mimc3cu_ctx *ctx, float *planes_dev, const mimc3cu_params *p, float *du_cp, float *dv_cp) {
    return post_finalize(ctx, planes_dev, p, du_cp, dv_cp);
}

/* ---- timing ---------------------------------------------------------------------------------- */
int mimc3cu_timing_enable(mimc3cu_ctx *ctx, int on) { ctx->timing = on != 0; return 0; }

int mimc3cu_timing_read(mimc3cu_ctx *ctx, double *ms, int64_t *counts) {
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int f = 0; f < 3; f++) {
        double tot = 0.0;
        for (auto &pr : ctx->timers[f]) {
            float t = 0.f;
            CU_CHECK(ctx, cudaEventElapsedTime(&t, pr.first, pr.second));
            tot += t;
            ctx->event_pool.push_back(pr.first); ctx->event_pool.push_back(pr.second);
        }
        if (ms) ms[f] = tot;
        if (counts) counts[f] = (int64_t)ctx->timers[f].size();
        ctx->timers[f].clear();
    }
    return 0;
}

/* ---- memory helpers ------------------------------------------------------------------------ */
int mimc3cu_malloc(mimc3cu_ctx *ctx, size_t bytes, void **dev) {
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaMalloc(dev, bytes ? bytes : 1));
    return 0;
}
int mimc3cu_free(mimc3cu_ctx *ctx, void *dev) {
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    CU_CHECK(ctx, cudaFree(dev));
    return 0;
}
int mimc3cu_memcpy_d2h(mimc3cu_ctx *ctx, void *host, const void *dev, size_t bytes) {
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
int mimc3cu_memcpy_h2d(mimc3cu_ctx *ctx, void *dev, const void *host, size_t bytes) {
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
