// Control-point (CP) stage: the integer image-to-image offset, get_offset_image
// (MIMC_module.c:33-492) + GMA_double_randperm_row (:494-540).
//
// Flow of the reference, kept step for step (the candidate order, the glibc rand() permutation,
// the float segment boundaries and the sequential float sums decide the result):
//   candidates (a-priori speed < thres_spd_cp, <= 50 % null pixels in the 61x61 window of i0)
//   -> random row permutation -> segments -> per segment, 4 image variants x chip half-widths
//   {vec_ocw[1], vec_ocw[2]} x {forward, swapped} = 16 attempts with the shared (2*AW_CRE+1)^2
//   rectangular pivot set on 85x85 tiles -> clusters -> clusters with support >= 0.6 are CPs.
// On the GPU: the tiles of a segment are cut (and filtered) by one kernel launch per variant, the
// 16 attempts run through the general matcher in its explicit-tile mode, clustering reuses the
// postprocess kernel.  Only the candidate bookkeeping and the final sums stay on the host.
//
// Tile-local conv2 (:273-308): the reference filters an 87x87 window per node into ONE output
// buffer reused for every node of the segment, so GMA_float_conv2's stale-border behaviour
//  chains the nodes together: the right border column is shifted again at every
// call and takes part in the next call's global minimum.  That recurrence is a scalar per
// (image, variant) and is evaluated by cp_shift_scan_kernel in node order.
#include <math_constants.h>
#include <stdlib.h>

#include <chrono>

#include "common.cuh"

namespace {

constexpr int kT = 256;

// number of pixels of the (2*ocw+1)^2 window of `img` around each node with value < 0.00001 (:100-103)
__global__ void __launch_bounds__(kT) cp_nullcount_kernel(const float *__restrict__ img, int H, int W, const int2 *__restrict__ uv,
                                                          int n, int ocw, int *__restrict__ count) {
    const int g = blockIdx.x;
    if (g >= n) return;
    const int S = 2 * ocw + 1;
    int c = 0;
    for (int i = threadIdx.x; i < S * S; i += kT) {
        const int r = i / S, q = i - r * S;
        const int y = uv[g].y + r - ocw, x = uv[g].x + q - ocw;
        if (x >= 0 && x < W && y >= 0 && y < H) c += ((double)__ldg(&img[(size_t)y * W + x]) < 0.00001) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&count[g], c);
}

// The same count from the image's summed-area table (sat.cu: the low 24 bits of .y count the pixels < 1e-10).  For an
// exact-class image (non-negative multiples of 1/8) "< 0.00001" and "< 1e-10" both mean "== 0", so the counts agree;
// one thread per node instead of one CTA: the candidate list is most of the grid (16 M nodes at 8-px spacing).
__global__ void __launch_bounds__(kT) cp_nullcount_sat_kernel(const ulonglong2 *__restrict__ sat, int H, int W, const int2 *__restrict__ uv,
                                                              int n, int ocw, int *__restrict__ count) {
    const int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    const int x0 = max(uv[g].x - ocw, 0), y0 = max(uv[g].y - ocw, 0), x1 = min(uv[g].x + ocw + 1, W), y1 = min(uv[g].y + ocw + 1, H);
    int c = 0;
    if (x1 > x0 && y1 > y0) {
        const size_t W1 = (size_t)W + 1;
        const unsigned long long a = __ldg(&sat[y1 * W1 + x1]).y, b = __ldg(&sat[y0 * W1 + x1]).y;
        const unsigned long long cc = __ldg(&sat[y1 * W1 + x0]).y, d = __ldg(&sat[y0 * W1 + x0]).y;
        c = (int)((a - b - cc + d) & 0xffffffull);
    }
    count[g] = c;
}

__device__ __forceinline__ float px_or_zero(const float *img, int H, int W, int y, int x) {
    return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(&img[(size_t)y * W + x]) : 0.0f;
}
// dn_in = (int32_t)(px + 0.5) ? px : NaN   (MIMC_module.c:2548)
__device__ __forceinline__ float null_to_nan(float px) {
    double t = (double)px + 0.5;
    return (t > -1.0 && t < 1.0) ? CUDART_NAN_F : px;
}

// One CTA per (node, image): the T x T tile (T = 2*ocw_chip+1) of the raw image (variant < 0) or
// of the tile-local filtered image before the shift; for filtered variants also the minimum over
// the interior of the (T+2)^2 working window (the part GMA_float_conv2 writes).
__global__ void __launch_bounds__(kT) cp_tiles_kernel(const float *__restrict__ i0, const float *__restrict__ i1, int H, int W,
                                                      const int2 *__restrict__ uv, int nsub, int ocw_chip, int variant,
                                                      const float *__restrict__ kern, int kh, int kw,
                                                      float *__restrict__ tiles, float *__restrict__ mins) {
    __shared__ float red[kT / 32];
    const int g = blockIdx.x, im = blockIdx.y;
    const float *img = im ? i1 : i0;
    const int T = 2 * ocw_chip + 1, Tw = T + 2;
    const int cu = uv[g].x, cv = uv[g].y;
    float *tile = tiles + ((size_t)im * nsub + g) * T * T;
    if (variant < 0) {
        for (int i = threadIdx.x; i < T * T; i += kT) {
            const int r = i / T, c = i - r * T;
            tile[i] = px_or_zero(img, H, W, cv + r - ocw_chip, cu + c - ocw_chip);
        }
        return;
    }
    const int ocwx = kw / 2, ocwy = kh / 2;
    float vmin = 1e+37f;
    // working window coordinates (wr, wc) in [0, Tw); window pixel (wr, wc) = image (cv - ocw_chip - 1 + wr, ...)
    for (int i = threadIdx.x; i < Tw * Tw; i += kT) {
        const int wr = i / Tw, wc = i - wr * Tw;
        if (wr < ocwy || wr >= Tw - ocwy || wc < ocwx || wc >= Tw - ocwx) continue;   // not written by the stencil loop
        float sum = 0.0f;
        for (int a = 0; a < kh; a++)
            for (int b = 0; b < kw; b++) {
                const float px = px_or_zero(img, H, W, cv - ocw_chip - 1 + wr + a - ocwy, cu - ocw_chip - 1 + wc + b - ocwx);
                sum = __fadd_rn(sum, __fmul_rn(null_to_nan(px), kern[a * kw + b]));
            }
        vmin = fminf(vmin, sum);   // fminf drops NaN like the reference's `<` scan
        if (wr >= 1 && wr <= T && wc >= 1 && wc <= T) tile[(wr - 1) * T + (wc - 1)] = sum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = vmin;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kT / 32; w++) vmin = fminf(vmin, red[w]);
        mins[(size_t)im * nsub + g] = vmin;
    }
}

// The per-call global minimum of GMA_float_conv2 on the reused output buffer, in node order:
// dn_min_k = min(interior_k, 0 [never-written cells], r_k [right border column, kernels with kw = 3]),
// r_{k+1} = r_k - (dn_min_k - 1).  One thread per image.
__global__ void cp_shift_scan_kernel(const float *__restrict__ mins, int nsub, int kw, float *__restrict__ shift) {
    const int im = threadIdx.x;
    if (im >= 2) return;
    float r = 0.0f;
    for (int k = 0; k < nsub; k++) {
        float m = 1e+37f;
        const float I = mins[(size_t)im * nsub + k];
        if (I < m) m = I;
        if (0.0f < m) m = 0.0f;
        if (kw == 3 && r < m) m = r;
        const float t = __fsub_rn(m, 1.0f);
        shift[(size_t)im * nsub + k] = t;
        if (kw == 3) r = __fsub_rn(r, t);
    }
}

__global__ void __launch_bounds__(kT) cp_apply_shift_kernel(float *__restrict__ tiles, const float *__restrict__ shift, int nsub,
                                                            int TT) {
    const int g = blockIdx.x, im = blockIdx.y;
    const float t = shift[(size_t)im * nsub + g];
    float *tile = tiles + ((size_t)im * nsub + g) * TT;
    for (int i = threadIdx.x; i < TT; i += kT) {
        const float v = tile[i];
        tile[i] = isnan(v) ? 0.0f : __fsub_rn(v, t);
    }
}

// Device buffers of one call come out of a pool the context keeps (grow-only): cudaMalloc / cudaFree per call cost a
// device-wide synchronisation each and made the stage take 75 ... 240 ms from call to call.  A buffer is a slice of the
// pool; the pool is rewound when a group of buffers goes out of use.
struct CpPool {
    mimc3cu_ctx *ctx;
    size_t used = 0;
    explicit CpPool(mimc3cu_ctx *c) : ctx(c) {}
    cudaError_t reserve(size_t bytes) {     // only while no slice is in use (the pool may move)
        used = 0;
        if (bytes <= ctx->cp_pool_bytes) return cudaSuccess;
        cudaStreamSynchronize(ctx->stream);
        if (ctx->cp_pool) cudaFree(ctx->cp_pool);
        ctx->cp_pool = nullptr; ctx->cp_pool_bytes = 0;
        const size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&ctx->cp_pool, want);
        if (e == cudaSuccess) ctx->cp_pool_bytes = want;
        return e;
    }
    void *take(size_t bytes) {
        void *p = (char *)ctx->cp_pool + used;
        used += (bytes + 255) & ~(size_t)255;
        return used <= ctx->cp_pool_bytes ? p : nullptr;
    }
};
struct DevBuf {
    void *p = nullptr;
    template <typename T> T *as() { return (T *)p; }
    cudaError_t take(CpPool &pool, size_t bytes) { p = pool.take(bytes ? bytes : 1); return p ? cudaSuccess : cudaErrorMemoryAllocation; }
};
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

int cp_get_offset_image(mimc3cu_ctx *ctx, Image *i0, Image *i1, const double *xyuvav, int32_t n, const mimc3cu_params *p,
                        const float *k1x3, const float *k3x1, const float *k3x3, uint32_t seed, int32_t *offset,
                        uint8_t *flag_cp, int32_t *result, int32_t *num_cp_found) {
    const int H = i0->H, W = i0->W;
    cudaStream_t st = ctx->stream;
    CpPool pool(ctx);
    *result = -1;
    if (num_cp_found) *num_cp_found = 0;
    const int ocw2 = p->vec_ocw[2];
    const int ocw_chip = (int)((float)ocw2 + p->AW_CRE + 2.0f);   // :48 (int + float + int, truncated)
    const int T = 2 * ocw_chip + 1;

    // number of control points sought (:57-64)
    int32_t num_cp;
    if ((float)n * p->ratio_cp > (float)p->num_cp_max) num_cp = p->num_cp_max;
    else num_cp = (int32_t)((float)n * p->ratio_cp);

    // wall-clock checkpoints on stderr with MIMC3CU_CP_TIMING=1 (development aid)
    const bool cp_timing = getenv("MIMC3CU_CP_TIMING") != nullptr;
    auto cp_t0 = std::chrono::steady_clock::now();
    auto checkpoint = [&](const char *what) {
        if (!cp_timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[mimc3cu cp] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - cp_t0).count());
        cp_t0 = now;
    };
    // candidates: slow nodes (:71-83) ...
    std::vector<int32_t> cand;
    for (int32_t g = 0; g < n; g++) {
        const float spd_sq = (float)(xyuvav[6 * (size_t)g + 4] * xyuvav[6 * (size_t)g + 4] + xyuvav[6 * (size_t)g + 5] * xyuvav[6 * (size_t)g + 5]);
        if (spd_sq < p->thres_spd_cp * p->thres_spd_cp) cand.push_back(g);
    }
    // ... whose 61x61 window of i0 is at most half null (:86-121; the reference tests i0 twice)
    if (!cand.empty()) {
        std::vector<int2> uv(cand.size());
        for (size_t k = 0; k < cand.size(); k++) {
            uv[k].x = (int32_t)xyuvav[6 * (size_t)cand[k] + 2];
            uv[k].y = (int32_t)xyuvav[6 * (size_t)cand[k] + 3];
        }
        DevBuf d_uv, d_cnt;
        CU_CHECK(ctx, pool.reserve(pad256(sizeof(int2) * uv.size()) + pad256(sizeof(int) * uv.size())));
        CU_CHECK(ctx, d_uv.take(pool, sizeof(int2) * uv.size()));
        CU_CHECK(ctx, d_cnt.take(pool, sizeof(int) * uv.size()));
        CU_CHECK(ctx, cudaMemcpyAsync(d_uv.p, uv.data(), sizeof(int2) * uv.size(), cudaMemcpyHostToDevice, st));
        CU_CHECK(ctx, cudaMemsetAsync(d_cnt.p, 0, sizeof(int) * uv.size(), st));
        // the table is needed by the matcher anyway (raw-pair attempts) and is cached with the image
        // (only for very long candidate lists -- 8-px node spacing -- where the per-candidate pixel loop costs seconds; at
        // C2 size the loop takes 10 ms and the table would have to be built before the matcher needs it)
        const bool use_sat = ctx->matcher != 1 && uv.size() > (size_t)4000000 && !getenv("MIMC3CU_CP_NO_SAT");
        if (use_sat) { if (int rc = ensure_image_sat(ctx, i0)) return rc; }
        if (use_sat && i0->exact_class && i0->sat_valid)
            cp_nullcount_sat_kernel<<<(unsigned)((uv.size() + kT - 1) / kT), kT, 0, st>>>((const ulonglong2 *)i0->sat, H, W, d_uv.as<int2>(),
                                                                                      (int)uv.size(), ocw2, d_cnt.as<int>());
        else
            cp_nullcount_kernel<<<(unsigned)uv.size(), kT, 0, st>>>(i0->d, H, W, d_uv.as<int2>(), (int)uv.size(), ocw2, d_cnt.as<int>());
        ctx->launches++;
        CU_CHECK(ctx, cudaGetLastError());
        std::vector<int> cnt(uv.size());
        CU_CHECK(ctx, cudaMemcpyAsync(cnt.data(), d_cnt.p, sizeof(int) * uv.size(), cudaMemcpyDeviceToHost, st));
        CU_CHECK(ctx, cudaStreamSynchronize(st));
        const int thres_numpx = (2 * ocw2 + 1) * (2 * ocw2 + 1) / 2;
        std::vector<int32_t> keep;
        for (size_t k = 0; k < cand.size(); k++)
            if (!(cnt[k] > thres_numpx)) keep.push_back(cand[k]);
        cand.swap(keep);
    }
    checkpoint("candidates + null counts");
    const int32_t num_cand = (int32_t)cand.size();
    if (num_cand < p->num_cp_min) return 0;   // :130-135 (result stays -1)
    if (num_cp > num_cand) num_cp = (int32_t)((float)num_cand * 0.75);   // :137-141 (float * double literal)
    if (num_cp < 1) num_cp = 1;

    // GMA_double_randperm_row (:494-540) on the candidate list, glibc srand/rand
    {
        std::vector<int32_t> tmp(cand), out(num_cand);
        srand(seed);
        for (int32_t lim = num_cand - 1; lim >= 0; lim--) {
            const int32_t idx = lim != 0 ? (int32_t)(rand() % lim) : 0;
            out[lim] = tmp[idx];
            tmp[idx] = tmp[0];
            tmp[0] = tmp[lim];
        }
        cand.swap(out);
    }

    checkpoint("random permutation");
    // segments (:183-194)
    const int32_t num_segment = (num_cand < p->num_cp_min) ? 1 : num_cand / num_cp;
    std::vector<int32_t> seg(num_segment + 1);
    seg[0] = 0;
    for (int32_t c = 1; c <= num_segment; c++) seg[c] = (int32_t)((float)num_cand * ((float)c / (float)num_segment));

    // shared rectangular pivot set (:165-176): u outer, v inner
    const int R = (int)p->AW_CRE;
    std::vector<int32_t> piv;
    for (int a = -R; a <= R; a++)
        for (int b = -R; b <= R; b++) { piv.push_back(a); piv.push_back(b); }
    const int P = (int)piv.size() / 2;
    float kall[15];
    memcpy(kall, k1x3, 3 * sizeof(float)); memcpy(kall + 3, k3x1, 3 * sizeof(float)); memcpy(kall + 6, k3x3, 9 * sizeof(float));
    const int kh[3] = {1, 3, 3}, kw[3] = {3, 1, 3}, koff[3] = {0, 3, 6};
    int32_t max_sub = 0;
    for (int32_t c = 0; c < num_segment; c++) max_sub = std::max(max_sub, seg[c + 1] - seg[c]);
    const size_t ms = (size_t)max_sub;
    const size_t sizes[9] = {sizeof(int32_t) * piv.size(), sizeof(kall), sizeof(int2) * ms, sizeof(float) * 2 * ms * T * T,
                             sizeof(float) * 2 * ms, sizeof(float) * 2 * ms, sizeof(float) * 16 * ms * 3,
                             sizeof(float) * ms * 16 * 5, sizeof(int32_t) * ms};
    size_t total = 0;
    for (size_t b : sizes) total += pad256(b ? b : 1);
    // the candidate buffers of the step above are no longer in use: the pool is rewound (and may grow) here
    CU_CHECK(ctx, pool.reserve(total));
    DevBuf d_piv, d_kern, d_uv, d_tiles, d_mins, d_shift, d_dp, d_mvn, d_ncl;
    DevBuf *bufs[9] = {&d_piv, &d_kern, &d_uv, &d_tiles, &d_mins, &d_shift, &d_dp, &d_mvn, &d_ncl};
    for (int b = 0; b < 9; b++) CU_CHECK(ctx, bufs[b]->take(pool, sizes[b]));
    CU_CHECK(ctx, cudaMemcpyAsync(d_piv.p, piv.data(), sizeof(int32_t) * piv.size(), cudaMemcpyHostToDevice, st));
    CU_CHECK(ctx, cudaMemcpyAsync(d_kern.p, kall, sizeof(kall), cudaMemcpyHostToDevice, st));

    checkpoint("buffers");
    float sduv[2] = {0.0f, 0.0f};
    int32_t num_cp_current = 0;
    bool ok = false;
    for (int32_t c = 0; c < num_segment; c++) {
        const int32_t nsub = seg[c + 1] - seg[c];
        if (nsub <= 0) continue;
        std::vector<int2> uv(nsub);
        for (int32_t k = 0; k < nsub; k++) {
            const int32_t g = cand[seg[c] + k];
            uv[k].x = (int32_t)xyuvav[6 * (size_t)g + 2];
            uv[k].y = (int32_t)xyuvav[6 * (size_t)g + 3];
        }
        CU_CHECK(ctx, cudaMemcpyAsync(d_uv.p, uv.data(), sizeof(int2) * (size_t)nsub, cudaMemcpyHostToDevice, st));
        float *tiles0 = d_tiles.as<float>(), *tiles1 = tiles0 + (size_t)nsub * T * T;
        for (int variant = -1; variant <= 2; variant++) {
            const int kk = variant < 0 ? 0 : variant;
            cp_tiles_kernel<<<dim3(nsub, 2), kT, 0, st>>>(i0->d, i1->d, H, W, d_uv.as<int2>(), nsub, ocw_chip, variant,
                                                         d_kern.as<float>() + koff[kk], kh[kk], kw[kk], tiles0, d_mins.as<float>());
            ctx->launches++;
            if (variant >= 0) {
                cp_shift_scan_kernel<<<1, 32, 0, st>>>(d_mins.as<float>(), nsub, kw[kk], d_shift.as<float>());
                cp_apply_shift_kernel<<<dim3(nsub, 2), kT, 0, st>>>(tiles0, d_shift.as<float>(), nsub, T * T);
                ctx->launches += 2;
            }
            CU_CHECK(ctx, cudaGetLastError());
            for (int c3 = 1; c3 < 3; c3++) {
                const int ocw = p->vec_ocw[c3];
                const int slot = (c3 - 1) * 8 + (variant + 1) * 2;
                for (int dir = 0; dir < 2; dir++) {
                    MatchLaunch L;
                    L.chips = (dir ? tiles1 : tiles0) + (size_t)(ocw_chip - ocw) * T + (ocw_chip - ocw);
                    L.chip_stride = (int64_t)T * T; L.chip_pitch = T;
                    L.sareas = dir ? tiles0 : tiles1;
                    L.D = T; L.P = P; L.piv = d_piv.as<int32_t>(); L.sign = 1;
                    L.n = nsub; L.ocw = ocw; L.negate = dir ? -1.0f : 1.0f;
                    L.dp = d_dp.as<float>() + (size_t)(slot + dir) * nsub * 3;
                    L.max_cells = (int64_t)(T - 2 * ocw - 1) * (T - 2 * ocw - 1);
                    L.max_sarea = (int64_t)T * T;
                    if (int rc = launch_match(ctx, L)) return rc;
                }
            }
        }
        // clusters of the 16 attempts (:389) and the prominent ones (:394-410)
        if (int rc = post_cluster(ctx, d_dp.as<float>(), nsub, 16, d_mvn.as<float>(), d_ncl.as<int32_t>())) return rc;
        std::vector<float> mvn((size_t)nsub * 16 * 5);
        std::vector<int32_t> ncl(nsub);
        CU_CHECK(ctx, cudaMemcpyAsync(mvn.data(), d_mvn.p, sizeof(float) * mvn.size(), cudaMemcpyDeviceToHost, st));
        CU_CHECK(ctx, cudaMemcpyAsync(ncl.data(), d_ncl.p, sizeof(int32_t) * ncl.size(), cudaMemcpyDeviceToHost, st));
        CU_CHECK(ctx, cudaStreamSynchronize(st));
        for (int32_t k = 0; k < nsub; k++)
            for (int32_t q = 0; q < ncl[k]; q++) {
                const float *row = &mvn[((size_t)k * 16 + q) * 5];
                if ((double)row[4] >= 0.6) {
                    sduv[0] += row[0]; sduv[1] += row[1];
                    flag_cp[cand[seg[c] + k]] = 1;
                    num_cp_current++;
                }
            }
        checkpoint("segment");
        if (num_cp <= num_cp_current) { ok = true; break; }
    }
    if (num_cp_found) *num_cp_found = num_cp_current;
    if (!ok && num_cp_current >= p->num_cp_min) ok = true;   // :437-441
    if (!ok) return 0;
    const float du = sduv[0] / (float)num_cp_current, dv = sduv[1] / (float)num_cp_current;
    offset[0] = du > 0 ? (int32_t)((double)du + 0.5) : (int32_t)((double)du - 0.5);   // :455-473
    offset[1] = dv > 0 ? (int32_t)((double)dv + 0.5) : (int32_t)((double)dv - 0.5);
    *result = 1;
    return 0;
}
