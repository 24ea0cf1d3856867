// Per-image statistics and summed-area tables (SAT) kept in HBM next to every image.
//
// The v2 matcher (match2.cu) needs, for every NCC cell, the window sums  sum(s), sum(fl(s*s))
// and the number of null pixels of the search window -- and the same three numbers for the
// reference chip.  In the reference these are re-accumulated pixel by pixel for every cell
// (MIMC_module.c:719-733: 5 of the 8 flop per pixel).  Here they come from a 2-D SAT with four
// 16-byte loads per cell.  Layout: (H+1) x (W+1) records, record (y,x) = sums over rows < y and
// columns < x of
//      .x  = (u64) fl32(v*v) * 4^f          (exactly the float product the reference adds)
//      .y  = ((u64)(v * 2^f) << 24) | (v < 1e-10)      (value sum and null count share a word)
// with f the number of fractional bits of the image (0 for integer DN, 3 after the 1/8-weight
// Laplacian).  All arithmetic is modulo 2^64: window sums are far below 2^64 (and the null count
// below 2^24), so the four-corner combination is exact even where the table itself wrapped.
//
// An image is "exact-class" when every pixel is finite, non-negative and a multiple of 2^-3
// (everything GMA_float_load_tiff or GMA_float_conv2 on integer data can produce); only then the
// SAT is built.  Anything else is matched by the general FP64 kernel (match.cu).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// stats[0] = max(v) as float bits (valid for non-negative floats), stats[1] = flags:
// bit0 negative or non-finite, bit1 not an integer, bit2 not a multiple of 1/8
__global__ void __launch_bounds__(kThreads) stats_kernel(const float *__restrict__ img, size_t count, unsigned int *stats) {
    unsigned int mx = 0, fl = 0;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (size_t)gridDim.x * kThreads) {
        float v = __ldg(&img[i]);
        if (!(v >= 0.0f) || !(v < 3.0e38f)) { fl |= 1u; continue; }
        mx = max(mx, __float_as_uint(v));
        if (v != floorf(v)) {
            fl |= 2u;
            float w = __fmul_rn(v, 8.0f);
            if (w != floorf(w)) fl |= 4u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&stats[0], mx);
        if (fl) atomicOr(&stats[1], fl);
    }
}

__device__ __forceinline__ ulonglong2 add2(ulonglong2 a, ulonglong2 b) { return make_ulonglong2(a.x + b.x, a.y + b.y); }

// Row pass: one CTA per image row, tiles of 256 pixels, inclusive prefix along x written to
// sat[(y+1)*(W+1) + x+1]; column 0 of the row is zeroed.
__global__ void __launch_bounds__(kThreads) sat_row_kernel(const float *__restrict__ img, int H, int W, float scale, float scale2,
                                                           float min_dn, ulonglong2 *__restrict__ sat) {
    __shared__ ulonglong2 wsum[kThreads / 32];
    __shared__ ulonglong2 carry_s;
    const int y = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *row = img + (size_t)y * W;
    ulonglong2 *out = sat + (size_t)(y + 1) * (W + 1);
    if (threadIdx.x == 0) { carry_s = make_ulonglong2(0, 0); out[0] = make_ulonglong2(0, 0); }
    __syncthreads();
    for (int x0 = 0; x0 < W; x0 += kThreads) {
        const int x = x0 + threadIdx.x;
        ulonglong2 v = make_ulonglong2(0, 0);
        if (x < W) {
            float p = __ldg(&row[x]);
            v.x = __float2ull_rn(__fmul_rn(__fmul_rn(p, p), scale2));
            v.y = (__float2ull_rn(__fmul_rn(p, scale)) << 24) | (p < min_dn ? 1ull : 0ull);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long ax = __shfl_up_sync(0xffffffffu, v.x, o), ay = __shfl_up_sync(0xffffffffu, v.y, o);
            if (lane >= o) { v.x += ax; v.y += ay; }
        }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        ulonglong2 base = carry_s;
        for (int w = 0; w < warp; w++) base = add2(base, wsum[w]);
        v = add2(v, base);
        if (x < W) out[x + 1] = v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) carry_s = v;
        __syncthreads();
    }
}

// Column pass in three steps over chunks of `R` rows: chunk totals, exclusive scan over the
// chunks, running sum within each chunk (in place).  Threads map to columns => coalesced.
__global__ void __launch_bounds__(kThreads) sat_col_total_kernel(const ulonglong2 *__restrict__ sat, int H, int W1, int R,
                                                                 ulonglong2 *__restrict__ tot) {
    const int x = blockIdx.x * kThreads + threadIdx.x, c = blockIdx.y;
    if (x >= W1) return;
    const int y0 = 1 + c * R, y1 = min(H + 1, y0 + R);
    ulonglong2 s = make_ulonglong2(0, 0);
    for (int y = y0; y < y1; y++) s = add2(s, sat[(size_t)y * W1 + x]);
    tot[(size_t)c * W1 + x] = s;
}
__global__ void __launch_bounds__(kThreads) sat_col_scan_kernel(ulonglong2 *__restrict__ tot, int W1, int nchunks) {
    const int x = blockIdx.x * kThreads + threadIdx.x;
    if (x >= W1) return;
    ulonglong2 run = make_ulonglong2(0, 0);
    for (int c = 0; c < nchunks; c++) {
        ulonglong2 t = tot[(size_t)c * W1 + x];
        tot[(size_t)c * W1 + x] = run;
        run = add2(run, t);
    }
}
__global__ void __launch_bounds__(kThreads) sat_col_apply_kernel(ulonglong2 *__restrict__ sat, int H, int W1, int R,
                                                                 const ulonglong2 *__restrict__ tot) {
    const int x = blockIdx.x * kThreads + threadIdx.x, c = blockIdx.y;
    if (x >= W1) return;
    const int y0 = 1 + c * R, y1 = min(H + 1, y0 + R);
    ulonglong2 run = tot[(size_t)c * W1 + x];
    if (c == 0) sat[x] = make_ulonglong2(0, 0);   // row 0 of the table
    for (int y = y0; y < y1; y++) {
        run = add2(run, sat[(size_t)y * W1 + x]);
        sat[(size_t)y * W1 + x] = run;
    }
}

}  // namespace

void image_invalidate(Image *im) { im->stats_valid = false; im->sat_valid = false; }

// Makes im->exact_class / max_value / frac_bits valid (one pass over the image + a 8-byte D2H).
int ensure_image_stats(mimc3cu_ctx *ctx, Image *im) {
    if (im->stats_valid) return 0;
    unsigned int *d = ctx->statbuf;
    CU_CHECK(ctx, cudaMemsetAsync(d, 0, 2 * sizeof(unsigned int), ctx->stream));
    const size_t count = (size_t)im->H * im->W;
    int blocks = (int)std::min<size_t>((count + kThreads - 1) / kThreads, (size_t)ctx->num_sms * 16);
    stats_kernel<<<blocks, kThreads, 0, ctx->stream>>>(im->d, count, d);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    unsigned int h[2];
    CU_CHECK(ctx, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    float mx;
    memcpy(&mx, &h[0], sizeof(float));
    im->max_value = mx;
    im->exact_class = !(h[1] & 1u) && !(h[1] & 4u);
    im->frac_bits = (h[1] & 2u) ? 3 : 0;
    // scaled values must stay exact integers in float and their squares exact in u64
    if ((double)mx * (1 << im->frac_bits) >= 16777216.0) im->exact_class = false;
    im->stats_valid = true;
    return 0;
}

int ensure_image_sat(mimc3cu_ctx *ctx, Image *im) {
    if (int rc = ensure_image_stats(ctx, im)) return rc;
    if (!im->exact_class) return 0;
    if (im->sat_valid) return 0;
    const int H = im->H, W = im->W, W1 = W + 1;
    const size_t elems = (size_t)(H + 1) * W1;
    if (!im->sat) {
        // 16 B per pixel; if HBM cannot hold it the image simply stays on the general FP64 matcher
        if (cudaMalloc(&im->sat, elems * sizeof(ulonglong2)) != cudaSuccess) {
            cudaGetLastError();
            im->sat = nullptr;
            im->exact_class = false;
            return 0;
        }
    }
    const int R = 128, nchunks = (H + R - 1) / R;
    if (int rc = ensure_scratch(ctx, (size_t)nchunks * W1 * sizeof(ulonglong2))) return rc;
    ulonglong2 *tot = (ulonglong2 *)ctx->scratch;
    const float scale = (float)(1 << im->frac_bits), scale2 = scale * scale;
    float min_dn = (float)1e-10;
    if ((double)min_dn < 1e-10) min_dn = nextafterf(min_dn, 1.0f);
    sat_row_kernel<<<H, kThreads, 0, ctx->stream>>>(im->d, H, W, scale, scale2, min_dn, (ulonglong2 *)im->sat);
    CU_CHECK(ctx, cudaGetLastError());
    dim3 grid((W1 + kThreads - 1) / kThreads, nchunks);
    sat_col_total_kernel<<<grid, kThreads, 0, ctx->stream>>>((const ulonglong2 *)im->sat, H, W1, R, tot);
    sat_col_scan_kernel<<<grid.x, kThreads, 0, ctx->stream>>>(tot, W1, nchunks);
    sat_col_apply_kernel<<<grid, kThreads, 0, ctx->stream>>>((ulonglong2 *)im->sat, H, W1, R, tot);
    CU_CHECK(ctx, cudaGetLastError());
    ctx->launches += 4;
    im->sat_valid = true;
    return 0;
}
