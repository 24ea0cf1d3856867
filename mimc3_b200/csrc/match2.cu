// DLC-constrained, null-excluding NCC matcher, v2: exact FP32 arithmetic for sm_100a.
//
// Same contract as match.cu (matching_ncc_dlc_2 + extract_refchip + extract_sarea +
// investigate_valid_grid + find_ncc_peak, MIMC_module.c:605-890) and the same bit-exact
// results, for "exact-class" image pairs (sat.cu: non-negative multiples of 2^-3, which is
// everything GMA_float_load_tiff / GMA_float_conv2 produce from integer DN).  What changed is
// where the arithmetic runs:
//
//  * v1 accumulates five sums per pixel in FP64 and is bound by the F2F.F64.F32 conversion
//    (16 lanes/clk/SM on B200, profiles/r01_ubench_sm100a.txt).
//  * v2 observes that for a cell whose window and chip contain no null pixel the joint mask of
//    MIMC_module.c:723 is all-ones, so n = S^2, sum(r), sum(r*r) are per-node constants and
//    sum(s), sum(s*s) are window sums: all five come from the per-image summed-area tables
//    (4 x 16-byte loads).  Only sum(fl(r*s)) needs the S^2 loop, and it is accumulated EXACTLY in
//    FP32 with an error-free transformation (Fast2Sum against a running accumulator biased by a
//    power of two A0 >= 16 * max product):
//        p = r*s;  t = acc + p;  z = t - acc;  e = p - z;  acc = t;  lo += e
//    i.e. 1 FMUL + 4 FADD per pixel on the FP32 pipes, no conversion -- issued as packed pairs
//    (FFMA2 / FADD2, two pixels per instruction; the even/odd accumulators are the two lanes).
//    acc stays in one binade [A0, 2*A0], so float_as_uint(acc) - float_as_uint(A0) is the partial
//    sum in units of ulp(A0); lo is biased by 1.5*2^23 units the same way.  Both are then reduced
//    as integers (REDUX.SUM), and the cell is normalised in FP64 with the reference's expression
//    (:734).  Pairs whose products are all exact in FP32 (u8 and filtered-u8 data) use the FFMA
//    form of the same transformation, one instruction less per pixel.
//  * The never-written last row/column of the search area only drops the chip's
//    last row/column from the joint mask: those cells stay on the fast path with trimmed SAT
//    rectangles.  Cells that touch a real null pixel or the zero-filled image border are
//    re-evaluated with the masked 5-sum FP64 loop in a separate round.
//
// Work decomposition: a "group" of G threads owns one node at a time: one warp for chip
// half-widths 7/15 (eight nodes per CTA), half a CTA for 30, a whole 256-thread CTA for 40 and
// for wide search areas, a 128-thread CTA holding two chip rows per thread for the middle bin of
// 40 (Cfg, bin_table).  Thread (k, r) keeps the L pixels of chip row r, segment k in REGISTERS;
// the search area is staged in shared memory (loads issued eight deep) with an odd pitch so that
// the row-per-lane access pattern is bank-conflict free.  The reference's hill-climbing state
// machine (pivot order, first-wins ties, "newly evaluated" stop rule, -2.0 placeholders) runs
// verbatim in warp 0 of the group, its state kept in the shared control block; the first 3x3
// probe of every pivot is unconditional and is evaluated up-front in rounds of <= 32 cells, then
// the second probes of all pivots are predicted lane-parallel and evaluated in one round.
// Launches are split into shared-memory bins by the nodes' pivot extent (bin_table), each bin
// with the instantiation compiled for its number of resident CTAs (register budget); nodes that
// fit no bin go to the general kernel within the same call.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
// per-cell state.  cval[] holds kUnknownBits (a NaN no computation produces) until the cell has been evaluated;
// cflag[]: kVisible = the reference's state machine has "evaluated" the cell (cmap >= -1, MIMC_module.c:713),
// kListed = queued for evaluation; ctab[]: kTabComplete | k once all nine cells of the 3x3 probe centred here
// are known, k = scan index (u outer, v inner, :709-741) of the first occurrence of their maximum.
constexpr unsigned char kVisible = 2, kListed = 4, kTabComplete = 0x10;
constexpr unsigned int kUnknownBits = 0x7fc0deadu;
constexpr int kWalkDone = 1 << 30;
__device__ __forceinline__ int job_cy(int job) { return (job >> 16) & 0x7fff; }
__device__ __forceinline__ int job_cx(int job) { return job & 0xffff; }

struct Match2Args {
    const float *ref, *srch;
    int H, W;
    const ulonglong2 *sat_ref, *sat_srch;
    double inv_ref, inv_ref2, inv_srch, inv_srch2;   // 2^-f, 4^-f of the two images
    const int2 *node_uv;
    int off_u, off_v;
    const int *csr_off;
    const int2 *piv;
    int sign;
    const int *node_list;
    int n_list;
    float negate;
    float *dp;
    int2 *peak;
    int *ncell;
    unsigned int *counter;          // dynamic node fetch for this launch
    int *overflow_list;             // nodes that do not fit this launch's shared memory
    unsigned int *overflow_count;
    int grp_bytes;                  // shared memory per group: search area + cmap values + cmap flags
    int ngroups;                    // node groups per CTA in this launch (blockDim.x / G)
    float A0, Mlo;                  // accumulator biases
    unsigned int A0_bits, Mlo_bits;
    double hi_unit, lo_unit;
    float min_dn;
    int num_sms;                    // leader-warp rotation: co-resident CTAs lead from different SM sub-partitions
};

#ifdef MIMC3CU_PROFILE
// cycles seen by lane 0 of the group's leader warp: 0 node total, 1 staging, 2 leader section (7 walk, 8 requests, 9 replay + fit),
// 3 compute (+ barrier waits), 4 finalize + masked round, 10 probe-table update; 5 rounds, 6 nodes
__device__ unsigned long long g_prof[24];
__device__ unsigned int g_busy[256];   // per SM: node streams inside a fast round right now (slots 11..: histogram seen mid-round)
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, v) do { if (lane == 0 && gwarp == lead) atomicAdd(&g_prof[slot], (unsigned long long)(v)); } while (0)
#else
#define PROF_T(var)
#define PROF_ADD(slot, v)
#endif

struct Sums {
    double sx, sy, sxx, syy, sxy;
    int n;
};

// Resident CTAs per SM each instantiation is compiled for (register cap = 65536 / (256 * n)) and
// that the first shared-memory bin is sized for: ocw 40 keeps 27 chip pixels per thread (80
// registers, 3 CTAs), ocw 30 keeps 31 in half-CTA groups; measured on B200 (profiles/).
// ocw 30 runs two 128-thread groups (two nodes) per CTA, 3 CTAs = 6 nodes per SM; its 256-thread
// instantiation only serves the wide-search-area bins (<= 2 CTAs per SM).
// ocw 40: one 256-thread CTA per node, four per SM at 64 registers; its wider-search-area bins run 128-thread CTAs
// (two chip rows per thread, three per SM) or 256-thread CTAs compiled for two per SM (no register pressure).
// The 256-thread instantiations of ocw 7/15 only serve bins with <= 3 CTAs per SM.
constexpr int min_ctas(int ocw, int G) { return G == 64 ? 2 : (G == 256 && ocw < 30 ? 3 : (ocw == 15 ? 2 : (ocw == 30 ? 3 : (G == 128 ? 3 : 4)))); }

template <int OCW, int G>
struct Cfg {
    static constexpr int S = 2 * OCW + 1;
    // A thread keeps RB chip rows (r, r + RPT, ...) of one column segment: RB = 2 lets half as many threads
    // hold the 81x81 chip (54 pixels each), which is what fits five nodes into an SM's register file.
    static constexpr int RB = ((OCW == 40 && G == 128) || (OCW == 30 && G == 64)) ? 2 : 1;
    static constexpr int RPT = (S + RB - 1) / RB;            // rows per row block
    static constexpr int NSEG = (G / RPT) < S ? (G / RPT) : S;   // row segments per chip row (at most one pixel each)
    static constexpr int L = (S + NSEG - 1) / NSEG;
    static constexpr int NGROUPS = (OCW == 40 && G == 128) ? 1 : kThreads / G;   // groups (nodes) per CTA
    static constexpr int CTA = G * NGROUPS;
    static constexpr int NWARPS = G / 32;
    // cells per evaluation round: thread c of the group owns cell c (SAT look-up, normalisation), so the ~39
    // first-probe cells of a static node go in one round wherever a group has 64 threads
    static constexpr int MAXJ = G == 32 ? 32 : 64;
    // pivots per node this instantiation can walk (explore state in shared memory); longer pivot lines go to
    // a wider instantiation or to the general kernel
    static constexpr int PMAX = G == 32 ? 64 : 128;
    // cells evaluated side by side in the inner loop (independent accumulator chains)
#ifndef MIMC3CU_CP
#define MIMC3CU_CP 2
#endif
    static constexpr int CP = (G >= 128) ? MIMC3CU_CP : 1;
    // 16-byte staging loads where the register budget has room for them (measured: +5 % at chip half-width 15; the
    // 64- and 80-register instantiations of the other sizes spill with them and lose 2-13 %)
    static constexpr bool VEC = OCW == 15;
    // chip pixels straight from global memory into registers instead of through the shared-memory tile: pays for the
    // smallest chip only (+7 % at half-width 7; the row-per-lane pattern is uncoalesced and costs 10-20 % at 30 / 40,
    // even when the warps that idle during the previous node's replay do the loading)
    static constexpr bool DIRECT_CHIP = OCW == 7 && G == 32;
#ifndef MIMC3CU_ASYNC
#define MIMC3CU_ASYNC 9
#endif
    // staging by asynchronous 4-byte copies: bit 0 search area / bit 1 chip of the 128- and 256-thread groups,
    // bit 2 search area at half-width 15, bit 3 at half-width 7 (32-thread groups)
    static constexpr bool ASYNC_AREA = ((MIMC3CU_ASYNC & 1) && G >= 128) || ((MIMC3CU_ASYNC & 4) && G == 32 && OCW == 15) ||
                                       ((MIMC3CU_ASYNC & 8) && G == 32 && OCW == 7);
    static constexpr bool ASYNC_CHIP = ((MIMC3CU_ASYNC & 2) && G >= 128) || ((MIMC3CU_ASYNC & 16) && G == 32 && OCW == 15);
    static_assert(NSEG >= 1, "group too small for this chip");
    static_assert((L + 1) / 2 <= 16, "at most 16 pixels per FP32 accumulator");
};

// Copies `rows` x `width` floats (global row stride `gstride`) into a shared tile with row pitch
// `spitch` (columns [width, spitch) are zero-filled), NT threads, thread index `tix`.  Loads are
// issued DEPTH at a time before the first store: staging is latency-bound, not bandwidth-bound.
template <int NT, int DEPTH>
__device__ __forceinline__ void stage_rows(const float *__restrict__ src, int gstride, float *dst, int spitch, int rows, int width,
                                           int tix) {
    const int total = rows * spitch;
    int y = tix / spitch, x = tix - y * spitch;
    const int dy = NT / spitch, dx = NT - dy * spitch;
    for (int e = tix; e < total; e += NT * DEPTH) {
        float v[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; u++) {
            const bool ok = e + u * NT < total && x < width;
            v[u] = ok ? __ldg(src + (size_t)y * gstride + x) : 0.0f;
            x += dx; y += dy;
            if (x >= spitch) { x -= spitch; y++; }
        }
#pragma unroll
        for (int u = 0; u < DEPTH; u++)
            if (e + u * NT < total) dst[e + u * NT] = v[u];
    }
}

// Same copy with asynchronous 4-byte global -> shared copies (LDGSTS): no register holds a pixel, so every load of the
// thread is in flight at once instead of DEPTH at a time; pad columns are zero-filled by the copy itself (src-size 0).
// The caller commits / waits (cp_async_wait_all) and then synchronises the group.
__device__ __forceinline__ void cp_async4(float *dst, const float *src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
template <int NT>
__device__ __forceinline__ void stage_rows_async(const float *__restrict__ src, int gstride, float *dst, int spitch, int rows, int width,
                                                 int tix) {
    const int total = rows * spitch;
    int y = tix / spitch, x = tix - y * spitch;
    const int dy = NT / spitch, dx = NT - dy * spitch;
#pragma unroll 4
    for (int e = tix; e < total; e += NT) {
        const bool ok = x < width;
        cp_async4(dst + e, ok ? src + (size_t)y * gstride + x : src, ok);
        x += dx; y += dy;
        if (x >= spitch) { x -= spitch; y++; }
    }
}

// Same copy with 16-byte global loads: every row is read from the 16-byte aligned address at or below its first
// pixel (`a` floats earlier; the same for all rows because the row stride is a multiple of four floats), so a row of
// `width` pixels takes (a + width + 3) / 4 LDG.128 instead of `width` LDG.32, four in flight per thread; the stores
// stay scalar (the tile pitch is odd so that the row-per-lane reads of the inner loop are conflict-free).
// The caller guarantees gstride % 4 == 0 and that the aligned spans stay inside the image buffer.
template <int NT, int DEPTH>
__device__ __forceinline__ void stage_rows_v4(const float *__restrict__ src, int gstride, float *dst, int spitch, int rows, int width,
                                              int tix) {
    const int a = (int)(((size_t)src >> 2) & 3);
    const int nch = (a + width + 3) >> 2;
    const int total = rows * nch;
    const float4 *base = (const float4 *)(src - a);
    const int g4 = gstride >> 2;
    int y = tix / nch, j = tix - y * nch;
    const int dy = NT / nch, dj = NT - dy * nch;
    for (int e = tix; e < total; e += NT * DEPTH) {
        float4 v[DEPTH];
        int yj[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; u++) {
            yj[u] = (y << 16) | j;
            if (e + u * NT < total) v[u] = __ldg(base + (size_t)y * g4 + j);
            j += dj; y += dy;
            if (j >= nch) { j -= nch; y++; }
        }
#pragma unroll
        for (int u = 0; u < DEPTH; u++) {
            if (e + u * NT < total) {
                float *row = dst + (yj[u] >> 16) * spitch;
                const int c = 4 * (yj[u] & 0xffff) - a;
                if (c >= 0 && c < width) row[c] = v[u].x;
                if (c + 1 >= 0 && c + 1 < width) row[c + 1] = v[u].y;
                if (c + 2 >= 0 && c + 2 < width) row[c + 2] = v[u].z;
                if (c + 3 < width) row[c + 3] = v[u].w;
            }
        }
    }
    const int pad = spitch - width;
    for (int i = tix; i < rows * pad; i += NT) {
        const int r = i / pad;
        dst[r * spitch + width + (i - r * pad)] = 0.0f;
    }
}

// Packed FP32 pairs (sm_100 FMUL2 / FADD2 / FFMA2): each lane rounds exactly like the scalar instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
// The product is written as fma(x, y, +0): ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one
// FFMA2 despite the explicit roundings, which would skip the float rounding of the product the reference
// performs; FFMA2 with RZ as addend is left alone (operands are non-negative, so no signed-zero issue).
// The u16 parity tests (products beyond 24 bits) fail if a toolchain ever changes this.
__device__ __forceinline__ f32x2 mul2(f32x2 x, f32x2 y) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(0ull)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 x, f32x2 y) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 x, f32x2 y) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 x, f32x2 y, f32x2 z) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z)); return r; }

// Order-preserving map float -> uint32 (NaN -> 0), for warp arg-max with REDUX.
__device__ __forceinline__ unsigned int ordered_key(float v) {
    if (v != v) return 0u;
    const unsigned int b = __float_as_uint(v + 0.0f);   // -0.0f -> +0.0f
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int G>
__device__ __forceinline__ void gsync() {
    if (G == 32) __syncwarp();
    else if (G == kThreads) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / G)), "n"(G) : "memory");   // one named barrier per group
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (float)((n*sxy - sx*sy) / sqrt((n*sxx - sx*sx) * (n*syy - sy*sy)))   MIMC_module.c:734
__device__ __forceinline__ float ncc_from_sums(const Sums &s) {
    double n = (double)s.n;
    double num = __dsub_rn(__dmul_rn(n, s.sxy), __dmul_rn(s.sx, s.sy));
    double a = __dsub_rn(__dmul_rn(n, s.sxx), __dmul_rn(s.sx, s.sx));
    double b = __dsub_rn(__dmul_rn(n, s.syy), __dmul_rn(s.sy, s.sy));
    double den = __dsqrt_rn(__dmul_rn(a, b));
    return __double2float_rn(__ddiv_rn(num, den));
}

// Window sums over image rows [y0,y1) x columns [x0,x1) (already clipped, possibly empty).
__device__ __forceinline__ void rect_query(const ulonglong2 *sat, int W1, int x0, int y0, int x1, int y1,
                                           unsigned long long &ss, unsigned long long &s, unsigned int &nul) {
    if (x1 <= x0 || y1 <= y0) { ss = 0; s = 0; nul = 0; return; }
    const ulonglong2 a = __ldg(&sat[(size_t)y1 * W1 + x1]), b = __ldg(&sat[(size_t)y0 * W1 + x1]);
    const ulonglong2 c = __ldg(&sat[(size_t)y1 * W1 + x0]), d = __ldg(&sat[(size_t)y0 * W1 + x0]);
    ss = a.x - b.x - c.x + d.x;
    const unsigned long long pk = a.y - b.y - c.y + d.y;
    s = pk >> 24;
    nul = (unsigned int)(pk & 0xffffffull);
}

// Same, leaving value sum and null count packed (fewer live registers for prefetched queries).
__device__ __forceinline__ void rect_query_packed(const ulonglong2 *sat, int W1, int x0, int y0, int x1, int y1,
                                                  unsigned long long &ss, unsigned long long &pk) {
    if (x1 <= x0 || y1 <= y0) { ss = 0; pk = 0; return; }
    const ulonglong2 a = __ldg(&sat[(size_t)y1 * W1 + x1]), b = __ldg(&sat[(size_t)y0 * W1 + x1]);
    const ulonglong2 c = __ldg(&sat[(size_t)y1 * W1 + x0]), d = __ldg(&sat[(size_t)y0 * W1 + x0]);
    ss = a.x - b.x - c.x + d.x;
    pk = a.y - b.y - c.y + d.y;
}

// 3x3 quadratic fit, MIMC_module.c:757-788, with the reference's float/double mix (H7).
__device__ void subpixel_fit(const float n9[9], int peak_du, int peak_dv, float &du, float &dv) {
#define FM(k, x) __fmul_rn((float)(k), (x))
#define FA(x, y) __fadd_rn((x), (y))
    float c0f = FA(FA(FA(FA(FA(FA(FA(FA(FM(6, n9[0]), -FM(12, n9[1])), FM(6, n9[2])), FM(6, n9[3])), -FM(12, n9[4])), FM(6, n9[5])), FM(6, n9[6])), -FM(12, n9[7])), FM(6, n9[8]));
    float c1f = FA(FA(FA(FM(9, n9[0]), -FM(9, n9[2])), -FM(9, n9[6])), FM(9, n9[8]));
    float c2f = FA(FA(FA(FA(FA(FA(FA(FA(FM(6, n9[0]), FM(6, n9[1])), FM(6, n9[2])), -FM(12, n9[3])), -FM(12, n9[4])), -FM(12, n9[5])), FM(6, n9[6])), FM(6, n9[7])), FM(6, n9[8]));
    float c3f = FA(FA(FA(FA(FA(FM(-6, n9[0]), FM(6, n9[2])), -FM(6, n9[3])), FM(6, n9[5])), -FM(6, n9[6])), FM(6, n9[8]));
    float c4f = FA(FA(FA(FA(FA(FM(-6, n9[0]), -FM(6, n9[1])), -FM(6, n9[2])), FM(6, n9[6])), FM(6, n9[7])), FM(6, n9[8]));
#undef FM
#undef FA
    double c0 = __ddiv_rn((double)c0f, 36.0), c1 = __ddiv_rn((double)c1f, 36.0), c2 = __ddiv_rn((double)c2f, 36.0);
    double c3 = __ddiv_rn((double)c3f, 36.0), c4 = __ddiv_rn((double)c4f, 36.0);
    float fu = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(-2.0, c2), c3), __dmul_rn(c1, c4)));
    float fv = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(-2.0, c0), c4), __dmul_rn(c1, c3)));
    double det = __dsub_rn(__dmul_rn(__dmul_rn(4.0, c0), c2), __dmul_rn(c1, c1));
    fu = __double2float_rn(__ddiv_rn((double)fu, det));
    fv = __double2float_rn(__ddiv_rn((double)fv, det));
    du = __fadd_rn(fu, (float)peak_du);
    dv = __fadd_rn(fv, (float)peak_dv);
}

// The thread's slice of the reference chip (extract_refchip, MIMC_module.c:845-855: no bounds check there; zero outside
// the image here) straight from global memory into registers: rows r, r + RPT, ... of column segment col0, `len` pixels
// each.  A lane per chip row is an uncoalesced pattern, but all loads of the slice are in flight at once, nothing goes
// through shared memory (the search area can be staged at the same time).  Used where Cfg::DIRECT_CHIP says it pays.
template <int OCW, int G>
__device__ __forceinline__ void load_chip_direct(const Match2Args &a, int u0, int v0, int r, int col0, int len,
                                                 float (&chip)[Cfg<OCW, G>::RB][Cfg<OCW, G>::L]) {
    using C = Cfg<OCW, G>;
    constexpr int S = C::S, L = C::L;
    const bool inside = u0 - OCW >= 0 && v0 - OCW >= 0 && u0 + OCW < a.W && v0 + OCW < a.H;
#pragma unroll
    for (int rb = 0; rb < C::RB; rb++) {
        const int row = r + rb * C::RPT;
        const int iv = v0 - OCW + row;
        const float *src = a.ref + (size_t)min(max(iv, 0), a.H - 1) * a.W + (u0 - OCW + col0);
        const bool rowok = row < S && (inside || (iv >= 0 && iv < a.H));
#pragma unroll
        for (int c = 0; c < L; c++) {
            const int iu = u0 - OCW + col0 + c;
            const bool ok = rowok && c < len && (inside || (iu >= 0 && iu < a.W));
            chip[rb][c] = ok ? __ldg(src + c) : 0.0f;
        }
    }
}

// Per-group control block in shared memory.
template <int NWARPS, int MAXJ, int PMAX>
struct Ctl {
    int job[MAXJ];                   // cells of the current round: cy << 16 | cx
    int slowjob[MAXJ];               // cells of the round that need the masked (null-excluding) evaluation
    int2 part[NWARPS][MAXJ];         // per-warp integer partial sums (hi units, lo units)
    Sums partd[NWARPS];              // masked-path partials
    // explore state of every pivot (cell coordinates cx | cy << 15 | kWalkDone, running maximum) and where it started
    int wpos[PMAX], wstart[PMAX];
    float wmax[PMAX];
    // Node geometry lives HERE, not in registers: only the leader warp needs it, and only between evaluation rounds;
    // keeping it out of the register file of the compute loop is what lets the 81x81 instantiation keep its chip
    // pixels without remat/spills.
    struct Geo { int g, P, su0, sv0, dx2, dy2, Dx2, Dy2, cw, ch, sa_elems, cell_elems; } geo;
    int blk[PMAX];              // probe centres (cell index) of the pivots blocked in this round
    int nblk;
    int m;                      // cells requested for this round (may exceed MAXJ: the excess is not listed)
    int nslow;                  // cells in slowjob[]
    int bbox[4];                // cx0, cx1, cy0, cy1 of the round's cells (probe table update)
    unsigned int node[2];       // dynamic node fetch, double-buffered: the next index is fetched a node ahead
    int valid;
    // chip constants from the reference SAT
    // sum(fl(r*r)), sum(r) of the chip without its last (tx) column / (ty) row: index tx + 2*ty
    unsigned long long chip_ss[4], chip_s[4];
    int chip_fast;
};

// Evaluation rounds of one node.  The reference's climb (MIMC_module.c:691-753) asks for NCC cells one 3x3 probe
// at a time, pivot after pivot; a round per probe would leave the group idle behind a serial state machine.
// Instead (oracle/leader_model.c restates this schedule on the CPU and the test-suite asserts that it is
// equivalent to the reference's):
//   explore  lane i of the leader warp walks pivot i's climb on the values known so far, ignoring the
//            reference's "nothing new in this probe => stop" rule -- that rule can only SHORTEN a path, so the
//            walked cells are a superset of the reference's (by +0.1 %).  One step is a look-up in the probe
//            table ctab (all nine cells known?  where is their maximum?), which all threads keep up to date
//            after every round.  A pivot whose next probe has unknown cells is blocked and lists them; the
//            lists of all blocked pivots form one evaluation round (<= MAXJ cells).
//   replay   once no pivot is blocked the reference's state machine runs verbatim over the known values and
//            decides visibility, the evaluated-cell count, the peak and its NCC.
// 4.4 rounds per node instead of 7.5 (fast glaciers: 9 instead of ~38), and no speculative cells.
template <int OCW, int G, bool EXACTP, typename CtlT>
__device__ __forceinline__ void node_rounds(const Match2Args &a, CtlT &ctl, float *sa, const float *sa_thread,
                                            const float (&chip)[Cfg<OCW, G>::RB][Cfg<OCW, G>::L], const int pitch, const int row2, const bool active, const int t,
                                            const int lane, const int gwarp, const int lead) {
    using C = Cfg<OCW, G>;
    constexpr int S = C::S, L = C::L;
    const int W1 = a.W + 1;
    const int cw = ctl.geo.cw, ch = ctl.geo.ch;
    float *cval = sa + ctl.geo.sa_elems;
    unsigned char *cflag = (unsigned char *)(cval + ctl.geo.cell_elems);
    unsigned char *ctab = cflag + ctl.geo.cell_elems;
    for (;;) {
        PROF_T(t_p0);
        // ---- explore: thread i walks pivot i as far as the probe table allows ------------------------------
        const int P = ctl.geo.P;
        if (t == 0) { ctl.bbox[0] = 0x7fff; ctl.bbox[1] = 0; ctl.bbox[2] = 0x7fff; ctl.bbox[3] = 0; }
        for (int i = t; i < P; i += G) {
            const int st = ctl.wpos[i];
            if (st & kWalkDone) continue;
            int cx = st & 0x7fff, cy = (st >> 15) & 0x7fff;
            float nmax = ctl.wmax[i];
            bool done = false;
            for (;;) {
                if ((unsigned)(cx - 1) > (unsigned)(cw - 3) || (unsigned)(cy - 1) > (unsigned)(ch - 3)) { done = true; break; }   // :701-707
                const unsigned char e = ctab[cy * cw + cx];
                if (!(e & kTabComplete)) break;                  // blocked: some of the nine cells are unknown
                const int k = e & 15;
                const int nx = cx + k / 3 - 1, ny = cy + k % 3 - 1;
                const float v = cval[ny * cw + nx];
                if (!(v > nmax)) { done = true; break; }      // nothing beats the running maximum: duv = 0
                nmax = v;
                if (k == 4) { done = true; break; }            // the maximum is the centre: duv = 0
                cx = nx; cy = ny;
            }
            ctl.wpos[i] = cx | (cy << 15) | (done ? kWalkDone : 0);
            ctl.wmax[i] = nmax;
            if (!done) ctl.blk[atomicAdd(&ctl.nblk, 1)] = cy * cw + cx;
        }
        PROF_T(t_w1);
        PROF_ADD(7, t_w1 - t_p0);
        gsync<G>();
        const int nblk = ctl.nblk;
        if (nblk == 0) {
            // ---- replay: the reference's state machine, MIMC_module.c:691-753, on the known values ---------
            // (the other warps wait for the leader at the group barrier that opens the next node)
            if (gwarp == lead) {
                PROF_T(t_r0);
                const int dx2 = ctl.geo.dx2, dy2 = ctl.geo.dy2;
                int pkx = dx2 - OCW - 1, pky = dy2 - OCW - 1, ncells = 0;
                float best = -2.0f;
                // lane k < 9 looks at probe cell k = (c1+1)*3 + (c2+1)  (c1: u outer, c2: v inner)
                const int loff = lane < 9 ? (lane % 3 - 1) * cw + (lane / 3 - 1) : 0;
                for (int ip = 0; ip < P; ip++) {
                    const int st = ctl.wstart[ip];
                    int cx = st & 0x7fff, cy = st >> 15;                                  // :693-694
                    float nccmax = -2.0f;
                    for (;;) {   // one iteration per probe (:699-745); every exit leaves (cx, cy) where the reference's loop ends
                        if ((unsigned)(cx - 1) > (unsigned)(cw - 3) || (unsigned)(cy - 1) > (unsigned)(ch - 3)) break;   // :701-707
                        const int pos = cy * cw + cx;
                        const unsigned char f = cflag[pos + loff];
                        const float v = cval[pos + loff];
                        const int k = ctab[pos] & 15;
                        const bool isnew = lane < 9 && (!(f & kVisible) || v < -1.0f);   // `cmap < -1.0` => evaluated now, :713
                        const unsigned int newm = __ballot_sync(0xffffffffu, isnew);
                        if (isnew) cflag[pos + loff] = f | kVisible;
                        ncells += __popc(newm);
                        // :736-741: sequential strict `>` scan == first occurrence of the maximum, taken only if it
                        // beats the running nccmax (NaN never wins)
                        const float vb = __shfl_sync(0xffffffffu, v, k);
                        if (!(vb > nccmax)) break;                                         // duv = 0
                        nccmax = vb;
                        if (k == 4) break;                                                 // duv = 0
                        cx += k / 3 - 1; cy += k % 3 - 1;                                  // :744-745
                        if (newm == 0) break;                                              // flag_newncc == 0, :699
                    }
                    if (nccmax > best) { pkx = cx; pky = cy; best = nccmax; }             // :747-752
                    __syncwarp();
                }
                // ---- sub-pixel fit and output (:757-788) -------------------------------------------------
                if (lane == 0) {
                    const int g = ctl.geo.g;
                    float n9[9];
                    for (int rr = 0; rr < 3; rr++)
                        for (int cc = 0; cc < 3; cc++) {
                            const int cx = pkx - 1 + cc, cy = pky - 1 + rr;
                            float v = -2.0f;   // never evaluated (or outside the evaluable region)
                            if (cx >= 0 && cx < cw && cy >= 0 && cy < ch) {
                                const int cell = cy * cw + cx;
                                if (cflag[cell] & kVisible) v = cval[cell];
                            }
                            n9[rr * 3 + cc] = v;
                        }
                    const int peak_du = pkx + OCW + 1 - dx2, peak_dv = pky + OCW + 1 - dy2;
                    float du, dv;
                    subpixel_fit(n9, peak_du, peak_dv, du, dv);
                    a.dp[3 * (size_t)g] = a.negate * du;
                    a.dp[3 * (size_t)g + 1] = a.negate * dv;
                    a.dp[3 * (size_t)g + 2] = best;
                    if (a.peak) a.peak[g] = make_int2(peak_du, peak_dv);
                    if (a.ncell) a.ncell[g] = ncells;
                }
                PROF_T(t_r1);
                PROF_ADD(9, t_r1 - t_r0);
                PROF_ADD(2, t_r1 - t_p0);
            }
            break;
        }
        // ---- requests: the unknown cells of the blocked pivots' probes form the round (<= MAXJ cells; a cell
        //      that does not get in stays unlisted and its pivot asks again) --------------------------------------
        for (int idx = t; idx < nblk * 9; idx += G) {
            const int j = idx / 9, k = idx - 9 * j;
            const int pos = ctl.blk[j];
            const int cell = pos + (k % 3 - 1) * cw + (k / 3 - 1);
            if (__float_as_uint(cval[cell]) != kUnknownBits) continue;
            unsigned int *word = (unsigned int *)cflag + (cell >> 2);
            const unsigned int bit = (unsigned int)kListed << (8 * (cell & 3));
            if (atomicOr(word, bit) & bit) continue;             // listed by another pivot
            const int slot = atomicAdd(&ctl.m, 1);
            if (slot < C::MAXJ) {
                const int cy = cell / cw;
                ctl.job[slot] = (cy << 16) | (cell - cy * cw);
            } else {
                atomicAnd(word, ~bit);
            }
        }
        PROF_T(t_p1);
        PROF_ADD(8, t_p1 - t_w1);
        PROF_ADD(2, t_p1 - t_p0);
        gsync<G>();
        const int m = min(ctl.m, C::MAXJ);
        PROF_ADD(5, 1);

        // ---- fast round: sum(fl(r*s)) for m cells, exact in FP32 ------------------------------------------
        // Thread c < m owns cell c: its SAT corner loads are issued first so that their latency hides behind
        // the loop.  sum(fl(s*s)), packed (sum(s) << 24 | nulls), flags = inside | (tx + 2*ty) << 1 where tx/ty
        // say that the window reaches the never-written last column / row of the search area
        unsigned long long w_ss = 0, w_pk = 1;
        int w_flag = 0;
        if (t < m) {
            const int su0 = ctl.geo.su0, sv0 = ctl.geo.sv0, dx2 = ctl.geo.dx2, dy2 = ctl.geo.dy2, Dx2 = ctl.geo.Dx2, Dy2 = ctl.geo.Dy2;
            const int job = ctl.job[t];
            const int x0 = job_cx(job) + 1, y0 = job_cy(job) + 1;                     // window origin in the search area
            const int ix0 = su0 - dx2 + x0, iy0 = sv0 - dy2 + y0;     // ... and in the image
            // The last column / row of the search area is never written by the reference (H1): those pixels are
            // nulls, i.e. the joint mask simply drops the chip's last column / row.  The FP32 loop already sees
            // zeros there; only the SAT rectangles and n shrink.
            const int tx = (x0 + S - 1 == Dx2 - 1), ty = (y0 + S - 1 == Dy2 - 1);
            const bool inside = ix0 >= 0 && iy0 >= 0 && ix0 + S - tx <= a.W && iy0 + S - ty <= a.H;
            w_flag = (inside ? 1 : 0) | ((tx + 2 * ty) << 1);
            if (inside) rect_query_packed(a.sat_srch, W1, ix0, iy0, ix0 + S - tx, iy0 + S - ty, w_ss, w_pk);
        }
        // CP cells at a time: their accumulator chains are independent, which keeps the FP32 pipe fed when few warps are
        // resident (the wide-search-area bins: +8 % on the fast-glacier scene; neutral at four CTAs per SM)
        constexpr int CP = C::CP;
#ifdef MIMC3CU_PROFILE
        unsigned int prof_smid = 0;
        if (lane == 0 && gwarp == lead) { asm("mov.u32 %0, %%smid;" : "=r"(prof_smid)); atomicAdd(&g_busy[prof_smid & 255u], 1u); }
#endif
        for (int c = 0; c < m; c += CP) {
#ifdef MIMC3CU_PROFILE
            if (lane == 0 && gwarp == lead && c == ((m / 2) & ~(CP - 1))) {
                const unsigned int k = *(volatile unsigned int *)&g_busy[prof_smid & 255u];
                atomicAdd(&g_prof[11 + min(k, 8u)], 1ull);
            }
#endif
            const float *spc[CP];
#pragma unroll
            for (int q = 0; q < CP; q++) {
                const int job = ctl.job[min(c + q, m - 1)];   // an odd tail evaluates its last cell twice
                spc[q] = sa_thread + (job_cy(job) + 1) * pitch + (job_cx(job) + 1);
            }
            unsigned int hi[CP];
            int lo[CP];
#pragma unroll
            for (int q = 0; q < CP; q++) { hi[q] = 0; lo[q] = 0; }
#pragma unroll
            for (int rb = 0; rb < C::RB; rb++) {
                // threads without chip pixels (r = col0 = 0, chip all zero) run the same code: no branch; a
                // thread whose second row does not exist reads its first row again (times zero)
                // Two pixels per instruction (FMUL2 / FADD2 / FFMA2, sm_100): lane 0 of the packed pair
                // accumulates the even pixels, lane 1 the odd ones -- the same two accumulators as a
                // scalar loop would keep, at half the issue slots.  One accumulator pair per chip row
                // (<= 16 pixels per accumulator).
                f32x2 acc[CP], lo2[CP];
#pragma unroll
                for (int q = 0; q < CP; q++) { acc[q] = pack2(a.A0, a.A0); lo2[q] = pack2(a.Mlo, a.Mlo); }
                // The search-area pixels go through a ring of PD register pairs: the pair of step k + PD is requested
                // when step k starts, so the shared-memory latency (~30 cycles) is off the accumulator chains.  The
                // loads are volatile asm: ptxas otherwise sinks every LDS to just before its FFMA2 and each step
                // waits out the full latency (17 % of all warp time in the round-1 build, ncu source view).
#pragma unroll
                for (int k = 0; k < L; k += 2) {
                    // an odd L ends with a (pixel, 0) pair: the zero is a literal, not a load
                    const f32x2 rv = pack2(chip[rb][k], k + 1 < L ? chip[rb][k + 1] : 0.0f);
#pragma unroll
                    for (int q = 0; q < CP; q++) {
                        const float *sp = spc[q] + (rb ? row2 : 0);
                        const f32x2 sv = pack2(sp[k], k + 1 < L ? sp[k + 1] : 0.0f);
                        if (EXACTP) {
                            // every product is exact in FP32 (scaled operands < 2^12): fma(r, s, acc) ==
                            // fadd(acc, fmul(r, s)) and fma(r, s, -z) == p - z, one instruction less per pixel
                            const f32x2 s1 = fma2(rv, sv, acc[q]);
                            const f32x2 nz = sub2(acc[q], s1);
                            lo2[q] = add2(lo2[q], fma2(rv, sv, nz));
                            acc[q] = s1;
                        } else {
                            const f32x2 pr = mul2(rv, sv);
                            const f32x2 s1 = add2(acc[q], pr);
                            const f32x2 z = sub2(s1, acc[q]);
                            lo2[q] = add2(lo2[q], sub2(pr, z));
                            acc[q] = s1;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < CP; q++) {
                    float acc0, acc1, lo0, lo1;
                    unpack2(acc[q], acc0, acc1);
                    unpack2(lo2[q], lo0, lo1);
                    hi[q] += (__float_as_uint(acc0) - a.A0_bits) + (__float_as_uint(acc1) - a.A0_bits);
                    lo[q] += (int)(__float_as_uint(lo0) - a.Mlo_bits) + (int)(__float_as_uint(lo1) - a.Mlo_bits);
                }
            }
#pragma unroll
            for (int q = 0; q < CP; q++) {
                hi[q] = __reduce_add_sync(0xffffffffu, hi[q]);
                lo[q] = __reduce_add_sync(0xffffffffu, lo[q]);
                if (lane == 0 && c + q < m) ctl.part[gwarp][c + q] = make_int2((int)hi[q], lo[q]);
            }
        }
#ifdef MIMC3CU_PROFILE
        if (lane == 0 && gwarp == lead) atomicSub(&g_busy[prof_smid & 255u], 1u);
#endif
        gsync<G>();
        PROF_T(t_c1);
        PROF_ADD(3, t_c1 - t_p1);
        // ---- finalize: thread c normalises cell c -------------------------------------------------------------
        if (t < m) {
            const int job = ctl.job[t];
            const int cell = job_cy(job) * cw + job_cx(job);
            const int trim = w_flag >> 1;
            // bounding box of the round's cells: the probe-table update only looks there
            atomicMin(&ctl.bbox[0], job_cx(job)); atomicMax(&ctl.bbox[1], job_cx(job));
            atomicMin(&ctl.bbox[2], job_cy(job)); atomicMax(&ctl.bbox[3], job_cy(job));
            if (ctl.chip_fast && (w_flag & 1) && (w_pk & 0xffffffull) == 0) {
                long long hs = 0, ls = 0;
#pragma unroll
                for (int w = 0; w < C::NWARPS; w++) {
                    const int2 q = ctl.part[w][t];
                    hs += (unsigned int)q.x; ls += q.y;
                }
                Sums s;
                s.n = (S - (trim & 1)) * (S - (trim >> 1));
                s.sxy = (double)hs * a.hi_unit + (double)ls * a.lo_unit;
                s.sx = (double)ctl.chip_s[trim] * a.inv_ref; s.sxx = (double)ctl.chip_ss[trim] * a.inv_ref2;
                s.sy = (double)(w_pk >> 24) * a.inv_srch; s.syy = (double)w_ss * a.inv_srch2;
                cval[cell] = ncc_from_sums(s);
            } else {
                // the window or the chip holds null pixels (no-data, zero-filled image border): masked evaluation
                ctl.slowjob[atomicAdd(&ctl.nslow, 1)] = job;
            }
        }
        gsync<G>();
        // ---- masked round: the reference's 5-sum loop with null exclusion (:719-733), FP64 ------------------------
        const int nslow = ctl.nslow;
        if (nslow > 0) {
            for (int c = 0; c < nslow; c++) {
                const int job = ctl.slowjob[c];
                const int cy = job_cy(job), cx = job_cx(job);
                const int cell = cy * cw + cx;
                Sums s = {0.0, 0.0, 0.0, 0.0, 0.0, 0};
#pragma unroll
                for (int rb = 0; rb < C::RB; rb++) {
                    if (!active) break;
                    const float *sp = sa_thread + (cy + 1) * pitch + (cx + 1) + (rb ? row2 : 0);
#pragma unroll
                    for (int k = 0; k < L; k++) {
                        const float rv = chip[rb][k], sv = sp[k];
                        if (rv >= a.min_dn && sv >= a.min_dn) {   // null exclusion, :723
                            s.n++;
                            s.sx += (double)rv; s.sy += (double)sv;
                            s.sxx += (double)__fmul_rn(rv, rv);
                            s.syy += (double)__fmul_rn(sv, sv);
                            s.sxy += (double)__fmul_rn(rv, sv);
                        }
                    }
                }
                s.n = __reduce_add_sync(0xffffffffu, s.n);
                s.sx = warp_sum_d(s.sx); s.sy = warp_sum_d(s.sy);
                s.sxx = warp_sum_d(s.sxx); s.syy = warp_sum_d(s.syy); s.sxy = warp_sum_d(s.sxy);
                if (G == 32) {
                    if (lane == 0) cval[cell] = ncc_from_sums(s);
                } else {
                    if (lane == 0) ctl.partd[gwarp] = s;
                    gsync<G>();
                    if (t == 0) {
                        Sums q = ctl.partd[0];
                        for (int w = 1; w < C::NWARPS; w++) {
                            const Sums &z = ctl.partd[w];
                            q.n += z.n; q.sx += z.sx; q.sy += z.sy; q.sxx += z.sxx; q.syy += z.syy; q.sxy += z.sxy;
                        }
                        cval[cell] = ncc_from_sums(q);
                    }
                    gsync<G>();
                }
            }
            if (t == 0) ctl.nslow = 0;   // next written after two more group barriers
            gsync<G>();
        }
        // ---- probe table: positions whose nine cells are all known now ---------------------------------------
        PROF_T(t_t0);
        PROF_ADD(4, t_t0 - t_c1);
        {
            const int bx0 = max(ctl.bbox[0] - 1, 1), bx1 = min(ctl.bbox[1] + 1, cw - 2);
            const int by0 = max(ctl.bbox[2] - 1, 1), by1 = min(ctl.bbox[3] + 1, ch - 2);
            const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
            const int total = (bw > 0 && bh > 0) ? bw * bh : 0;
            for (int idx = t; idx < total; idx += G) {
                const int yy = idx / bw;
                const int pos = (by0 + yy) * cw + bx0 + (idx - yy * bw);
                if (ctab[pos] & kTabComplete) continue;
                float top = -CUDART_INF_F;
                int kb = 4;
                bool complete = true;
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const float v = cval[pos + (k % 3 - 1) * cw + (k / 3 - 1)];
                    complete = complete && __float_as_uint(v) != kUnknownBits;
                    if (v > top) { top = v; kb = k; }
                }
                if (complete) ctab[pos] = (unsigned char)(kTabComplete | kb);
            }
            if (t == 0) { ctl.m = 0; ctl.nblk = 0; }   // every thread has read both (two barriers ago)
        }
        PROF_T(t_f1);
        PROF_ADD(10, t_f1 - t_t0);
        gsync<G>();
    }
}

template <int OCW, int G, bool EXACTP, int MINCTA = min_ctas(OCW, G)>
__global__ void __launch_bounds__(Cfg<OCW, G>::CTA, MINCTA) match2_kernel(const Match2Args a) {
    using C = Cfg<OCW, G>;
    constexpr int S = C::S, L = C::L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Ctl<C::NWARPS, C::MAXJ, C::PMAX> ctl_all[C::NGROUPS];

    const int tid = threadIdx.x;
    const int grp = tid / G, t = tid - grp * G;
    const int lane = tid & 31, gwarp = t >> 5;
    Ctl<C::NWARPS, C::MAXJ, C::PMAX> &ctl = ctl_all[grp];
    // CTAs that share an SM (blockIdx.x = sm + j * num_sms in the first wave) run their serial sections in
    // warps of different SM sub-partitions
    const int lead = (int)((blockIdx.x / (unsigned)a.num_sms + grp) % C::NWARPS);

    // dynamic shared memory, per group and per node: sa[Dy2*pitch] floats | cval[cells] floats | cflag[cells] | ctab[cells] bytes
    float *sa = (float *)(smem_raw + (size_t)grp * a.grp_bytes);

    // this thread's chip slice: row r, columns [col0, col0+len)
    const int seg = t / C::RPT, r = t - seg * C::RPT;   // rows r, r + RPT, ... of column segment seg
    const bool active = seg < C::NSEG;
    const int col0 = seg * L;
    const int len = active ? min(L, S - col0) : 0;
    const int W1 = a.W + 1;

    if (t == 0) { ctl.node[0] = atomicAdd(a.counter, 1u); ctl.nslow = 0; }
    float chip[C::RB][L];        // this thread's chip pixels, in registers for the whole node
    for (int iter = 0;; iter++) {
        gsync<G>();
        const unsigned int idx = ctl.node[iter & 1];
        if (idx >= (unsigned int)a.n_list) break;
        if (t == 0) ctl.node[(iter + 1) & 1] = atomicAdd(a.counter, 1u);   // latency hidden behind this node
        const int g = a.node_list ? a.node_list[idx] : (int)idx;
        PROF_T(t_node0);

        // ---- node geometry (uniform over the group) -----------------------------------------
        const int pb = a.csr_off[g], P = a.csr_off[g + 1] - pb;
        const int2 *piv = a.piv + pb;
        const int2 uv = a.node_uv[g];
        const int u0 = uv.x, v0 = uv.y, su0 = uv.x + a.off_u, sv0 = uv.y + a.off_v;
        if (P <= 0) {   // undefined behaviour in the reference (MIMC_module.c:589-591)
            if (t == 0) {
                a.dp[3 * (size_t)g] = CUDART_NAN_F; a.dp[3 * (size_t)g + 1] = CUDART_NAN_F; a.dp[3 * (size_t)g + 2] = -2.0f;
                if (a.peak) a.peak[g] = make_int2(0, 0);
                if (a.ncell) a.ncell[g] = 0;
            }
            continue;
        }
        const int2 last = piv[P - 1];
        const int dx2 = abs(last.x) + OCW + 2, dy2 = abs(last.y) + OCW + 2;   // :863-866
        const int Dx2 = 2 * dx2 + 1, Dy2 = 2 * dy2 + 1;
        const int cw = Dx2 - 2 * OCW - 1, ch = Dy2 - 2 * OCW - 1;   // cells whose 3x3 probe can be requested
        const int pitch = (Dx2 + 3) | 1;
        const int sa_elems = (Dy2 * pitch + 3) & ~3, cell_elems = (cw * ch + 15) & ~15;
        if (sa_elems * 4 + cell_elems * 6 > a.grp_bytes || P > C::PMAX) {
            if (t == 0) a.overflow_list[atomicAdd(a.overflow_count, 1u)] = g;
            continue;
        }
        float *cval = sa + sa_elems;
        unsigned char *cflag = (unsigned char *)(cval + cell_elems);

        // ---- chip and search-area statistics from the SATs (investigate_valid_grid :605-644); the loads
        //      are issued first so that their latency hides behind the chip staging below ----------------
        unsigned long long q_ss = 0, q_s = 0;
        unsigned int q_nul = 0;
        int q_area = 0, q_full = 0;
        if (t == 0) {
            const int x0 = max(u0 - OCW, 0), y0 = max(v0 - OCW, 0), x1 = min(u0 + OCW + 1, a.W), y1 = min(v0 + OCW + 1, a.H);
            rect_query(a.sat_ref, W1, x0, y0, x1, y1, q_ss, q_s, q_nul);
            q_area = max(x1 - x0, 0) * max(y1 - y0, 0); q_full = S * S;
        } else if (t >= 2 && t <= 4) {   // the chip without its last column (t=2: tx), last row (t=3: ty), or both (t=4)
            const int tx = (t - 1) & 1, ty = (t - 1) >> 1;
            const int x0 = max(u0 - OCW, 0), y0 = max(v0 - OCW, 0), x1 = min(u0 + OCW + 1 - tx, a.W), y1 = min(v0 + OCW + 1 - ty, a.H);
            rect_query(a.sat_ref, W1, x0, y0, x1, y1, q_ss, q_s, q_nul);
        } else if (t == 1) {   // written part of the search area: rows [0,Dy2-1) x columns [0,Dx2-1)
            const int ax = su0 - dx2, ay = sv0 - dy2;
            const int x0 = max(ax, 0), y0 = max(ay, 0), x1 = min(ax + Dx2 - 1, a.W), y1 = min(ay + Dy2 - 1, a.H);
            rect_query(a.sat_srch, W1, x0, y0, x1, y1, q_ss, q_s, q_nul);
            q_area = max(x1 - x0, 0) * max(y1 - y0, 0); q_full = Dx2 * Dy2;
        }
        // every pivot starts its climb at its own cell (:693-694); cell (cx, cy) <-> search-area position (cx+OCW+1, cy+OCW+1)
        for (int i = t; i < P; i += G) {
            const int2 pv = piv[i];
            const int st = (a.sign * pv.x + dx2 - OCW - 1) | ((a.sign * pv.y + dy2 - OCW - 1) << 15);
            ctl.wstart[i] = st; ctl.wpos[i] = st; ctl.wmax[i] = -2.0f;
        }

        // ---- the chip (extract_refchip :845-855): into registers, straight from global memory or through the shared tile ----
        const bool vec_ok = (a.W & 3) == 0;   // 16-byte loads need a row stride of whole float4s
        if (C::DIRECT_CHIP) {
            load_chip_direct<OCW, G>(a, u0, v0, r, col0, len, chip);
        } else if (u0 - OCW >= 0 && v0 - OCW >= 0 && u0 + OCW < a.W && v0 + OCW < a.H) {
            const size_t first = (size_t)(v0 - OCW) * a.W + (u0 - OCW);
            if (C::ASYNC_CHIP) {
                stage_rows_async<G>(a.ref + first, a.W, sa, S, S, S, t);
                cp_async_wait_all();
            } else if (C::VEC && vec_ok && first >= 3 && first + (size_t)(S - 1) * a.W + S + 3 <= (size_t)a.H * a.W)
                stage_rows_v4<G, 4>(a.ref + first, a.W, sa, S, S, S, t);
            else
                stage_rows<G, 8>(a.ref + first, a.W, sa, S, S, S, t);
        } else {
            for (int i = t; i < S * S; i += G) {
                const int rr = i / S, cc = i - rr * S;
                const int iv = v0 + rr - OCW, iu = u0 + cc - OCW;
                sa[i] = (iu >= 0 && iu < a.W && iv >= 0 && iv < a.H) ? __ldg(&a.ref[(size_t)iv * a.W + iu]) : 0.0f;
            }
        }
        if (t < 32) {
            const unsigned long long ss = q_ss, s = q_s;
            const unsigned int nul = q_nul;
            const int area = q_area, full = q_full;
            const int cnt = (int)nul + (full - area);   // pixels < 1e-10, zero fill included
            const int cnt_ref = __shfl_sync(0xffffffffu, cnt, 0), cnt_sa = __shfl_sync(0xffffffffu, cnt, 1);
            if (lane == 0) {
                ctl.valid = !(((float)cnt_ref / (float)(S * S) > 0.8f) || ((float)cnt_sa / (float)(Dx2 * Dy2) > 0.8f));
                ctl.chip_ss[0] = ss; ctl.chip_s[0] = s;
                ctl.chip_fast = (cnt == 0);
            }
            if (lane >= 2 && lane <= 4) { ctl.chip_ss[lane - 1] = ss; ctl.chip_s[lane - 1] = s; }
        }
        if (!C::DIRECT_CHIP) {
            gsync<G>();
            if (!ctl.valid) {
                if (t == 0) {
                    a.dp[3 * (size_t)g] = CUDART_NAN_F; a.dp[3 * (size_t)g + 1] = CUDART_NAN_F; a.dp[3 * (size_t)g + 2] = -3.0f;
                    if (a.peak) a.peak[g] = make_int2(0, 0);
                    if (a.ncell) a.ncell[g] = 0;
                }
                continue;
            }
#pragma unroll
            for (int rb = 0; rb < C::RB; rb++)
#pragma unroll
                for (int c = 0; c < L; c++) chip[rb][c] = (c < len && r + rb * C::RPT < S) ? sa[(r + rb * C::RPT) * S + col0 + c] : 0.0f;
            gsync<G>();
        }

        // ---- stage the search area (extract_sarea :857-890): zero outside the image, zero in the
        //      never-written last row / column, zero in the pad columns [Dx2, pitch) ----------------
        if (su0 - dx2 >= 0 && sv0 - dy2 >= 0 && su0 - dx2 + Dx2 - 1 <= a.W && sv0 - dy2 + Dy2 - 1 <= a.H) {
            // written part entirely inside the image (the common case): no per-pixel bounds tests
            const size_t first = (size_t)(sv0 - dy2) * a.W + (su0 - dx2);
            if (C::ASYNC_AREA)
                stage_rows_async<G>(a.srch + first, a.W, sa, pitch, Dy2 - 1, Dx2 - 1, t);   // waited for below, after the table reset
            else if (C::VEC && vec_ok && first >= 3 && first + (size_t)(Dy2 - 2) * a.W + Dx2 + 2 <= (size_t)a.H * a.W)
                stage_rows_v4<G, (G >= 128 ? 2 : 4)>(a.srch + first, a.W, sa, pitch, Dy2 - 1, Dx2 - 1, t);   // the chip pixels are live here: fewer loads in flight
            else
                stage_rows<G, 8>(a.srch + first, a.W, sa, pitch, Dy2 - 1, Dx2 - 1, t);
            for (int x = t; x < pitch; x += G) sa[(Dy2 - 1) * pitch + x] = 0.0f;
        } else {
            for (int y = gwarp; y < Dy2; y += C::NWARPS) {
                const int iv = sv0 - dy2 + y;
                const bool rowok = (y < Dy2 - 1) && iv >= 0 && iv < a.H;
                const float *src = a.srch + (size_t)(rowok ? iv : 0) * a.W;
                for (int x = lane; x < pitch; x += 32) {
                    const int iu = su0 - dx2 + x;
                    sa[y * pitch + x] = (rowok && x < Dx2 - 1 && iu >= 0 && iu < a.W) ? __ldg(&src[iu]) : 0.0f;
                }
            }
        }
        // nothing known, nothing visible, no probe complete (cflag and ctab are adjacent: one sweep of words)
        for (int i = t; i < cw * ch; i += G) cval[i] = __uint_as_float(kUnknownBits);
        for (int i = t; i < cell_elems / 2; i += G) ((unsigned int *)cflag)[i] = 0u;
        if (t == 0) {   // written BEFORE the barrier: the leader warp reads it right after
            ctl.geo = {g, P, su0, sv0, dx2, dy2, Dx2, Dy2, cw, ch, sa_elems, cell_elems};
            ctl.m = 0; ctl.nblk = 0;
        }
        if (C::ASYNC_AREA) cp_async_wait_all();
        gsync<G>();
        if (!ctl.valid) {   // investigate_valid_grid: the search area was staged for nothing (rare)
            if (t == 0) {
                a.dp[3 * (size_t)g] = CUDART_NAN_F; a.dp[3 * (size_t)g + 1] = CUDART_NAN_F; a.dp[3 * (size_t)g + 2] = -3.0f;
                if (a.peak) a.peak[g] = make_int2(0, 0);
                if (a.ncell) a.ncell[g] = 0;
            }
            continue;
        }
        PROF_T(t_stage1);
        PROF_ADD(1, t_stage1 - t_node0);
        // this thread's view of the tile: row r, first column col0 (threads without chip pixels: the origin)
        node_rounds<OCW, G, EXACTP>(a, ctl, sa, sa + (active ? r * pitch + col0 : 0), chip, pitch,
                                    (C::RB > 1 && active && r + C::RPT < S) ? C::RPT * pitch : 0, active, t, lane, gwarp, lead);
        PROF_T(t_node1);
        PROF_ADD(0, t_node1 - t_node0);
        PROF_ADD(6, 1);
    }
}

float min_dn_float() {
    float f = (float)1e-10;
    if ((double)f < 1e-10) f = nextafterf(f, 1.0f);
    return f;
}

template <int OCW, int G, bool EXACTP, int MINCTA = min_ctas(OCW, G)>
int launch_one(mimc3cu_ctx *ctx, Match2Args &a, int groups_per_cta, size_t smem, int n_list, long long *grid_out) {
    auto kern = match2_kernel<OCW, G, EXACTP, MINCTA>;
    const int threads = G * groups_per_cta;   // <= kThreads; the kernel only needs whole groups
    CU_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return mimc3cu_fail(ctx, "match2: kernel does not fit on an SM (smem %zu)", smem);
    long long grid = (long long)per_sm * ctx->num_sms;
    const long long groups = ((long long)n_list + groups_per_cta - 1) / groups_per_cta;
    if (grid > groups) grid = groups;
    if (grid_out) *grid_out = grid;
    kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

inline int pivot_max(int G) { return G == 32 ? 64 : 128; }   // Cfg::PMAX

// Shared-memory bins of a launch: bin k runs `groups` node groups per CTA and is sized so that
// `ctas` CTAs fit on an SM.  A node goes to the first bin it fits; the warp-per-node kernels get
// two extra bins with few warps per CTA so that wide search areas (fast glaciers) stay on this
// kernel instead of falling back to the general FP64 one.
struct BinCfg { int G, groups, ctas; };
constexpr int kMaxBins = 5;
inline int bin_table(int ocw, BinCfg *t) {
    // ocw 30: half a CTA per node (chip rows in two 31-pixel segments), 6 or 4 nodes per SM at <= 85 registers
    // (measured: 8 nodes per SM at 64 registers spill and lose 20 %), then whole CTAs for the wide search areas
    if (ocw == 30) { t[0] = {128, 2, 3}; t[1] = {128, 2, 2}; t[2] = {256, 1, 2}; t[3] = {256, 1, 1}; return 4; }
    if (ocw >= 30) { t[0] = {256, 1, 4}; t[1] = {128, 1, 3}; t[2] = {256, 1, 2}; t[3] = {256, 1, 1}; return 4; }
    // small chips: a warp per node, eight nodes per CTA, as many CTAs per SM as the nodes' search areas allow; nodes with
    // very wide search areas (fast ice) would leave the SM with one or two warps that way, so they get a whole
    // 256-thread CTA each (row segments of <= 4 pixels per thread): more instructions per cell, but
    // many more resident warps (measured: +10 % on the fast-glacier workload, a loss for moderate areas)
    if (ocw == 7) { t[0] = {32, 8, 4}; t[1] = {32, 8, 3}; t[2] = {32, 8, 2}; t[3] = {256, 1, 3}; t[4] = {256, 1, 1}; }
    // half-width 15: between two CTAs of eight nodes and one CTA of eight nodes, two CTAs of five nodes (ten warps per SM)
    else { t[0] = {32, 8, 2}; t[1] = {32, 5, 2}; t[2] = {32, 8, 1}; t[3] = {256, 1, 3}; t[4] = {256, 1, 1}; }
    return 5;
}
// Development aid: MIMC3CU_BINS="40:0:128,1,5;30:1:256,1,3" replaces bin <index> of the named chip half-widths
// (only combinations with a compiled instantiation: see the dispatch in launch_match2).
inline int bin_table_env(int ocw, BinCfg *t) {
    const int nb = bin_table(ocw, t);
    if (const char *e = getenv("MIMC3CU_BINS")) {
        for (const char *p = e; p && *p;) {
            int o = 0, idx = 0, G = 0, groups = 0, ctas = 0;
            if (sscanf(p, "%d:%d:%d,%d,%d", &o, &idx, &G, &groups, &ctas) == 5 && o == ocw && idx >= 0 && idx < nb) t[idx] = {G, groups, ctas};
            p = strchr(p, ';');
            if (p) p++;
        }
    }
    return nb;
}

}  // namespace

bool match2_supported(const MatchLaunch &L, const Image *ref, const Image *srch) {
    if (!L.csr_off || L.chips) return false;                       // explicit (CP) mode stays on the general kernel
    if (!(L.ocw == 7 || L.ocw == 15 || L.ocw == 30 || L.ocw == 40)) return false;
    if (!ref->exact_class || !srch->exact_class || !ref->sat_valid || !srch->sat_valid) return false;
    // exactness budget of the FP32 accumulation (see the header): product bits + fraction bits <= 37
    const double maxprod = (double)ref->max_value * (double)srch->max_value;
    int b = 0;
    while (ldexp(1.0, b) < maxprod) b++;
    return b + ref->frac_bits + srch->frac_bits <= 37;
}

template <int OCW, int G>
static size_t static_smem_of() {
    cudaFuncAttributes fa;
    return cudaFuncGetAttributes(&fa, match2_kernel<OCW, G, false>) == cudaSuccess ? fa.sharedSizeBytes : 16384;
}
static size_t static_smem_bytes(int ocw, int G) {
    switch (ocw * 1000 + G) {
        case 7032: return static_smem_of<7, 32>();
        case 7256: return static_smem_of<7, 256>();
        case 15032: return static_smem_of<15, 32>();
        case 15256: return static_smem_of<15, 256>();
        case 30064: return static_smem_of<30, 64>();
        case 30128: return static_smem_of<30, 128>();
        case 30256: return static_smem_of<30, 256>();
        case 40128: return static_smem_of<40, 128>();
        case 40256: return static_smem_of<40, 256>();
    }
    return 16384;
}

static int build_bins(mimc3cu_ctx *ctx, PivotSet *ps, PivotSet::Bins &B, int ocw) {
    BinCfg tab[kMaxBins];
    const int nb = bin_table_env(ocw, tab);
    const size_t usable = ctx->smem_optin;
    for (int k = 0; k < nb; k++) {
        const size_t fixed = static_smem_bytes(ocw, tab[k].G) + 256;         // static control blocks + slack
        size_t per_cta = (228 * 1024 - tab[k].ctas * 1024) / tab[k].ctas;   // 1 KB reserved per resident CTA
        per_cta = std::min(per_cta, usable) - fixed;
        B.grp_bytes[k] = (int64_t)((per_cta / tab[k].groups) & ~(size_t)15);
    }
    // two passes (count, then fill) over the host copy of the last pivots
    std::vector<uint8_t> which((size_t)ps->n);
    B.max_global_cells = 0;
    int64_t cnt[kMaxBins + 1] = {0, 0, 0, 0, 0, 0};
    for (int32_t g = 0; g < ps->n; g++) {
        const int64_t Dx2 = 2 * (ps->last_u[g] + ocw + 2) + 1, Dy2 = 2 * (ps->last_v[g] + ocw + 2) + 1;
        const int64_t pitch = (Dx2 + 3) | 1, need_sa = (Dy2 * pitch + 3) & ~3LL;
        const int64_t need_cells = ((Dx2 - 2 * ocw - 1) * (Dy2 - 2 * ocw - 1) + 15) & ~15LL;
        const int64_t need = need_sa * 4 + need_cells * 6;   // same formula as the kernel's fit test
        const int P = std::max(ps->last_u[g], ps->last_v[g]) + 1;   // get_uv_pivot steps by one pixel along the dominant axis
        int k = 0;
        while (k < nb && (need > B.grp_bytes[k] || P > pivot_max(tab[k].G))) k++;
        if (k == nb) k = kMaxBins;   // general kernel
        which[g] = (uint8_t)k; cnt[k]++;
    }
    std::vector<int32_t> all((size_t)ps->n);
    int64_t pos[kMaxBins + 1];
    for (int k = 0, acc = 0; k <= kMaxBins; k++) { B.start[k] = acc; B.count[k] = (int32_t)cnt[k]; pos[k] = acc; acc += (int)cnt[k]; }
    for (int32_t g = 0; g < ps->n; g++) all[(size_t)pos[which[g]]++] = g;
    if ((size_t)ps->n > B.lists_cap) {
        if (B.lists) { CU_CHECK(ctx, cudaFree(B.lists)); B.lists = nullptr; B.lists_cap = 0; }
        CU_CHECK(ctx, cudaMalloc(&B.lists, sizeof(int32_t) * (size_t)ps->n));
        B.lists_cap = (size_t)ps->n;
    }
    // ordered against the matcher launches on the context stream (upload_sync in api.cu explains why not cudaMemcpy)
    if (int rc = upload_sync(ctx, B.lists, all.data(), sizeof(int32_t) * all.size())) return rc;
    B.ocw = ocw;
    if (getenv("MIMC3CU_DEBUG_BINS")) {
        fprintf(stderr, "[mimc3cu] ocw %d bins:", ocw);
        for (int k = 0; k < nb; k++) fprintf(stderr, " {G %d x %d, %d CTA/SM, %lld B/node}: %lld", tab[k].G, tab[k].groups, tab[k].ctas, (long long)B.grp_bytes[k], (long long)cnt[k]);
        fprintf(stderr, "  general kernel: %lld of %d nodes\n", (long long)cnt[kMaxBins], ps->n);
    }
    return 0;
}

int launch_match2(mimc3cu_ctx *ctx, const MatchLaunch &L, const Image *ref, const Image *srch, PivotSet *ps) {
    if (L.n <= 0) return 0;
    // bins cached per (pivot set, ocw); two cache entries cover main's usage (one ocw per slot)
    PivotSet::Bins *B = nullptr;
    for (auto &b : ps->bins) if (b.ocw == L.ocw) B = &b;
    if (!B) {
        B = ps->bins[0].ocw < 0 ? &ps->bins[0] : &ps->bins[1];
        // no launch that reads this slot's node lists can be in flight: set_pivots waited for them, and a second
        // chip size on the same slot uses the other cache entry (a third one recycles: wait for the slot then)
        if (B->ocw >= 0 && ps->last_use) CU_CHECK(ctx, cudaEventSynchronize(ps->last_use));
        if (int rc = build_bins(ctx, ps, *B, L.ocw)) return rc;
    }
    if (ctx->overflow_cap < (size_t)L.n) {
        CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->overflow_list) CU_CHECK(ctx, cudaFree(ctx->overflow_list));
        CU_CHECK(ctx, cudaMalloc(&ctx->overflow_list, sizeof(int) * (size_t)L.n));
        ctx->overflow_cap = (size_t)L.n;
    }

    Match2Args a;
    a.ref = L.ref; a.srch = L.srch; a.H = L.H; a.W = L.W;
    a.sat_ref = (const ulonglong2 *)ref->sat; a.sat_srch = (const ulonglong2 *)srch->sat;
    a.inv_ref = ldexp(1.0, -ref->frac_bits); a.inv_ref2 = ldexp(1.0, -2 * ref->frac_bits);
    a.inv_srch = ldexp(1.0, -srch->frac_bits); a.inv_srch2 = ldexp(1.0, -2 * srch->frac_bits);
    a.node_uv = L.node_uv; a.off_u = L.off_u; a.off_v = L.off_v; a.csr_off = L.csr_off; a.piv = (const int2 *)L.piv;
    a.sign = L.sign; a.negate = L.negate; a.dp = L.dp; a.peak = (int2 *)L.peak; a.ncell = L.ncell;
    a.min_dn = min_dn_float();
    a.num_sms = std::max(1, ctx->num_sms);
    // accumulator biases: A0 = 2^(b+4) with 2^b >= max product; lo unit = product granularity
    const double maxprod = std::max(1.0, (double)ref->max_value * (double)srch->max_value);
    int b = 0;
    while (ldexp(1.0, b) < maxprod) b++;
    a.A0 = (float)ldexp(1.0, b + 4);
    a.lo_unit = ldexp(1.0, -(ref->frac_bits + srch->frac_bits));
    a.Mlo = (float)(ldexp(1.5, 23) * a.lo_unit);
    a.hi_unit = ldexp(1.0, b + 4 - 23);
    memcpy(&a.A0_bits, &a.A0, 4); memcpy(&a.Mlo_bits, &a.Mlo, 4);

    // counters: [0] general kernel, [1..3] the v2 bins, [8] overflow count (seeded with bin 3)
    CU_CHECK(ctx, cudaMemsetAsync(ctx->counter, 0, 16 * sizeof(unsigned int), ctx->stream));
    a.overflow_list = ctx->overflow_list; a.overflow_count = ctx->counter + 8;
    if (B->count[kMaxBins] > 0) {
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->overflow_list, B->lists + B->start[kMaxBins], sizeof(int) * (size_t)B->count[kMaxBins],
                                      cudaMemcpyDeviceToDevice, ctx->stream));
        unsigned int c3 = (unsigned int)B->count[kMaxBins];
        // small H2D from the stack is safe: the value is copied at enqueue time for pageable memory
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->counter + 8, &c3, sizeof(c3), cudaMemcpyHostToDevice, ctx->stream));
    }
    BinCfg tab[kMaxBins];
    const int nb = bin_table_env(L.ocw, tab);
    // products of the scaled operands below 2^24 are exact in FP32: the FFMA form of the inner loop applies
    const bool exactp = b + ref->frac_bits + srch->frac_bits <= 24;
    for (int k = 0; k < nb; k++) {
        if (B->count[k] == 0) continue;
        a.node_list = B->lists + B->start[k]; a.n_list = B->count[k];
        a.counter = ctx->counter + 1 + k;
        a.grp_bytes = (int)B->grp_bytes[k];
        a.ngroups = tab[k].groups;
        const size_t smem = (size_t)a.grp_bytes * tab[k].groups;
        int rc = 0;
        switch (L.ocw) {
            case 7:
                if (tab[k].G == 256) rc = exactp ? launch_one<7, 256, true>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<7, 256, false>(ctx, a, 1, smem, a.n_list, nullptr);
                else rc = exactp ? launch_one<7, 32, true>(ctx, a, tab[k].groups, smem, a.n_list, nullptr) : launch_one<7, 32, false>(ctx, a, tab[k].groups, smem, a.n_list, nullptr);
                break;
            case 15:
                if (tab[k].G == 256) rc = exactp ? launch_one<15, 256, true>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<15, 256, false>(ctx, a, 1, smem, a.n_list, nullptr);
                else rc = exactp ? launch_one<15, 32, true>(ctx, a, tab[k].groups, smem, a.n_list, nullptr) : launch_one<15, 32, false>(ctx, a, tab[k].groups, smem, a.n_list, nullptr);
                break;
            case 30:
                if (tab[k].G == 256) rc = exactp ? launch_one<30, 256, true>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<30, 256, false>(ctx, a, 1, smem, a.n_list, nullptr);
                else if (tab[k].G == 64) rc = exactp ? launch_one<30, 64, true>(ctx, a, tab[k].groups, smem, a.n_list, nullptr) : launch_one<30, 64, false>(ctx, a, tab[k].groups, smem, a.n_list, nullptr);
                else rc = exactp ? launch_one<30, 128, true>(ctx, a, tab[k].groups, smem, a.n_list, nullptr) : launch_one<30, 128, false>(ctx, a, tab[k].groups, smem, a.n_list, nullptr);
                break;
            case 40:
                if (tab[k].G == 128 && tab[k].ctas >= 5) rc = exactp ? launch_one<40, 128, true, 5>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<40, 128, false, 5>(ctx, a, 1, smem, a.n_list, nullptr);
                else if (tab[k].G == 128) rc = exactp ? launch_one<40, 128, true>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<40, 128, false>(ctx, a, 1, smem, a.n_list, nullptr);
                else if (tab[k].ctas <= 2) rc = exactp ? launch_one<40, 256, true, 2>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<40, 256, false, 2>(ctx, a, 1, smem, a.n_list, nullptr);
                else if (tab[k].ctas == 3) rc = exactp ? launch_one<40, 256, true, 3>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<40, 256, false, 3>(ctx, a, 1, smem, a.n_list, nullptr);
                else rc = exactp ? launch_one<40, 256, true>(ctx, a, 1, smem, a.n_list, nullptr) : launch_one<40, 256, false>(ctx, a, 1, smem, a.n_list, nullptr);
                break;
            default: return mimc3cu_fail(ctx, "match2: unsupported ocw %d", L.ocw);
        }
        if (rc) return rc;
    }
#ifdef MIMC3CU_PROFILE
    {
        unsigned long long h[24];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
        if (h[6]) {
            double tot = 0;
            for (int k = 11; k < 20; k++) tot += (double)h[k];
            fprintf(stderr, "[prof ocw %d] streams of the SM inside a fast round, seen mid-round:", L.ocw);
            for (int k = 11; k < 20; k++) if (h[k]) fprintf(stderr, " %d: %.0f%%", k - 11, 100.0 * h[k] / tot);
            fprintf(stderr, "\n");
        }
        if (h[6])
            fprintf(stderr, "[prof ocw %d] nodes %llu rounds/node %.2f  cycles/node: total %.0f staging %.0f leader %.0f (walk %.0f requests %.0f replay %.0f) "
                            "compute %.0f finalize %.0f table %.0f\n",
                    L.ocw, h[6], (double)h[5] / h[6], (double)h[0] / h[6], (double)h[1] / h[6], (double)h[2] / h[6], (double)h[7] / h[6],
                    (double)h[8] / h[6], (double)h[9] / h[6], (double)h[3] / h[6], (double)h[4] / h[6], (double)h[10] / h[6]);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_prof, h, sizeof(h));
    }
#endif
    // whatever did not fit goes to the general kernel (device-side count)
    MatchLaunch L1 = L;
    L1.node_list = ctx->overflow_list; L1.list_count = ctx->counter + 8; L1.list_n = 0;
    if (B->count[kMaxBins] == 0) {
        // nothing pre-seeded and the v2 kernels only overflow if host and device sizing disagree;
        // still run a minimal grid so that such nodes are never dropped
        L1.n = std::min<int32_t>(L.n, ctx->num_sms);
    }
    return launch_match(ctx, L1);
}
