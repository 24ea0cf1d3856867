// DLC-constrained, null-excluding NCC matcher for sm_100a.
//
// Replaces matching_ncc_dlc_2 / extract_refchip / extract_sarea / investigate_valid_grid /
// find_ncc_peak of the reference (MIMC_module.c:605-890).  One persistent CTA per SM slot
// pulls grid nodes from a global counter; for each node it stages the reference chip and
// the search area in shared memory, evaluates NCC cells cooperatively, and runs the
// reference's hill-climbing state machine verbatim so that the set of evaluated cells,
// the first-wins tie-breaks and the -2.0 placeholders of the sub-pixel fit are identical.
//
// Exactness: the five sums of a cell are accumulated from float products
// (__fmul_rn, like the reference's `float*float`) into FP64; for every image the
// reference can load (integer DN, or multiples of 1/8 after the Laplacian filter) all
// partial sums are exactly representable, so the result does not depend on summation
// order and the cell value is bit-identical to the CPU's sequential loop.  The final
// normalisation and the 3x3 quadratic fit use explicitly rounded FP64/FP32 intrinsics
// (no FMA contraction).
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned char kComputed = 1, kVisible = 2;

struct MatchArgs {
    const float *ref, *srch;
    int H, W;
    const int2 *node_uv;
    int off_u, off_v;
    const int *csr_off;
    const int2 *piv;
    int sign;
    const float *chips, *sareas;
    int D, P;
    long long chip_stride;   // explicit mode: floats between consecutive problems' chips
    int chip_pitch;          // ... and between chip rows (chips may be windows of larger tiles)
    int n, ocw;
    float negate;
    float *dp;
    int2 *peak;
    int *ncell;
    float *scr_val;
    int *scr_list;
    unsigned char *scr_flag;
    long long scr_stride;
    unsigned int *counter;
    const int *node_list;            // optional indirection (NULL: nodes 0..n-1)
    const unsigned int *list_count;  // optional device-side length of node_list
    int list_n;
    int sa_cap;      // floats of shared memory available for the search area
    float min_dn;    // smallest float >= 1e-10 (the reference compares against a double literal)
};

struct Sums {
    double sx, sy, sxx, syy, sxy;
    int n;
};

// Per-node geometry, identical for all threads of the CTA.
struct Node {
    int u0, v0;        // chip centre in the reference image
    int su0, sv0;      // search-area centre in the search image (u0 + offset)
    int P;             // number of pivots
    const int2 *piv;   // pivot list (global)
    int sign;
    int dx2, dy2, Dx2, Dy2;
    int cw, ch;        // reachable cmap region (cells whose window fits), see below
    bool staged;       // search area lives in shared memory
    const float *chip_src;   // explicit mode
    const float *sa_src;     // explicit mode
};

__device__ __forceinline__ float sarea_global(const MatchArgs &a, const Node &nd, int y, int x) {
    if (nd.sa_src) return nd.sa_src[(size_t)y * nd.Dx2 + x];
    // extract_sarea (MIMC_module.c:857-890): loops run to `< dx2`, so the last row and
    // column are never written (zero under the zero-initialised-allocation semantics, H1).
    if (y >= nd.Dy2 - 1 || x >= nd.Dx2 - 1) return 0.0f;
    int iv = nd.sv0 - nd.dy2 + y, iu = nd.su0 - nd.dx2 + x;
    if (iu < 0 || iu >= a.W || iv < 0 || iv >= a.H) return 0.0f;
    return __ldg(&a.srch[(size_t)iv * a.W + iu]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
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

// Rows [r0, r1) of one NCC cell whose window centre is (px, py) in search-area coordinates.
template <bool STAGED>
__device__ __forceinline__ Sums cell_rows(const MatchArgs &a, const Node &nd, const float *chip, const float *sa,
                                          int px, int py, int r0, int r1, int lane) {
    const int ocw = a.ocw, S = 2 * ocw + 1;
    Sums s = {0.0, 0.0, 0.0, 0.0, 0.0, 0};
    const int y0 = py - ocw, x0 = px - ocw;
    for (int r = r0; r < r1; r++) {
        for (int c = lane; c < S; c += 32) {
            float rv = chip[r * S + c];
            float sv = STAGED ? sa[(y0 + r) * nd.Dx2 + x0 + c] : sarea_global(a, nd, y0 + r, x0 + c);
            if (rv >= a.min_dn && sv >= a.min_dn) {   // null exclusion, MIMC_module.c:723
                s.n++;
                s.sx += (double)rv;
                s.sy += (double)sv;
                s.sxx += (double)__fmul_rn(rv, rv);
                s.syy += (double)__fmul_rn(sv, sv);
                s.sxy += (double)__fmul_rn(rv, sv);
            }
        }
    }
    s.n = warp_sum(s.n);
    s.sx = warp_sum(s.sx); s.sy = warp_sum(s.sy);
    s.sxx = warp_sum(s.sxx); s.syy = warp_sum(s.syy); s.sxy = warp_sum(s.sxy);
    return s;
}

struct Shared {
    Sums part[kWarps];
    int job[9];
    int m;                // >0: cells requested, 0: none, -1: done
    int cnt_ref, cnt_sa;  // null counts for investigate_valid_grid
    int list_len;
    unsigned int node;
};

// Evaluate `m` cells (linear indices into the cmap region) cooperatively. Block-uniform.
template <bool STAGED>
__device__ void eval_cells(const MatchArgs &a, const Node &nd, const float *chip, const float *sa, Shared &sh,
                           const int *list, int m, float *cval, unsigned char *cflag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = 2 * a.ocw + 1;
    if (m <= 0) return;   // block-uniform
    int split = (m >= kWarps) ? 1 : kWarps / m;
    if (split > S) split = S;
    const int items = m * split;
    for (int it = warp; it < items; it += kWarps) {
        int c = it / split, part = it - c * split;
        int cell = list[c];
        int cy = cell / nd.cw, cx = cell - cy * nd.cw;
        int px = cx + a.ocw + 1, py = cy + a.ocw + 1;
        int r0 = (part * S) / split, r1 = ((part + 1) * S) / split;
        Sums s = cell_rows<STAGED>(a, nd, chip, sa, px, py, r0, r1, lane);
        if (lane == 0) {
            if (split == 1) {
                cval[cell] = ncc_from_sums(s);
                cflag[cell] |= kComputed;
            } else {
                sh.part[it] = s;
            }
        }
    }
    if (split > 1) {
        __syncthreads();
        if (tid < m) {
            Sums s = sh.part[tid * split];
            for (int k = 1; k < split; k++) {
                const Sums &q = sh.part[tid * split + k];
                s.n += q.n; s.sx += q.sx; s.sy += q.sy; s.sxx += q.sxx; s.syy += q.syy; s.sxy += q.sxy;
            }
            int cell = list[tid];
            cval[cell] = ncc_from_sums(s);
            cflag[cell] |= kComputed;
        }
    }
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

__global__ void __launch_bounds__(kThreads) match_kernel(const MatchArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ Shared sh;
    const int tid = threadIdx.x;
    const int ocw = a.ocw, S = 2 * ocw + 1;
    float *chip = smem;
    float *sa = smem + S * S;
    float *cval = a.scr_val + (size_t)blockIdx.x * a.scr_stride;
    int *clist = a.scr_list + (size_t)blockIdx.x * a.scr_stride;
    unsigned char *cflag = a.scr_flag + (size_t)blockIdx.x * a.scr_stride;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sh.node = atomicAdd(a.counter, 1u);
            sh.cnt_ref = 0; sh.cnt_sa = 0; sh.list_len = 0; sh.m = 0;
        }
        __syncthreads();
        unsigned int g = sh.node;
        if (a.node_list) {
            const unsigned int lim = a.list_count ? *a.list_count : (unsigned int)a.list_n;
            if (g >= lim) break;
            g = (unsigned int)a.node_list[g];
        } else if (g >= (unsigned int)a.n) break;

        // ---- node geometry ---------------------------------------------------------
        Node nd;
        nd.sign = a.sign;
        if (a.csr_off) {
            int b = a.csr_off[g];
            nd.P = a.csr_off[g + 1] - b;
            nd.piv = a.piv + b;
            int2 uv = a.node_uv[g];
            nd.u0 = uv.x; nd.v0 = uv.y;
            nd.su0 = uv.x + a.off_u; nd.sv0 = uv.y + a.off_v;
            nd.chip_src = nullptr; nd.sa_src = nullptr;
        } else {
            nd.P = a.P; nd.piv = a.piv;
            nd.u0 = nd.v0 = nd.su0 = nd.sv0 = 0;
            nd.chip_src = a.chips + (size_t)g * a.chip_stride;
            nd.sa_src = a.sareas + (size_t)g * a.D * a.D;
        }
        if (nd.P <= 0) {   // undefined behaviour in the reference (MIMC_module.c:589-591)
            if (tid == 0) {
                a.dp[3 * (size_t)g] = CUDART_NAN_F; a.dp[3 * (size_t)g + 1] = CUDART_NAN_F; a.dp[3 * (size_t)g + 2] = -2.0f;
                if (a.peak) a.peak[g] = make_int2(0, 0);
                if (a.ncell) a.ncell[g] = 0;
            }
            continue;
        }
        if (nd.sa_src) {
            nd.Dx2 = a.D; nd.Dy2 = a.D; nd.dx2 = a.D / 2; nd.dy2 = a.D / 2;
        } else {
            int2 last = nd.piv[nd.P - 1];
            nd.dx2 = abs(last.x) + ocw + 2; nd.dy2 = abs(last.y) + ocw + 2;   // :863-866
            nd.Dx2 = 2 * nd.dx2 + 1; nd.Dy2 = 2 * nd.dy2 + 1;
        }
        // cells whose 3x3 probe can ever be requested: window centre in [ocw+1, D-ocw-1]
        nd.cw = nd.Dx2 - 2 * ocw - 1; nd.ch = nd.Dy2 - 2 * ocw - 1;
        const int ncells_region = nd.cw * nd.ch;
        const int sa_elems = nd.Dx2 * nd.Dy2;
        nd.staged = sa_elems <= a.sa_cap;

        // ---- stage chip + search area, count null pixels ---------------------------
        int inv_ref = 0, inv_sa = 0;
        for (int i = tid; i < S * S; i += kThreads) {
            float v;
            int r = i / S, c = i - r * S;
            if (nd.chip_src) v = nd.chip_src[r * a.chip_pitch + c];
            else {
                int iv = nd.v0 + r - ocw, iu = nd.u0 + c - ocw;   // extract_refchip :845-855 (no bounds check there)
                v = (iu >= 0 && iu < a.W && iv >= 0 && iv < a.H) ? __ldg(&a.ref[(size_t)iv * a.W + iu]) : 0.0f;
            }
            chip[i] = v;
            inv_ref += (v < a.min_dn);
        }
        for (int i = tid; i < sa_elems; i += kThreads) {
            int y = i / nd.Dx2, x = i - y * nd.Dx2;
            float v = sarea_global(a, nd, y, x);
            if (nd.staged) sa[i] = v;
            inv_sa += (v < a.min_dn);
        }
        for (int i = tid; i < ncells_region; i += kThreads) cflag[i] = 0;
        inv_ref = warp_sum(inv_ref); inv_sa = warp_sum(inv_sa);
        if ((tid & 31) == 0) { atomicAdd(&sh.cnt_ref, inv_ref); atomicAdd(&sh.cnt_sa, inv_sa); }
        __syncthreads();

        // investigate_valid_grid :605-644
        const bool invalid = ((float)sh.cnt_ref / (float)(S * S) > 0.8f) || ((float)sh.cnt_sa / (float)sa_elems > 0.8f);
        if (invalid) {
            if (tid == 0) {
                a.dp[3 * (size_t)g] = CUDART_NAN_F; a.dp[3 * (size_t)g + 1] = CUDART_NAN_F; a.dp[3 * (size_t)g + 2] = -3.0f;
                if (a.peak) a.peak[g] = make_int2(0, 0);
                if (a.ncell) a.ncell[g] = 0;
            }
            continue;
        }

        // ---- up-front batch: the first 3x3 probe of every pivot is unconditional ------
        for (int i = tid; i < nd.P * 9; i += kThreads) {
            int ip = i / 9, k = i - ip * 9;
            int2 pv = nd.piv[ip];
            int px = nd.sign * pv.x + nd.dx2, py = nd.sign * pv.y + nd.dy2;
            if (px - ocw <= 1 || px + ocw >= nd.Dx2 - 1 || py - ocw <= 1 || py + ocw >= nd.Dy2 - 1) continue;
            int cx = px + (k / 3 - 1) - (ocw + 1), cy = py + (k % 3 - 1) - (ocw + 1);
            int cell = cy * nd.cw + cx;
            // claim the cell once: byte-wide flags, so use a CAS on the containing word
            unsigned int *w = (unsigned int *)(cflag + (cell & ~3));
            unsigned int bit = (unsigned int)4u << (8 * (cell & 3));   // temporary "listed" bit (value 4)
            unsigned int old = atomicOr(w, bit);
            if (!(old & bit)) clist[atomicAdd(&sh.list_len, 1)] = cell;
        }
        __syncthreads();
        if (nd.staged) eval_cells<true>(a, nd, chip, sa, sh, clist, sh.list_len, cval, cflag);
        else eval_cells<false>(a, nd, chip, sa, sh, clist, sh.list_len, cval, cflag);
        __syncthreads();

        // ---- the reference's hill-climbing state machine (thread 0), lazy cells ---------
        // registers of thread 0 only
        int ip = 0, px = 0, py = 0, duv0 = -1, duv1 = -1, flag_new = 1;
        int peak_x = nd.dx2, peak_y = nd.dy2, ncells = 0;
        float nccmax = -2.0f, best = -2.0f;
        bool in_pivot = false;
        for (;;) {
            if (tid == 0) {
                int m = -1;
                while (true) {
                    if (!in_pivot) {
                        if (ip >= nd.P) { m = -1; break; }
                        int2 pv = nd.piv[ip];
                        px = nd.sign * pv.x + nd.dx2; py = nd.sign * pv.y + nd.dy2;   // :693-694
                        nccmax = -2.0f; duv0 = -1; duv1 = -1; flag_new = 1;
                        in_pivot = true;
                    }
                    bool stop = !((duv0 != 0 || duv1 != 0) && flag_new != 0);       // :699
                    if (!stop) {
                        if (px - ocw <= 1 || px + ocw >= nd.Dx2 - 1 || py - ocw <= 1 || py + ocw >= nd.Dy2 - 1) {
                            duv0 = 0; duv1 = 0; stop = true;                         // :701-707 (break)
                        }
                    }
                    if (stop) {
                        if (nccmax > best) { peak_x = px; peak_y = py; best = nccmax; }   // :747-752
                        in_pivot = false; ip++;
                        continue;
                    }
                    // probe the 3x3: first make sure every needed value exists
                    int need = 0;
                    for (int c1 = -1; c1 <= 1; c1++)
                        for (int c2 = -1; c2 <= 1; c2++) {
                            int cell = (py + c2 - ocw - 1) * nd.cw + (px + c1 - ocw - 1);
                            unsigned char f = cflag[cell];
                            if (!(f & kVisible) && !(f & kComputed)) sh.job[need++] = cell;
                        }
                    if (need) { m = need; break; }
                    duv0 = 0; duv1 = 0; flag_new = 0;
                    for (int c1 = -1; c1 <= 1; c1++)
                        for (int c2 = -1; c2 <= 1; c2++) {
                            int cell = (py + c2 - ocw - 1) * nd.cw + (px + c1 - ocw - 1);
                            float v = cval[cell];
                            bool vis = cflag[cell] & kVisible;
                            if (!vis || v < -1.0f) {            // `cmap < -1.0` => (re)evaluated, :713
                                flag_new++; ncells++;
                                cflag[cell] |= kVisible;
                            }
                            if (v > nccmax) { nccmax = v; duv0 = c1; duv1 = c2; }   // :736-741
                        }
                    px += duv0; py += duv1;                                           // :744-745
                }
                sh.m = m;
            }
            __syncthreads();
            const int m = sh.m;
            if (m < 0) break;
            if (nd.staged) eval_cells<true>(a, nd, chip, sa, sh, sh.job, m, cval, cflag);
            else eval_cells<false>(a, nd, chip, sa, sh, sh.job, m, cval, cflag);
            __syncthreads();
        }

        // ---- sub-pixel fit and output (thread 0) ------------------------------------------
        if (tid == 0) {
            float n9[9];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    int cx = peak_x - 1 + c - (ocw + 1), cy = peak_y - 1 + r - (ocw + 1);
                    float v = -2.0f;   // never evaluated (or outside the evaluable region)
                    if (cx >= 0 && cx < nd.cw && cy >= 0 && cy < nd.ch) {
                        int cell = cy * nd.cw + cx;
                        if (cflag[cell] & kVisible) v = cval[cell];
                    }
                    n9[r * 3 + c] = v;
                }
            float du, dv;
            subpixel_fit(n9, peak_x - nd.dx2, peak_y - nd.dy2, du, dv);
            a.dp[3 * (size_t)g] = a.negate * du;
            a.dp[3 * (size_t)g + 1] = a.negate * dv;
            a.dp[3 * (size_t)g + 2] = best;
            if (a.peak) a.peak[g] = make_int2(peak_x - nd.dx2, peak_y - nd.dy2);
            if (a.ncell) a.ncell[g] = ncells;
        }
    }
}

float min_dn_float() {
    // smallest float f with (double)f >= 1e-10, so that `f32 >= 1e-10 (double)` == `f32 >= f`
    float f = (float)1e-10;
    if ((double)f < 1e-10) f = nextafterf(f, 1.0f);
    return f;
}

}  // namespace

int launch_match(mimc3cu_ctx *ctx, const MatchLaunch &L) {
    if (L.n <= 0) return 0;
    const int S = 2 * L.ocw + 1;
    if (L.ocw < 1) return mimc3cu_fail(ctx, "match: ocw must be >= 1");
    const size_t chip_bytes = (size_t)S * S * sizeof(float);
    const size_t smem_max = ctx->smem_optin - 2048;   // static shared + slack
    if (chip_bytes + 1024 > smem_max) return mimc3cu_fail(ctx, "match: chip %dx%d does not fit shared memory", S, S);
    size_t want = chip_bytes + (size_t)L.max_sarea * sizeof(float);
    size_t smem = want <= smem_max ? want : smem_max;
    // keep at least 2 CTAs per SM resident when the search areas are small enough
    MatchArgs a;
    a.ref = L.ref; a.srch = L.srch; a.H = L.H; a.W = L.W; a.node_uv = L.node_uv;
    a.off_u = L.off_u; a.off_v = L.off_v; a.csr_off = L.csr_off; a.piv = (const int2 *)L.piv; a.sign = L.sign;
    a.chips = L.chips; a.sareas = L.sareas; a.D = L.D; a.P = L.P;
    a.chip_stride = L.chip_stride ? L.chip_stride : (long long)S * S;
    a.chip_pitch = L.chip_pitch ? L.chip_pitch : S;
    a.n = L.n; a.ocw = L.ocw; a.negate = L.negate; a.dp = L.dp; a.peak = (int2 *)L.peak; a.ncell = L.ncell;
    a.node_list = L.node_list; a.list_count = L.list_count; a.list_n = L.list_n;
    a.sa_cap = (int)((smem - chip_bytes) / sizeof(float));
    a.min_dn = min_dn_float();

    CU_CHECK(ctx, cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel, kThreads, smem));
    if (per_sm < 1) return mimc3cu_fail(ctx, "match: kernel does not fit on an SM (smem %zu)", smem);
    long long grid = (long long)per_sm * ctx->num_sms;
    const long long work = L.node_list ? (L.list_count ? (long long)L.n : (long long)L.list_n) : (long long)L.n;
    if (grid > work) grid = work;
    if (grid < 1) return 0;

    // per-CTA scratch: cmap values, flags, cell list
    long long stride = (L.max_cells + 15) & ~15LL;
    if (stride < 16) stride = 16;
    size_t need = (size_t)grid * stride * (sizeof(float) + sizeof(int) + 1) + 256;
    if (int rc = ensure_scratch(ctx, need)) return rc;
    a.scr_val = (float *)ctx->scratch;
    a.scr_list = (int *)(a.scr_val + (size_t)grid * stride);
    a.scr_flag = (unsigned char *)(a.scr_list + (size_t)grid * stride);
    a.scr_stride = stride;
    a.counter = ctx->counter;   // slot 0: the general kernel's node counter
    CU_CHECK(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    match_kernel<<<(unsigned)grid, kThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}
