// Postprocess kernels: multi-match clustering, prominent-cluster field, flow-direction
// guided hole filling (Jacobi sweeps), 3x3 smoothing, snap-to-cluster and pseudosmoothing.
//
// Replaces mimc2_postprocess and its callees (MIMC_module.c:893-2496).  Every stage is a
// per-node (thread-per-node) kernel over the dimy x dimx node grid; the iterative stages
// are Jacobi sweeps (reads of committed values only), so they parallelise without changing
// the result.  Sweep control (the reference's nested while loops) stays on the host and
// reads back one small counter block per sweep.
//
// Arithmetic follows the reference's float/double mix operation by operation (this TU is
// built with -fmad=false).  Two library calls can differ from glibc in the last ulp and are
// the only source of non-bit-equality: expf() in the dpf1 weights (:1514; evaluated through the
// double-precision exp here, which agrees with glibc's expf except for rare double-rounding
// cases) and exp() in the pseudosmoothing weights (:2157; CUDA 1 ulp vs glibc < 1 ulp).  Both
// only influence which cluster is nearest to an interpolated value, i.e. they matter on exact
// ties only.
#include <math_constants.h>

#include <vector>

#include "common.cuh"

namespace {
constexpr int kT = 128;
constexpr int MAXDP = 32;     // attempts per node supported by the clustering kernel
constexpr int MAXNB = 128;    // neighbour offsets (29 for radius 3, 81 for radius 5)

inline int nblocks(size_t n) { return (int)((n + kT - 1) / kT); }
}  // namespace

// Geometry of one band of node rows.  The local arrays cover the owned rows plus `halo` rows of
// the neighbouring bands on each side (clipped at the grid edges), so that "inside the local
// array" == "inside the global grid" for every neighbour offset a sweep can reach.
struct Band {
    int dimx;          // nodes per grid row
    int rows;          // local rows (owned + halo)
    int own0, own1;    // owned local rows [own0, own1)
    int grow0;         // global row index of local row 0
    int gdimy;         // rows of the global grid
};

struct Post {
    int32_t n = 0, K = 0, dimx = 0, dimy = 0;   // n = OWNED nodes of the last run
    int32_t own_off = 0;                        // first owned node in the local arrays
    Band band;
    float *mvn = nullptr;       // (n, K, 5)
    int32_t *ncl = nullptr;     // (n)
    int32_t *dpf0 = nullptr;    // after get_dpf0
    int32_t *id = nullptr;      // working cluster-id field
    float *dx = nullptr, *dy = nullptr, *dxb = nullptr, *dyb = nullptr, *noi = nullptr;
    int32_t *id1 = nullptr;     // after dpf1
    float *dx1 = nullptr, *dy1 = nullptr;
    int32_t *bid = nullptr;
    uint8_t *mask[2] = {nullptr, nullptr};
    uint8_t *stack = nullptr;   // (kMaxStack, n)
    int32_t *ruv = nullptr;     // (MAXNB, 2)
    int32_t *ctr = nullptr;     // device counters (256 ints)
    int32_t *list[2] = {nullptr, nullptr};   // work lists of the sweeps (local node indices): holes still to fill / dirty nodes
    double *apv = nullptr;      // (n, 2) a-priori vx, vy (xyuvav cols 4,5)
    size_t cap_n = 0;
    int32_t stack_cap = 0;
};

namespace {
constexpr int kMaxStack = 104;

// ---------------------------------------------------------------------------------------
// calc_mean_var_num_dp_cluster + cluster_euclidian + mark_row, MIMC_module.c:994-1194
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) cluster_kernel(const float *__restrict__ dp, int n, int K, float *__restrict__ mvn,
                                                     int *__restrict__ ncl) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    float sx[MAXDP], sy[MAXDP];
    unsigned adj[MAXDP];
    unsigned char lab[MAXDP];
    int k = 0;
    for (int a = 0; a < K; a++) {
        const float *d = dp + ((size_t)a * n + g) * 3;
        if (d[2] > 0.1f) { sx[k] = d[0]; sy[k] = d[1]; k++; }                  // :1048
    }
    const float min_dist_sq = 0.25f;
    for (int i = 0; i < k; i++) {
        unsigned m = 0;
        for (int j = 0; j < k; j++) {
            // the reference evaluates (j >= i ? v[j]-v[i] : v[i]-v[j]); squares are identical
            float ddx = __fsub_rn(sx[j], sx[i]), ddy = __fsub_rn(sy[j], sy[i]);
            float d2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
            if (d2 < min_dist_sq) m |= 1u << j;                                 // :1156
        }
        adj[i] = m;
        lab[i] = 0;
    }
    // mark_row's recursive DFS labels exactly the set reachable from row i through
    // still-unlabelled candidates; ids are handed out in order of the first member (:1170-1176).
    int id_curr = 0, max_id = 0;
    unsigned unl = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
    for (int i = 0; i < k; i++) {
        if (lab[i] != 0) continue;
        id_curr++;
        unsigned frontier = adj[i] & unl, comp = 0;
        while (frontier) {
            comp |= frontier; unl &= ~frontier;
            unsigned nxt = 0;
            for (unsigned f = frontier; f; f &= f - 1) nxt |= adj[__ffs(f) - 1];
            frontier = nxt & unl;
        }
        for (unsigned c = comp; c; c &= c - 1) lab[__ffs(c) - 1] = (unsigned char)id_curr;
        if (comp) max_id = id_curr;
        // a candidate with a NaN displacement never labels itself: it consumes an id and stays
        // 0; the reference then corrupts its stack, we leave an empty cluster.
    }
    float *o = mvn + (size_t)g * K * 5;
    for (int c = 1; c <= max_id; c++) {
        float ax = 0.f, ay = 0.f, axx = 0.f, ayy = 0.f;
        int ns = 0;
        for (int i = 0; i < k; i++)
            if (lab[i] == c) {                                                  // attempt order, float sums :1084-1090
                ax = __fadd_rn(ax, sx[i]); ay = __fadd_rn(ay, sy[i]);
                axx = __fadd_rn(axx, __fmul_rn(sx[i], sx[i])); ayy = __fadd_rn(ayy, __fmul_rn(sy[i], sy[i]));
                ns++;
            }
        float fn = (float)ns;
        float mx = __fdiv_rn(ax, fn), my = __fdiv_rn(ay, fn);
        o[5 * (c - 1) + 0] = mx;
        o[5 * (c - 1) + 1] = my;
        o[5 * (c - 1) + 2] = __fsub_rn(__fdiv_rn(axx, fn), __fmul_rn(mx, mx));
        o[5 * (c - 1) + 3] = __fsub_rn(__fdiv_rn(ayy, fn), __fmul_rn(my, my));
        o[5 * (c - 1) + 4] = __fdiv_rn(fn, (float)K);
    }
    ncl[g] = max_id;
}

// get_dpf0 :1224-1263 + the field initialisation of get_dpf1 :1358-1376
__global__ void __launch_bounds__(kT) dpf0_kernel(const float *__restrict__ mvn, const int *__restrict__ ncl, int n, int K,
                                                  int *__restrict__ dpf0, int *__restrict__ id, float *__restrict__ dx,
                                                  float *__restrict__ dy, float *__restrict__ dxb, float *__restrict__ dyb,
                                                  float *__restrict__ noi, int *__restrict__ ctr) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    int sel = -1;
    for (int c = 0; c < ncl[g]; c++)
        if (mvn[((size_t)g * K + c) * 5 + 4] > 0.6f) { sel = c; break; }
    dpf0[g] = sel; id[g] = sel;
    if (sel >= 0) { dx[g] = mvn[((size_t)g * K + sel) * 5]; dy[g] = mvn[((size_t)g * K + sel) * 5 + 1]; }
    else { dx[g] = CUDART_NAN_F; dy[g] = CUDART_NAN_F; if (ncl[g] != 0) atomicAdd(&ctr[2], 1); }
    dxb[g] = CUDART_NAN_F; dyb[g] = CUDART_NAN_F; noi[g] = 1.0f;
}

// Fields of the halo rows before the first exchange: "nothing known yet".
__global__ void __launch_bounds__(kT) halo_init_kernel(int n, int *dpf0, int *id, float *dx, float *dy, float *dxb, float *dyb,
                                                       float *noi, int *ncl, int *bid, uint8_t *m0, uint8_t *m1) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    dpf0[g] = -1; id[g] = -1; ncl[g] = 0; bid[g] = -1;
    dx[g] = CUDART_NAN_F; dy[g] = CUDART_NAN_F; dxb[g] = CUDART_NAN_F; dyb[g] = CUDART_NAN_F; noi[g] = 1.0f;
    m0[g] = 0; m1[g] = 0;
}

// One Jacobi sweep of get_dpf1's interpolation, MIMC_module.c:1408-1566.
__global__ void __launch_bounds__(kT) dpf1_sweep_kernel(const float *__restrict__ dx, const float *__restrict__ dy,
                                                        float *__restrict__ dxb, float *__restrict__ dyb, float *noi,
                                                        const int *__restrict__ ncl, const double *__restrict__ apv,
                                                        const int *__restrict__ ruv, int nruv, const Band B,
                                                        float factor, int thres_n, float thres_weight, int *ctr,
                                                        const int *__restrict__ list, const int *__restrict__ list_len) {
    const int dimx = B.dimx, dimy = B.rows;
    // the holes still to fill are kept as a list (dense warps: a thread per grid node left one lane in thirty busy)
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i >= *list_len) return;
    const int g = list[i];
    if (!(isnan(__fadd_rn(dx[g], dy[g])) && ncl[g] != 0)) return;               // :1412
    const int cv = g / dimx, cu = g - cv * dimx;
    float dpe0 = __double2float_rn(__dmul_rn(apv[2 * (size_t)g], (double)factor));
    float dpe1 = __double2float_rn(__dmul_rn(-apv[2 * (size_t)g + 1], (double)factor));
    float mag_dpe = __fsqrt_rn(__fadd_rn(__fmul_rn(dpe0, dpe0), __fmul_rn(dpe1, dpe1)));
    // columns of vec_ruv_w_scale_mag: 0,1 offset; 2 weight; 3 scale; 4 |d|; 5 |apv|; 6 noi
    float r0[32], r1[32], r2[32], r3[32], r4[32], r5[32], r6[32];
    int nn = 0;
    for (int k = 0; k < nruv; k++) {
        int u = cu + ruv[2 * k], v = cv + ruv[2 * k + 1];
        if (u < 0 || u >= dimx || v < 0 || v >= dimy) continue;
        int h = v * dimx + u;
        float dn0 = dx[h], dn1 = dy[h];
        if (isnan(__fadd_rn(dn0, dn1))) continue;                                // :1430
        float apv0 = __fmul_rn((float)apv[2 * (size_t)h], factor);
        float apv1 = __fmul_rn(-(float)apv[2 * (size_t)h + 1], factor);
        float t = __fadd_rn(__fmul_rn(apv0, apv0), __fmul_rn(apv1, apv1));
        r0[nn] = (float)ruv[2 * k]; r1[nn] = (float)ruv[2 * k + 1];
        r4[nn] = __fsqrt_rn(__fadd_rn(__fmul_rn(dn0, dn0), __fmul_rn(dn1, dn1)));
        r5[nn] = __fsqrt_rn(t);
        r6[nn] = noi[h];
        r3[nn] = __double2float_rn(__ddiv_rn((double)r4[nn], __dsqrt_rn((double)t)));   // float / double :1444
        nn++;
    }
    if (nn < thres_n) return;                                                   // :1453
    float w_min = 1E+37f, w_max = -1E+37f;
    int id_w_max = 0, id_w_min = 0;
    for (int k = 0; k < nn; k++) {
        float mag_dxy = __fsqrt_rn(__fadd_rn(__fmul_rn(r0[k], r0[k]), __fmul_rn(r1[k], r1[k])));
        float w = __fdiv_rn(__fadd_rn(__fmul_rn(dpe0, r0[k]), __fmul_rn(dpe1, r1[k])), __fmul_rn(mag_dpe, mag_dxy));
        w = w > 0 ? w : -w;
        if (w >= thres_weight) {
            r2[k] = w;
            if (r3[k] > w_max) { w_max = r3[k]; id_w_max = k; }
            if (r3[k] < w_min) { w_min = r3[k]; id_w_min = k; }
        } else r2[k] = 0.0f;
    }
    r2[id_w_max] = 0.0f; r2[id_w_min] = 0.0f;                                   // :1496-1497
    float sum_w = 0.f, sum_w_dp = 0.f, sum_w_dpe = 0.f, sum_noi = 0.f;
    for (int k = 0; k < nn; k++) {
        // expf through the double-precision exp: correctly rounded to float except for one-in-millions double-rounding
        // cases, like glibc's expf -- CUDA's own expf (2 ulp) made the interpolated fields differ in the last bit
        const float ex = __double2float_rn(exp((double)__fadd_rn(-r5[k], 5.0f)));
        float w2 = __fdiv_rn(__fdiv_rn(1.0f, __fadd_rn(1.0f, ex)), 1.0f);   // :1514
        sum_w = __fadd_rn(sum_w, r2[k]);
        sum_w_dp = __fadd_rn(sum_w_dp, __fdiv_rn(__fmul_rn(__fmul_rn(r2[k], w2), r4[k]), r6[k]));
        sum_w_dpe = __fadd_rn(sum_w_dpe, __fdiv_rn(__fmul_rn(__fmul_rn(r2[k], w2), r5[k]), r6[k]));
        sum_noi = __fadd_rn(sum_noi, r6[k]);
    }
    if (sum_w >= 1.0f) {                                                         // :1547
        float factor_mag = __fdiv_rn(sum_w_dp, sum_w_dpe);
        dxb[g] = __fmul_rn(dpe0, factor_mag);
        dyb[g] = __fmul_rn(dpe1, factor_mag);
        // written in place like the reference; only nodes that are still NaN write, and
        // neighbours are only read where the committed field is non-NaN, so no sweep-internal
        // reader can observe this store (Jacobi-safe).
        noi[g] = __fadd_rn(__fdiv_rn(sum_noi, (float)nn), 1.0f);
        atomicAdd(&ctr[0], 1);                                                   // num_processed
    }
}

// Jacobi commit :1577-1589 + count of still-unprocessed nodes :1600-1611, over the hole list; the holes that remain
// form the next list (its order does not matter: the sweep is a Jacobi iteration)
__global__ void __launch_bounds__(kT) dpf1_commit_kernel(float *dx, float *dy, float *dxb, float *dyb, const int *ncl,
                                                         const int *__restrict__ list, const int *__restrict__ list_len,
                                                         int *__restrict__ next_list, int *next_len, int *ctr) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i >= *list_len) return;
    const int g = list[i];
    float a = dxb[g], b = dyb[g];
    if (!isnan(a) && !isnan(b)) { dx[g] = a; dy[g] = b; dxb[g] = CUDART_NAN_F; dyb[g] = CUDART_NAN_F; }
    if ((isnan(dx[g]) || isnan(dy[g])) && ncl[g] != 0) {
        atomicAdd(&ctr[1], 1);
        next_list[atomicAdd(next_len, 1)] = g;
    }
}

// Work list of a sweep: the (local) indices of the owned nodes with flag[g] != 0 / of the holes left by get_dpf0.
__global__ void __launch_bounds__(kT) list_flagged_kernel(const uint8_t *__restrict__ flag, int first, int n, int *__restrict__ list, int *len) {
    const int g = first + blockIdx.x * kT + threadIdx.x;
    if (g >= first + n) return;
    if (flag[g]) list[atomicAdd(len, 1)] = g;
}
__global__ void __launch_bounds__(kT) list_holes_kernel(const float *__restrict__ dx, const float *__restrict__ dy, const int *__restrict__ ncl,
                                                        int first, int n, int *__restrict__ list, int *len) {
    const int g = first + blockIdx.x * kT + threadIdx.x;
    if (g >= first + n) return;
    if ((isnan(dx[g]) || isnan(dy[g])) && ncl[g] != 0) list[atomicAdd(len, 1)] = g;
}

// 3x3 box smoothing of the filled nodes :1623-1666 (reads dx/dy, writes dxb/dyb for interior nodes)
__global__ void __launch_bounds__(kT) smooth_kernel(const float *dx, const float *dy, float *dxb, float *dyb,
                                                    const int *dpf0, const Band B) {
    const int dimx = B.dimx;
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= dimx * (B.own1 - B.own0)) return;
    g += B.own0 * dimx;
    int cv = g / dimx, cu = g - cv * dimx;
    const int cvg = cv + B.grow0;   // the interior test is about the GLOBAL grid
    if (cvg < 1 || cvg >= B.gdimy - 1 || cu < 1 || cu >= dimx - 1) return;
    if (dpf0[g] < 0 && !isnan(__fadd_rn(dx[g], dy[g]))) {
        float num = 0.f, sdx = 0.f, sdy = 0.f;
        for (int dv = -1; dv <= 1; dv++)
            for (int du = -1; du <= 1; du++) {
                int h = (cv + dv) * dimx + cu + du;
                if (!isnan(__fadd_rn(dx[h], dy[h]))) { sdx = __fadd_rn(sdx, dx[h]); sdy = __fadd_rn(sdy, dy[h]); num = __fadd_rn(num, 1.0f); }
            }
        dxb[g] = __fdiv_rn(sdx, num); dyb[g] = __fdiv_rn(sdy, num);
    } else { dxb[g] = dx[g]; dyb[g] = dy[g]; }
}

// copy-back of the smoothing (:1668-1675) fused with snap-to-nearest-cluster (:1680-1706)
__global__ void __launch_bounds__(kT) snap_kernel(float *dx, float *dy, const float *dxb, const float *dyb, int *id,
                                                  const float *mvn, const int *ncl, int K, const Band B) {
    const int dimx = B.dimx;
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= dimx * (B.own1 - B.own0)) return;
    g += B.own0 * dimx;
    int cv = g / dimx, cu = g - cv * dimx;
    const int cvg = cv + B.grow0;
    float x = dx[g], y = dy[g];
    if (cvg >= 1 && cvg < B.gdimy - 1 && cu >= 1 && cu < dimx - 1) { x = dxb[g]; y = dyb[g]; }
    if (id[g] < 0 && ncl[g] != 0) {
        float best = 1E+37f; int sel = 0;
        for (int c = 0; c < ncl[g]; c++) {
            float d0 = __fsub_rn(x, mvn[((size_t)g * K + c) * 5]), d1 = __fsub_rn(y, mvn[((size_t)g * K + c) * 5 + 1]);
            float sq = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
            if (sq < best) { best = sq; sel = c; }
        }
        id[g] = sel;
        x = mvn[((size_t)g * K + sel) * 5]; y = mvn[((size_t)g * K + sel) * 5 + 1];
    }
    dx[g] = x; dy[g] = y;
}

// ---------------------------------------------------------------------------------------
// pseudosmoothing :1986-2312, quadfit2 :2314-2409, GMA_double_inv :2430-2496
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) ps_init_kernel(const int *id, const float *mvn, int K, int n, uint8_t *mask0,
                                                     uint8_t *stack0, float *bx, float *by, int *bid) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    int c = id[g];
    uint8_t m = 0;
    if (c >= 0) m = ((double)mvn[((size_t)g * K + c) * 5 + 4] >= 0.6) ? 0 : 1;   // :2041
    mask0[g] = m; stack0[g] = m;
    bx[g] = CUDART_NAN_F; by[g] = CUDART_NAN_F; bid[g] = -1;
}

__device__ void inv6(double b[6][6], double I[6][6]) {
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) I[i][j] = (i == j) ? 1.0 : 0.0;
    for (int c1 = 0; c1 < 5; c1++) {
        double pivot = b[c1][c1];
        for (int c2 = c1 + 1; c2 < 6; c2++) {
            double coeff = __ddiv_rn(b[c2][c1], pivot);
            for (int c3 = 0; c3 < 6; c3++) {
                b[c2][c3] = __dsub_rn(b[c2][c3], __dmul_rn(b[c1][c3], coeff));
                I[c2][c3] = __dsub_rn(I[c2][c3], __dmul_rn(I[c1][c3], coeff));
            }
        }
    }
    for (int c1 = 5; c1 >= 0; c1--) {
        double pivot = b[c1][c1];
        for (int c2 = c1 - 1; c2 >= 0; c2--) {
            double coeff = __ddiv_rn(b[c2][c1], pivot);
            for (int c3 = 5; c3 >= 0; c3--) {
                b[c2][c3] = __dsub_rn(b[c2][c3], __dmul_rn(b[c1][c3], coeff));
                I[c2][c3] = __dsub_rn(I[c2][c3], __dmul_rn(I[c1][c3], coeff));
            }
        }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) I[i][j] = __ddiv_rn(I[i][j], b[i][i]);
}

__global__ void __launch_bounds__(kT) ps_sweep_kernel(const uint8_t *__restrict__ mask, uint8_t *next,
                                                      const uint8_t *__restrict__ stack0, const int *__restrict__ id,
                                                      const float *__restrict__ dx, const float *__restrict__ dy,
                                                      float *bx, float *by, int *bid, const float *__restrict__ mvn,
                                                      const int *__restrict__ ncl, int K, const double *__restrict__ apv,
                                                      const int *__restrict__ ruv, int nruv, const Band B, int *ctr,
                                                      const int *__restrict__ list, const int *__restrict__ list_len) {
    const int dimx = B.dimx, dimy = B.rows;
    const int i = blockIdx.x * kT + threadIdx.x;   // a thread per dirty node of the list (dense warps)
    if (i >= *list_len) return;
    const int g = list[i];
    if (!mask[g]) return;
    const int cv = g / dimx, cu = g - cv * dimx;
    signed char nu[MAXNB], nv[MAXNB];
    int nn = 0;
    for (int k = 0; k < nruv; k++) {
        int u = cu + ruv[2 * k], v = cv + ruv[2 * k + 1];
        if (u >= 0 && u < dimx && v >= 0 && v < dimy) {
            int h = v * dimx + u;
            if (!isnan(dx[h]) && !isnan(dy[h])) { nu[nn] = (signed char)ruv[2 * k]; nv[nn] = (signed char)ruv[2 * k + 1]; nn++; }
        }
    }
    if (nn < 10) return;                                                        // :2131
    const double e0 = 1500.0 / 300.0, e1 = e0 / 3.0;
    const double vx = apv[2 * (size_t)g], vy = apv[2 * (size_t)g + 1];
    const double den = __dmul_rn(__dmul_rn(e0, e1), __dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
    const double ITM0 = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(e1, vx), vx), __dmul_rn(__dmul_rn(e0, vy), vy)), den);
    const double ITM1 = __ddiv_rn(__dmul_rn(__dmul_rn(__dsub_rn(e0, e1), vx), vy), den);
    const double ITM3 = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(e1, vy), vy), __dmul_rn(__dmul_rn(e0, vx), vx)), den);
    // normal equations N = A^T W A and the two right-hand sides, observation order = ruv order
    double N[6][6], IN[6][6], rhs0[6], rhs1[6];
    for (int r = 0; r < 6; r++) { rhs0[r] = 0.0; rhs1[r] = 0.0; for (int c = 0; c < 6; c++) N[r][c] = 0.0; }
    for (int o = 0; o < nn; o++) {
        const double x = (double)nu[o], y = (double)nv[o];
        const double w = exp(-__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(ITM0, x), x), __dmul_rn(__dmul_rn(__dmul_rn(2.0, ITM1), x), y)),
                                        __dmul_rn(__dmul_rn(ITM3, y), y)));    // :2157
        const int h = (cv + nv[o]) * dimx + cu + nu[o];
        const double z0 = (double)dx[h], z1 = (double)dy[h];
        const double A[6] = {__dmul_rn(x, x), __dmul_rn(x, y), __dmul_rn(y, y), x, y, 1.0};
        for (int r = 0; r < 6; r++) {
            const double aw = __dmul_rn(A[r], w);
            for (int c = 0; c < 6; c++) N[r][c] = __dadd_rn(N[r][c], __dmul_rn(aw, A[c]));   // :2350
            rhs0[r] = __dadd_rn(rhs0[r], __dmul_rn(aw, z0));                                  // :2367
            rhs1[r] = __dadd_rn(rhs1[r], __dmul_rn(aw, z1));
        }
    }
    inv6(N, IN);
    double interp[2];
    for (int oc = 0; oc < 2; oc++) {
        const double *rhs = oc ? rhs1 : rhs0;
        double out = 0.0;
        for (int r = 0; r < 6; r++) {
            double coeff = 0.0;
            for (int c = 0; c < 6; c++) coeff = __dadd_rn(coeff, __dmul_rn(IN[r][c], rhs[c]));
            out = __dadd_rn(out, __dmul_rn(r == 5 ? 1.0 : 0.0, coeff));          // terms = (0,0,0,0,0,1) at (0,0)
        }
        interp[oc] = out;
    }
    const int cur = id[g], nc = ncl[g];
    double sq_min = 1E+37; int closest = -1;
    for (int c = 0; c < nc; c++) {
        double c0 = (double)mvn[((size_t)g * K + c) * 5], c1 = (double)mvn[((size_t)g * K + c) * 5 + 1];
        double a = __dsub_rn(interp[0], c0), b = __dsub_rn(interp[1], c1);
        double sq = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
        if (sq < sq_min) { sq_min = sq; closest = c; }
    }
    // closest < 0 (all distances NaN): the reference falls through with the PREVIOUS node's
    // values (:2182-2188, sequential state); defined here as "no update".
    if (closest < 0) return;
    const double g0 = (double)mvn[((size_t)g * K + cur) * 5], g1 = (double)mvn[((size_t)g * K + cur) * 5 + 1];
    const double k0 = (double)mvn[((size_t)g * K + closest) * 5], k1 = (double)mvn[((size_t)g * K + closest) * 5 + 1];
    const double d0 = __dsub_rn(g0, k0), d1 = __dsub_rn(g1, k1);
    if (__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)) < 0.0001) return;       // :2190
    bx[g] = (float)k0; by[g] = (float)k1; bid[g] = closest;
    ctr[0] = 1;                                                                  // flag_any_modification
    for (int o = 0; o < nn; o++) {
        int h = (cv + nv[o]) * dimx + cu + nu[o];
        if (stack0[h]) next[h] = 1;                                              // :2205
    }
}

__global__ void __launch_bounds__(kT) ps_commit_kernel(float *dx, float *dy, int *id, float *bx, float *by, int *bid, int n) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    if (bid[g] >= 0) {
        dx[g] = bx[g]; dy[g] = by[g]; id[g] = bid[g];
        bx[g] = CUDART_NAN_F; by[g] = CUDART_NAN_F; bid[g] = -1;
    }
}

// diff[k] = 1 iff stack[k] != next somewhere (fluctuation check :2237-2260); also counts next.
__global__ void __launch_bounds__(kT) ps_compare_kernel(const uint8_t *stack, size_t stride, const uint8_t *next, int nstack,
                                                        int n, int *diff, int *count) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    uint8_t v = next[g];
    if (v) atomicAdd(count, 1);
    for (int k = 0; k < nstack; k++)
        if (stack[(size_t)k * stride + g] != v) diff[k] = 1;
}

__global__ void __launch_bounds__(kT) pack_kernel(const int *id, const float *mvn, int K, int n, float *planes) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    int c = id[g];
    for (int k = 0; k < 5; k++) planes[(size_t)k * n + g] = c >= 0 ? mvn[((size_t)g * K + c) * 5 + k] : CUDART_NAN_F;   // :950-973
}

// main negates du, dv of the swapped passes on the host (MIMC_main.c:289-293); the drop-in redoes that on its device copy
__global__ void __launch_bounds__(kT) negate_uv_kernel(float *dp, int n) {
    int g = blockIdx.x * kT + threadIdx.x;
    if (g >= n) return;
    dp[3 * (size_t)g] = -dp[3 * (size_t)g];
    dp[3 * (size_t)g + 1] = -dp[3 * (size_t)g + 1];
}

// get_ruv_neighbor :1266-1327, restricted to the window that can satisfy the radius test
// (same float arithmetic, same row-major order).
int ruv_neighbor_host(const double *xyuvav, int dimx, int dimy, float radius, float mps, std::vector<int32_t> &out) {
    out.clear();
    int cu = dimx / 2, cv = dimy / 2;
    float cx = (float)xyuvav[6 * (size_t)cu + 0], cy = (float)xyuvav[6 * (size_t)cv * dimx + 1];
    float lim = (radius * mps) * (radius * mps);
    int reach = (int)radius + 3;
    int v0 = cv - reach < 0 ? 0 : cv - reach, v1 = cv + reach >= dimy ? dimy - 1 : cv + reach;
    int u0 = cu - reach < 0 ? 0 : cu - reach, u1 = cu + reach >= dimx ? dimx - 1 : cu + reach;
    for (int v = v0; v <= v1; v++)
        for (int u = u0; u <= u1; u++) {
            float fx = (float)xyuvav[6 * (size_t)u + 0], fy = (float)xyuvav[6 * (size_t)v * dimx + 1];
            float d0 = fx - cx, d1 = fy - cy;
            float sq = d0 * d0 + d1 * d1;
            if (sq <= lim) { out.push_back(u - cu); out.push_back(v - cv); }
        }
    return (int)out.size() / 2;
}

int post_alloc(mimc3cu_ctx *ctx, int32_t n, int32_t K) {
    if (!ctx->post) ctx->post = new Post();
    Post &P = *ctx->post;
    if ((size_t)n <= P.cap_n && K == P.K) return 0;
    post_free(ctx);
    ctx->post = new Post();
    Post &Q = *ctx->post;
    size_t N = (size_t)n;
    CU_CHECK(ctx, cudaMalloc(&Q.mvn, N * K * 5 * sizeof(float)));
    CU_CHECK(ctx, cudaMalloc(&Q.ncl, N * 4)); CU_CHECK(ctx, cudaMalloc(&Q.dpf0, N * 4)); CU_CHECK(ctx, cudaMalloc(&Q.id, N * 4));
    CU_CHECK(ctx, cudaMalloc(&Q.id1, N * 4)); CU_CHECK(ctx, cudaMalloc(&Q.bid, N * 4));
    for (float **p : {&Q.dx, &Q.dy, &Q.dxb, &Q.dyb, &Q.noi, &Q.dx1, &Q.dy1}) CU_CHECK(ctx, cudaMalloc(p, N * 4));
    CU_CHECK(ctx, cudaMalloc(&Q.mask[0], N)); CU_CHECK(ctx, cudaMalloc(&Q.mask[1], N));
    CU_CHECK(ctx, cudaMalloc(&Q.ruv, MAXNB * 2 * sizeof(int32_t)));
    CU_CHECK(ctx, cudaMalloc(&Q.ctr, 256 * sizeof(int32_t)));
    CU_CHECK(ctx, cudaMalloc(&Q.list[0], N * 4)); CU_CHECK(ctx, cudaMalloc(&Q.list[1], N * 4));
    CU_CHECK(ctx, cudaMalloc(&Q.apv, N * 2 * sizeof(double)));
    Q.cap_n = N; Q.n = n; Q.K = K;
    return 0;
}

int ensure_stack(mimc3cu_ctx *ctx, int32_t slots) {
    Post &P = *ctx->post;
    if (slots <= P.stack_cap) return 0;
    int32_t want = slots < 8 ? 8 : (slots * 2 > kMaxStack ? kMaxStack : slots * 2);
    uint8_t *ns = nullptr;
    CU_CHECK(ctx, cudaMalloc(&ns, (size_t)want * P.cap_n));
    if (P.stack) {
        CU_CHECK(ctx, cudaMemcpyAsync(ns, P.stack, (size_t)P.stack_cap * P.cap_n, cudaMemcpyDeviceToDevice, ctx->stream));
        CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        CU_CHECK(ctx, cudaFree(P.stack));
    }
    P.stack = ns; P.stack_cap = want;
    return 0;
}

}  // namespace

int post_negate_uv(mimc3cu_ctx *ctx, float *dp, int32_t n) {
    if (n <= 0) return 0;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    negate_uv_kernel<<<nblocks(n), kT, 0, ctx->stream>>>(dp, n);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

void post_free(mimc3cu_ctx *ctx) {
    if (!ctx->post) return;
    Post &P = *ctx->post;
    for (void *p : {(void *)P.mvn, (void *)P.ncl, (void *)P.dpf0, (void *)P.id, (void *)P.id1, (void *)P.bid, (void *)P.dx,
                    (void *)P.dy, (void *)P.dxb, (void *)P.dyb, (void *)P.noi, (void *)P.dx1, (void *)P.dy1, (void *)P.mask[0],
                    (void *)P.mask[1], (void *)P.stack, (void *)P.ruv, (void *)P.ctr, (void *)P.apv, (void *)P.list[0], (void *)P.list[1]})
        if (p) cudaFree(p);
    delete ctx->post;
    ctx->post = nullptr;
}

int post_cluster(mimc3cu_ctx *ctx, const float *dp, int32_t n, int32_t num_dp, float *mvn, int32_t *ncl) {
    if (num_dp < 1 || num_dp > MAXDP) return mimc3cu_fail(ctx, "cluster: num_dp must be in 1..%d", MAXDP);
    if (n <= 0) return 0;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    // rows of clusters that do not exist stay zero (the oracle's dense layout does the same)
    CU_CHECK(ctx, cudaMemsetAsync(mvn, 0, (size_t)n * num_dp * 5 * sizeof(float), ctx->stream));
    cluster_kernel<<<nblocks(n), kT, 0, ctx->stream>>>(dp, n, num_dp, mvn, ncl);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

// get_ruv_neighbor works on the GLOBAL grid geometry (its window is centred on the grid centre).
static int band_halo_rows(const mimc3cu_params *p) {
    float r = p->radius_neighbor_dpf1 > p->radius_neighbor_ps ? p->radius_neighbor_dpf1 : p->radius_neighbor_ps;
    int h = (int)r;
    if ((float)h < r) h++;
    return h < 1 ? 1 : h;
}
int post_band_halo(const mimc3cu_params *p) { return band_halo_rows(p); }

int post_run_band(mimc3cu_ctx *ctx, const float *dp, const double *xyuvav, const mimc3cu_params *p, int32_t own_row0,
                  int32_t own_rows, const mimc3cu_band_comm *comm, float *planes, int32_t *stats) {
    if (!p || !dp || !xyuvav || !planes) return mimc3cu_fail(ctx, "postprocess: null argument");
    const int32_t dimx = p->dimx, gdimy = p->dimy, K = p->num_dp;
    if (dimx < 2 || gdimy < 1) return mimc3cu_fail(ctx, "postprocess: bad grid %dx%d", gdimy, dimx);
    if (own_row0 < 0 || own_rows < 1 || own_row0 + own_rows > gdimy) return mimc3cu_fail(ctx, "postprocess: bad band [%d,+%d) of %d rows", own_row0, own_rows, gdimy);
    const int halo = band_halo_rows(p);
    // communication: the caller's callbacks (any transport), or the context's own NCCL communicator (comm.cu)
    const bool nccl = !comm && ctx->comm && ctx->comm->world > 1;
    const bool banded = comm || nccl;
    if (banded && own_rows < halo && own_rows != gdimy)
        return mimc3cu_fail(ctx, "postprocess: a band needs at least %d node rows (has %d)", halo, own_rows);
    Band B;
    B.dimx = dimx; B.gdimy = gdimy;
    const int ht = banded ? std::min(halo, own_row0) : 0, hb = banded ? std::min(halo, gdimy - own_row0 - own_rows) : 0;
    if (!banded && own_rows != gdimy) return mimc3cu_fail(ctx, "postprocess: a partial band needs a communicator");
    if (nccl && ((ht != 0 && ht != halo) || (hb != 0 && hb != halo)))
        return mimc3cu_fail(ctx, "postprocess: every band needs at least %d node rows", halo);
    B.grow0 = own_row0 - ht; B.rows = own_rows + ht + hb; B.own0 = ht; B.own1 = ht + own_rows;
    const int32_t nl = B.rows * dimx, n = own_rows * dimx, off = B.own0 * dimx;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    if (int rc = post_alloc(ctx, nl, K)) return rc;
    Post &P = *ctx->post;
    P.dimx = dimx; P.dimy = gdimy; P.n = n; P.own_off = off; P.band = B;
    cudaStream_t st = ctx->stream;
    const int nb = nblocks(n), nbl = nblocks(nl);
    int32_t h_ctr[4];

    auto exchange = [&](std::initializer_list<void *> arrays, std::initializer_list<int32_t> bytes) -> int {
        if (!banded || (ht == 0 && hb == 0 && own_rows == gdimy)) return 0;
        std::vector<void *> a(arrays); std::vector<int32_t> eb(bytes);
        if (nccl) return bandcomm_halo_exchange(ctx, a.data(), eb.data(), (int32_t)a.size(), dimx, B.rows, B.own0, B.own1, halo);   // enqueued, no host wait
        CU_CHECK(ctx, cudaStreamSynchronize(st));
        if (comm->halo_exchange(comm->user, a.data(), eb.data(), (int32_t)a.size())) return mimc3cu_fail(ctx, "postprocess: halo exchange failed");
        return 0;
    };
    // sweep counters, summed over all bands, on the host: the one synchronisation per sweep
    auto read_counters = [&](int32_t *host, int32_t count) -> int {
        if (nccl)
            if (int rc = bandcomm_allreduce_sum(ctx, P.ctr, count)) return rc;
        CU_CHECK(ctx, cudaMemcpyAsync(host, P.ctr, count * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CU_CHECK(ctx, cudaStreamSynchronize(st));
        if (comm && comm->allreduce_sum(comm->user, host, count)) return mimc3cu_fail(ctx, "postprocess: all-reduce failed");
        return 0;
    };

    // a-priori velocity columns (xyuvav[:,4:6]) of the local rows to the device
    {
        std::vector<double> apv((size_t)nl * 2);
        const double *src = xyuvav + 6 * (size_t)B.grow0 * dimx;
        for (int32_t g = 0; g < nl; g++) { apv[2 * (size_t)g] = src[6 * (size_t)g + 4]; apv[2 * (size_t)g + 1] = src[6 * (size_t)g + 5]; }
        CU_CHECK(ctx, cudaMemcpyAsync(P.apv, apv.data(), sizeof(double) * 2 * (size_t)nl, cudaMemcpyHostToDevice, st));
        CU_CHECK(ctx, cudaStreamSynchronize(st));
    }
    halo_init_kernel<<<nbl, kT, 0, st>>>(nl, P.dpf0, P.id, P.dx, P.dy, P.dxb, P.dyb, P.noi, P.ncl, P.bid, P.mask[0], P.mask[1]);
    if (int rc = post_cluster(ctx, dp, n, K, P.mvn + (size_t)off * K * 5, P.ncl + off)) return rc;
    CU_CHECK(ctx, cudaMemsetAsync(P.ctr, 0, 256 * sizeof(int32_t), st));
    dpf0_kernel<<<nb, kT, 0, st>>>(P.mvn + (size_t)off * K * 5, P.ncl + off, n, K, P.dpf0 + off, P.id + off, P.dx + off, P.dy + off,
                                   P.dxb + off, P.dyb + off, P.noi + off, P.ctr);
    ctx->launches += 2;
    if (int rc = read_counters(h_ctr, 4)) return rc;
    const int32_t holes0 = h_ctr[2];
    if (int rc = exchange({P.dx, P.dy, P.noi}, {4, 4, 4})) return rc;

    // ---- get_dpf1 :1330-1718 -------------------------------------------------------------
    std::vector<int32_t> ruv;
    int nruv = ruv_neighbor_host(xyuvav, dimx, gdimy, p->radius_neighbor_dpf1, p->meter_per_spacing, ruv);
    if (nruv > 32) return mimc3cu_fail(ctx, "postprocess: %d dpf1 neighbours exceed the kernel's limit of 32", nruv);
    for (int k = 0; k < nruv; k++) if (abs(ruv[2 * k + 1]) > halo) return mimc3cu_fail(ctx, "postprocess: neighbour offset beyond the halo");
    CU_CHECK(ctx, cudaMemcpyAsync(P.ruv, ruv.data(), sizeof(int32_t) * 2 * (size_t)nruv, cudaMemcpyHostToDevice, st));
    const float factor = (float)(1.0 / 365.0 * (double)p->dt / (double)p->mpp);   // :1393
    // the sweeps work on the list of holes: ctr[200 + k] is the length of list[k]
    int cur = 0;
    CU_CHECK(ctx, cudaMemsetAsync(P.ctr + 200, 0, 2 * sizeof(int32_t), st));
    list_holes_kernel<<<nb, kT, 0, st>>>(P.dx, P.dy, P.ncl, off, n, P.list[0], P.ctr + 200);
    ctx->launches++;
    // own holes of this band: sizes the sweep launches (the list only shrinks)
    int32_t own_holes = 0;
    CU_CHECK(ctx, cudaMemcpyAsync(&own_holes, P.ctr + 200, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU_CHECK(ctx, cudaStreamSynchronize(st));
    const int nbh = std::max(1, nblocks((size_t)own_holes));
    int32_t NOI = 0, unprocessed = 1;
    for (int thres_n = nruv - 1; thres_n >= 3; thres_n--) {
        float thres_weight = 0.5f;
        while (unprocessed != 0 && thres_weight >= 0.5) {
            thres_weight = (float)((double)thres_weight - 0.02);
            int32_t processed = 1;
            while (processed != 0) {
                NOI++;
                CU_CHECK(ctx, cudaMemsetAsync(P.ctr, 0, 2 * sizeof(int32_t), st));
                CU_CHECK(ctx, cudaMemsetAsync(P.ctr + 200 + (cur ^ 1), 0, sizeof(int32_t), st));
                dpf1_sweep_kernel<<<nbh, kT, 0, st>>>(P.dx, P.dy, P.dxb, P.dyb, P.noi, P.ncl, P.apv, P.ruv, nruv, B, factor,
                                                      thres_n, thres_weight, P.ctr, P.list[cur], P.ctr + 200 + cur);
                dpf1_commit_kernel<<<nbh, kT, 0, st>>>(P.dx, P.dy, P.dxb, P.dyb, P.ncl, P.list[cur], P.ctr + 200 + cur, P.list[cur ^ 1],
                                                       P.ctr + 200 + (cur ^ 1), P.ctr);
                cur ^= 1;
                ctx->launches += 2;
                if (int rc = read_counters(h_ctr, 2)) return rc;
                processed = h_ctr[0];
                unprocessed = h_ctr[1];
                if (processed != 0)
                    if (int rc = exchange({P.dx, P.dy, P.noi}, {4, 4, 4})) return rc;
            }
        }
    }
    smooth_kernel<<<nb, kT, 0, st>>>(P.dx, P.dy, P.dxb, P.dyb, P.dpf0, B);
    snap_kernel<<<nb, kT, 0, st>>>(P.dx, P.dy, P.dxb, P.dyb, P.id, P.mvn, P.ncl, K, B);
    ctx->launches += 2;
    CU_CHECK(ctx, cudaMemcpyAsync(P.id1, P.id, (size_t)nl * 4, cudaMemcpyDeviceToDevice, st));
    CU_CHECK(ctx, cudaMemcpyAsync(P.dx1, P.dx, (size_t)nl * 4, cudaMemcpyDeviceToDevice, st));
    CU_CHECK(ctx, cudaMemcpyAsync(P.dy1, P.dy, (size_t)nl * 4, cudaMemcpyDeviceToDevice, st));
    if (int rc = exchange({P.dx, P.dy}, {4, 4})) return rc;

    // ---- get_dpf_pseudosmoothing :1986-2312 ---------------------------------------------------
    nruv = ruv_neighbor_host(xyuvav, dimx, gdimy, p->radius_neighbor_ps, p->meter_per_spacing, ruv);
    if (nruv > MAXNB) return mimc3cu_fail(ctx, "postprocess: %d pseudosmoothing neighbours exceed the limit of %d", nruv, MAXNB);
    for (int k = 0; k < nruv; k++) if (abs(ruv[2 * k + 1]) > halo) return mimc3cu_fail(ctx, "postprocess: neighbour offset beyond the halo");
    CU_CHECK(ctx, cudaMemcpyAsync(P.ruv, ruv.data(), sizeof(int32_t) * 2 * (size_t)nruv, cudaMemcpyHostToDevice, st));
    if (int rc = ensure_stack(ctx, 8)) return rc;
    CU_CHECK(ctx, cudaMemsetAsync(P.stack, 0, (size_t)nl, st));
    // dxb/dyb double as dxy_ps_buffer
    ps_init_kernel<<<nb, kT, 0, st>>>(P.id + off, P.mvn + (size_t)off * K * 5, K, n, P.mask[0] + off, P.stack + off, P.dxb + off,
                                      P.dyb + off, P.bid + off);
    ctx->launches++;
    if (int rc = exchange({P.stack}, {1})) return rc;   // mask_dp_ps_initial of the halo rows (:2205)
    int32_t ps_noi = 0, nstack = 1;
    bool any = true;
    while (ps_noi <= 100 && any) {
        uint8_t *mask = P.mask[ps_noi % 2], *next = P.mask[(ps_noi + 1) % 2];
        ps_noi++;
        if (int rc = ensure_stack(ctx, nstack + 1)) return rc;
        CU_CHECK(ctx, cudaMemsetAsync(next, 0, (size_t)nl, st));
        CU_CHECK(ctx, cudaMemsetAsync(P.ctr, 0, 128 * sizeof(int32_t), st));
        // the dirty nodes of this sweep as a list (ctr[200] = its length); a dirty node does a 6x6 weighted least-squares fit
        // over up to 81 neighbours, so dense warps matter
        CU_CHECK(ctx, cudaMemsetAsync(P.ctr + 200, 0, sizeof(int32_t), st));
        list_flagged_kernel<<<nb, kT, 0, st>>>(mask, off, n, P.list[0], P.ctr + 200);
        ps_sweep_kernel<<<nb, kT, 0, st>>>(mask, next, P.stack, P.id, P.dx, P.dy, P.dxb, P.dyb, P.bid, P.mvn, P.ncl, K, P.apv, P.ruv,
                                           nruv, B, P.ctr, P.list[0], P.ctr + 200);
        ctx->launches++;
        ps_commit_kernel<<<nb, kT, 0, st>>>(P.dx + off, P.dy + off, P.id + off, P.dxb + off, P.dyb + off, P.bid + off, n);
        ctx->launches += 2;
        if (nccl) {                   // dirty flags scattered into the neighbours' rows: OR them into their owners
            if (int rc = bandcomm_halo_or_reduce(ctx, next, dimx, B.rows, B.own0, B.own1, halo)) return rc;
        } else if (comm && (ht || hb)) {
            CU_CHECK(ctx, cudaStreamSynchronize(st));
            if (comm->halo_or_reduce(comm->user, next)) return mimc3cu_fail(ctx, "postprocess: halo OR-reduce failed");
        }
        ps_compare_kernel<<<nb, kT, 0, st>>>(P.stack + off, P.cap_n, next + off, nstack, n, P.ctr + 8, P.ctr + 1);
        ctx->launches++;
        int32_t h[128];
        if (int rc = read_counters(h, 128)) return rc;
        any = h[0] != 0;
        bool fluct = false;
        for (int k = nstack - 1; k >= 0; k--) if (h[8 + k] == 0) { fluct = true; break; }
        if (fluct) { ps_noi--; break; }
        CU_CHECK(ctx, cudaMemcpyAsync(P.stack + (size_t)nstack * P.cap_n, next, (size_t)nl, cudaMemcpyDeviceToDevice, st));
        nstack++;
        if (any)
            if (int rc = exchange({P.dx, P.dy}, {4, 4})) return rc;
    }
    pack_kernel<<<nb, kT, 0, st>>>(P.id + off, P.mvn + (size_t)off * K * 5, K, n, planes);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    CU_CHECK(ctx, cudaStreamSynchronize(st));
    if (stats) { stats[0] = NOI; stats[1] = ps_noi; stats[2] = holes0; stats[3] = 0; }
    return 0;
}

int post_run(mimc3cu_ctx *ctx, const float *dp, const double *xyuvav, const mimc3cu_params *p, float *planes, int32_t *stats) {
    if (!p) return mimc3cu_fail(ctx, "postprocess: null argument");
    return post_run_band(ctx, dp, xyuvav, p, 0, p->dimy, nullptr, planes, stats);
}

int post_stage(mimc3cu_ctx *ctx, int32_t which, void *host) {
    if (!ctx->post || ctx->post->n == 0) return mimc3cu_fail(ctx, "postprocess_stage: no postprocess run yet");
    Post &P = *ctx->post;
    const void *src = nullptr;
    const size_t off = (size_t)P.own_off;
    switch (which) {
        case 0: src = P.dpf0 + off; break;
        case 1: src = P.id1 + off; break;
        case 2: src = P.dx1 + off; break;
        case 3: src = P.dy1 + off; break;
        case 4: src = P.id + off; break;
        case 5: src = P.dx + off; break;
        case 6: src = P.dy + off; break;
        case 7: src = P.ncl + off; break;
        default: return mimc3cu_fail(ctx, "postprocess_stage: unknown field %d", which);
    }
    CU_CHECK(ctx, cudaMemcpyAsync(host, src, (size_t)P.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// main()'s tail, MIMC_main.c:356-402.  The mean is a sequential float sum over the grid in
// row-major order (host code in the reference's main, kept on the host for bit equality);
// the element-wise conversion is done on the host copy as well and written back.
int post_finalize(mimc3cu_ctx *ctx, float *planes, const mimc3cu_params *p, float *du_cp_out, float *dv_cp_out) {
    const size_t n = (size_t)p->dimx * p->dimy;
    std::vector<float> h(4 * n);
    CU_CHECK(ctx, cudaMemcpyAsync(h.data(), planes, 4 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    float sdu = 0.0f, sdv = 0.0f;
    int32_t num_cp = 0;
    for (size_t g = 0; g < n; g++) {
        float a = h[g], b = h[n + g];
        if (!isnan(a) && !isnan(b)) { sdu += a; sdv += b; num_cp++; }
    }
    const float du_cp = sdu / (float)num_cp, dv_cp = sdv / (float)num_cp;
    const float f = p->mpp / p->dt * 365;
    for (size_t g = 0; g < n; g++) {
        float a = h[g] - du_cp, b = h[n + g] - dv_cp;
        h[g] = a * f;
        h[n + g] = -b * f;
        h[2 * n + g] = (float)(sqrt((double)h[2 * n + g]) * (double)f);
        h[3 * n + g] = (float)(sqrt((double)h[3 * n + g]) * (double)f);
    }
    CU_CHECK(ctx, cudaMemcpyAsync(planes, h.data(), 4 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (du_cp_out) *du_cp_out = du_cp;
    if (dv_cp_out) *dv_cp_out = dv_cp;
    return 0;
}
