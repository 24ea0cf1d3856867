/* Flat-array harness around the UNMODIFIED reference translation units.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs may load the library this file is linked into (oracle/_ref/libmimc3ref.so).
 * Nothing in the shipped CUDA path depends on it.
 *
 * The reference sources are compiled where they lie (/root/reference/MIMC_module.c,
 * GMA.c) by oracle/Makefile with the README flags (-fopenmp -O3, README.md:21) plus
 * `-include shim/zalloc.h`.  This file is new code: it defines the
 * globals the module imports from MIMC_main.c:38-42, builds the GMA_* structs the
 * module's entry points expect (GMA.h:43-91) out of flat row-major arrays, calls
 * the reference functions, and copies results back into flat arrays so that
 * Python/ctypes can drive the reference itself as the parity oracle.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <omp.h>
#include "GMA.h"
#include "MIMC_module.h"

/* globals normally defined by MIMC_main.c:38-42 */
float dt;
int32_t num_dp;
int32_t num_grid, dimx_vmap, dimy_vmap;
param param_mimc2;
GMA_float **kernel;

/* srand(time(NULL)) in GMA_double_randperm_row (MIMC_module.c:516) makes the CP
 * stage non-deterministic.  The library is linked -Bsymbolic, so the
 * reference's call binds to this definition and tests can pin the seed. */
static time_t g_fake_time = 0;
static int g_use_fake_time = 0;
time_t time(time_t *t) {
    time_t v;
    if (g_use_fake_time) v = g_fake_time;
    else { struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts); v = ts.tv_sec; }
    if (t) *t = v;
    return v;
}
void ref_set_fake_time(int64_t t, int enable) { g_fake_time = (time_t)t; g_use_fake_time = enable; }

static void make_kernels(void) {
    /* MIMC_main.c:175-196 */
    if (kernel) return;
    kernel = (GMA_float **)malloc(sizeof(GMA_float *) * 3);
    kernel[0] = GMA_float_create(1, 3);
    kernel[1] = GMA_float_create(3, 1);
    kernel[2] = GMA_float_create(3, 3);
    kernel[0]->val[0][0] = -1; kernel[0]->val[0][1] = 0; kernel[0]->val[0][2] = 1;
    kernel[1]->val[0][0] = -1; kernel[1]->val[1][0] = 0; kernel[1]->val[2][0] = 1;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) kernel[2]->val[i][j] = -1.0 / 8;
    kernel[2]->val[1][1] = 1.0;
}

/* Defaults of MIMC_main.c:134-168, then the three values main derives from xyuvav
 * rows 0-1 (:221-223) and the grid dimensions (:211-219). */
void ref_set_globals(float dt_, int32_t num_dp_, int32_t num_grid_, int32_t dimx, int32_t dimy,
                     float mpp, float spacing_grid, float meter_per_spacing) {
    dt = dt_; num_dp = num_dp_; num_grid = num_grid_; dimx_vmap = dimx; dimy_vmap = dimy;
    param_mimc2.vec_ocw[0] = 7; param_mimc2.vec_ocw[1] = 15;
    param_mimc2.vec_ocw[2] = 30; param_mimc2.vec_ocw[3] = 40;
    param_mimc2.AW_CRE = 10.0; param_mimc2.AW_SF = 1.8;
    param_mimc2.radius_neighbor = 5.0;
    param_mimc2.radius_neighbor_dpf1 = 1000 / 300;
    param_mimc2.radius_neighbor_ps = 5.0;
    param_mimc2.num_cp_max = 500; param_mimc2.num_cp_min = 50;
    param_mimc2.ratio_cp = 0.03; param_mimc2.thres_spd_cp = 10;
    param_mimc2.mpp = mpp; param_mimc2.spacing_grid = spacing_grid;
    param_mimc2.meter_per_spacing = meter_per_spacing;
    make_kernels();
}

/* Borrowing views: struct + row-pointer table over caller memory (never destroyed
 * with GMA_*_destroy). */
static GMA_float *view_float(float *data, int32_t nrows, int32_t ncols) {
    GMA_float *g = (GMA_float *)malloc(sizeof(GMA_float));
    g->nrows = nrows; g->ncols = ncols; g->data = data;
    g->val = (float **)malloc(sizeof(float *) * (size_t)nrows);
    for (int32_t r = 0; r < nrows; r++) g->val[r] = data + (size_t)r * ncols;
    return g;
}
static GMA_double *view_double(double *data, int32_t nrows, int32_t ncols) {
    GMA_double *g = (GMA_double *)malloc(sizeof(GMA_double));
    g->nrows = nrows; g->ncols = ncols; g->data = data;
    g->val = (double **)malloc(sizeof(double *) * (size_t)nrows);
    for (int32_t r = 0; r < nrows; r++) g->val[r] = data + (size_t)r * ncols;
    return g;
}
static void drop_view(void *g_, void *val) { free(val); free(g_); }

/* get_uv_pivot (MIMC_module.c:543-602) -> CSR.  Returns total pivots, or -(needed)
 * if `cap` is too small. csr_off has n+1 entries. */
int64_t ref_get_uv_pivot(double *xyuvav, int32_t n, float dt_, int32_t ocw, int32_t H, int32_t W,
                         int32_t *csr_off, int32_t *piv, int64_t cap) {
    GMA_double *x = view_double(xyuvav, n, 6);
    GMA_float img; img.nrows = H; img.ncols = W; img.val = NULL; img.data = NULL;
    GMA_int32 **p = get_uv_pivot(x, dt_, param_mimc2, ocw, &img);
    int64_t tot = 0;
    for (int32_t i = 0; i < n; i++) tot += p[i]->nrows;
    int ok = tot <= cap;
    int64_t k = 0;
    for (int32_t i = 0; i < n; i++) {
        csr_off[i] = (int32_t)k;
        if (ok) for (int32_t j = 0; j < p[i]->nrows; j++) { piv[2 * (k + j)] = p[i]->val[j][0]; piv[2 * (k + j) + 1] = p[i]->val[j][1]; }
        k += p[i]->nrows;
        GMA_int32_destroy(p[i]);
    }
    csr_off[n] = (int32_t)k;
    free(p);
    drop_view(x, x->val);
    return ok ? tot : -tot;
}

/* matching_ncc_dlc_2 (MIMC_module.c:805-842).  `sign` = -1 reproduces main's in-place
 * negation of the pivots before the swapped pass (MIMC_main.c:272-279). Returns seconds. */
double ref_match(float *i0, float *i1, int32_t H, int32_t W, double *xyuvav, int32_t n, int32_t *offset,
                 const int32_t *csr_off, const int32_t *piv, int32_t sign, int32_t ocw, float *out) {
    GMA_float *g0 = view_float(i0, H, W), *g1 = view_float(i1, H, W);
    GMA_double *x = view_double(xyuvav, n, 6);
    GMA_int32 **p = (GMA_int32 **)malloc(sizeof(GMA_int32 *) * (size_t)n);
    for (int32_t i = 0; i < n; i++) {
        int32_t np = csr_off[i + 1] - csr_off[i];
        p[i] = GMA_int32_create(np, 2);
        for (int32_t j = 0; j < np; j++) {
            p[i]->val[j][0] = sign * piv[2 * ((int64_t)csr_off[i] + j)];
            p[i]->val[j][1] = sign * piv[2 * ((int64_t)csr_off[i] + j) + 1];
        }
    }
    double t0 = omp_get_wtime();
    GMA_float *dp = matching_ncc_dlc_2(g0, g1, x, offset, p, ocw, param_mimc2.AW_CRE, param_mimc2.AW_SF);
    double t1 = omp_get_wtime();
    memcpy(out, dp->data, sizeof(float) * 3 * (size_t)n);
    GMA_float_destroy(dp);
    for (int32_t i = 0; i < n; i++) GMA_int32_destroy(p[i]);
    free(p);
    drop_view(x, x->val); drop_view(g0, g0->val); drop_view(g1, g1->val);
    return t1 - t0;
}

/* GMA_float_conv2 (MIMC_module.c:2517-2585); `out` is read-modify-written in place
 * exactly like main's reused i0c/i1c buffers (MIMC_main.c:304-310). */
void ref_conv2(float *in, int32_t H, int32_t W, int32_t kernel_id, float *out) {
    make_kernels();
    GMA_float *gi = view_float(in, H, W), *go = view_float(out, H, W);
    GMA_float_conv2(gi, kernel[kernel_id], go);
    drop_view(gi, gi->val); drop_view(go, go->val);
}

/* calc_mean_var_num_dp_cluster (MIMC_module.c:994-1130) -> dense (n, num_dpoi, 5) +
 * per-node cluster counts. dp is num_dpoi arrays of n*3 floats, concatenated. */
void ref_cluster(float *dp, int32_t n, int32_t num_dpoi, float *mvn, int32_t *ncl) {
    GMA_float **d = (GMA_float **)malloc(sizeof(GMA_float *) * (size_t)num_dpoi);
    for (int32_t a = 0; a < num_dpoi; a++) d[a] = view_float(dp + (size_t)a * n * 3, n, 3);
    GMA_float **m = calc_mean_var_num_dp_cluster(d, num_dpoi);
    for (int32_t g = 0; g < n; g++) {
        ncl[g] = m[g]->nrows;
        for (int32_t c = 0; c < m[g]->nrows; c++)
            for (int k = 0; k < 5; k++) mvn[((size_t)g * num_dpoi + c) * 5 + k] = m[g]->val[c][k];
        GMA_float_destroy(m[g]);
    }
    free(m);
    for (int32_t a = 0; a < num_dpoi; a++) drop_view(d[a], d[a]->val);
    free(d);
}

/* mimc2_postprocess (MIMC_module.c:893-991): 32 x (n,3) -> 5 planes (dimy, dimx). */
void ref_postprocess(float *dp, double *xyuvav, int32_t n, float *planes) {
    GMA_float **d = (GMA_float **)malloc(sizeof(GMA_float *) * (size_t)num_dp);
    for (int32_t a = 0; a < num_dp; a++) d[a] = view_float(dp + (size_t)a * n * 3, n, 3);
    GMA_double *x = view_double(xyuvav, n, 6);
    GMA_float **v = mimc2_postprocess(d, x, dt);
    size_t np = (size_t)dimx_vmap * dimy_vmap;
    for (int k = 0; k < 5; k++) { memcpy(planes + k * np, v[k]->data, sizeof(float) * np); GMA_float_destroy(v[k]); }
    free(v);
    drop_view(x, x->val);
    for (int32_t a = 0; a < num_dp; a++) drop_view(d[a], d[a]->val);
    free(d);
}

/* Stage-level access for differential tests of the postprocess chain
 * (MIMC_module.c:1224-1718, 1986-2312): runs cluster -> dpf0 -> dpf1 -> pseudosmoothing
 * and returns the intermediate fields. */
void ref_postprocess_stages(float *dp, double *xyuvav, int32_t n,
                            int32_t *dpf0_out, int32_t *dpf1_id, float *dpf1_dx, float *dpf1_dy,
                            int32_t *ps_id, float *ps_dx, float *ps_dy) {
    GMA_float **d = (GMA_float **)malloc(sizeof(GMA_float *) * (size_t)num_dp);
    for (int32_t a = 0; a < num_dp; a++) d[a] = view_float(dp + (size_t)a * n * 3, n, 3);
    GMA_double *x = view_double(xyuvav, n, 6);
    size_t np = (size_t)dimx_vmap * dimy_vmap;
    GMA_float **mvn = calc_mean_var_num_dp_cluster(d, num_dp);
    GMA_int32 *dpf0 = get_dpf0(mvn, 0.6);
    memcpy(dpf0_out, dpf0->data, sizeof(int32_t) * np);
    GMA_int32 *ruv = get_ruv_neighbor(x, param_mimc2.radius_neighbor_dpf1);
    GMA_float *dx = GMA_float_create(dimy_vmap, dimx_vmap), *dy = GMA_float_create(dimy_vmap, dimx_vmap);
    get_dpf1(dpf0, dx, dy, ruv, mvn, x);
    memcpy(dpf1_id, dpf0->data, sizeof(int32_t) * np);
    memcpy(dpf1_dx, dx->data, sizeof(float) * np);
    memcpy(dpf1_dy, dy->data, sizeof(float) * np);
    GMA_int32_destroy(ruv);
    ruv = get_ruv_neighbor(x, param_mimc2.radius_neighbor_ps);
    get_dpf_pseudosmoothing(dpf0, dx, dy, ruv, mvn, x);
    memcpy(ps_id, dpf0->data, sizeof(int32_t) * np);
    memcpy(ps_dx, dx->data, sizeof(float) * np);
    memcpy(ps_dy, dy->data, sizeof(float) * np);
    GMA_int32_destroy(ruv); GMA_int32_destroy(dpf0); GMA_float_destroy(dx); GMA_float_destroy(dy);
    for (int32_t g = 0; g < n; g++) GMA_float_destroy(mvn[g]);
    free(mvn);
    drop_view(x, x->val);
    for (int32_t a = 0; a < num_dp; a++) drop_view(d[a], d[a]->val);
    free(d);
}

/* get_offset_image (MIMC_module.c:33-492). Returns the reference's return code. */
int ref_get_offset_image(float *i0, float *i1, int32_t H, int32_t W, double *xyuvav, int32_t n,
                         int32_t *offset, uint8_t *flag_cp) {
    make_kernels();
    GMA_float *g0 = view_float(i0, H, W), *g1 = view_float(i1, H, W);
    GMA_double *x = view_double(xyuvav, n, 6);
    GMA_uint8 *f = GMA_uint8_create(n, 1);
    for (int32_t i = 0; i < n; i++) f->val[i][0] = 0;
    offset[0] = 0; offset[1] = 0;
    int rc = get_offset_image(g0, g1, kernel, x, offset, f);
    for (int32_t i = 0; i < n; i++) flag_cp[i] = f->val[i][0];
    GMA_uint8_destroy(f);
    drop_view(x, x->val); drop_view(g0, g0->val); drop_view(g1, g1->val);
    return rc;
}

int ref_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to its workers: bench.py's reference arm sets the team size explicitly */
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* The reference prints progress and debug matrices to stdout (e.g. the node (15,30)
 * dump inside get_dpf1, MIMC_module.c:1501-1544).  Tests silence it. */
#include <fcntl.h>
#include <unistd.h>
static int g_saved_stdout = -1;
void ref_quiet(int on) {
    fflush(stdout);
    if (on && g_saved_stdout < 0) {
        g_saved_stdout = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1); close(nul);
    } else if (!on && g_saved_stdout >= 0) {
        dup2(g_saved_stdout, 1); close(g_saved_stdout); g_saved_stdout = -1;
    }
}
