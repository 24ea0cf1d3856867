/* mimc3_dropin -- the reference's own operator interface, re-declared as an ABI contract.
 *
 * libmimc3cu_dropin.a implements the five entry points the reference driver (MIMC_main.c)
 * calls in MIMC_module.c, with the signatures of MIMC_module.h:34-67, on top of the C ABI of
 * include/mimc3cu.h.  Linking it (plus libmimc3cu.so) INSTEAD of MIMC_module.c turns the
 * unchanged four-argument CLI into a GPU program.
 *
 * The types below restate layouts, not code: they must stay bit-compatible with
 *   GMA.h:43-91          {int32 ncols; int32 nrows; T **val; T *data;}  (ncols FIRST)
 *   MIMC_module.h:10-25  struct param
 * A translation unit that already includes the reference's GMA.h / MIMC_module.h must define
 * MIMC3_DROPIN_USE_REFERENCE_TYPES before including this header.
 */
#ifndef MIMC3_DROPIN_H
#define MIMC3_DROPIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MIMC3_DROPIN_USE_REFERENCE_TYPES
typedef struct { int32_t ncols; int32_t nrows; uint8_t **val; uint8_t *data; } GMA_uint8;   /* GMA.h:43-49; val rows are separate mallocs, data unset (GMA.c:10-20) */
typedef struct { int32_t ncols; int32_t nrows; int32_t **val; int32_t *data; } GMA_int32;   /* GMA.h:67-73 */
typedef struct { int32_t ncols; int32_t nrows; float **val; float *data; } GMA_float;       /* GMA.h:75-82 */
typedef struct { int32_t ncols; int32_t nrows; double **val; double *data; } GMA_double;    /* GMA.h:84-90 */

typedef struct param {                                                                       /* MIMC_module.h:10-25 */
    int32_t vec_ocw[4];
    float AW_CRE;
    float AW_SF;
    float spacing_grid;
    float radius_neighbor;
    float radius_neighbor_dpf1;
    float radius_neighbor_ps;
    float meter_per_spacing;
    float mpp;
    int32_t num_cp_max;
    int32_t num_cp_min;
    float ratio_cp;
    float thres_spd_cp;
} param;
#endif

/* Globals DEFINED by the driver (MIMC_main.c:38-42) and read by the module (MIMC_module.h:28-32). */
extern float dt;
extern int32_t num_dp;
extern int32_t num_grid, dimx_vmap, dimy_vmap;
extern param param_mimc2;
extern GMA_float **kernel;

/* Allocators of the driver's GMA.c (GMA.c:54-84): every object returned to the driver is built
 * with them, because the driver releases it with GMA_*_destroy / free(). */
GMA_int32 *GMA_int32_create(int32_t size_row, int32_t size_col);
GMA_float *GMA_float_create(int32_t size_row, int32_t size_col);

/* ---- the five entry points (MIMC_module.h) ------------------------------------------------ */
/* :34  returns 1 (ok) / -1 (not enough control points); offset[2], flag_cp (n x 1) out */
int get_offset_image(GMA_float *i0, GMA_float *i1, GMA_float **kernel, GMA_double *xyuvav, int32_t *offset, GMA_uint8 *flag_cp);
/* :39  n ragged (P x 2) pivot lists; `param` by value */
GMA_int32 **get_uv_pivot(GMA_double *xyuvav, float dt, param param_mimc2, int32_t ocw, GMA_float *i1);
/* :44  new (n x 3) [du, dv, ncc] */
GMA_float *matching_ncc_dlc_2(GMA_float *i0, GMA_float *i1, GMA_double *xyuvav, int32_t *offset, GMA_int32 **uv_pivot,
                              int32_t ocw, float AW_CRE, float AW_SF);
/* :48  five (dimy x dimx) planes [du, dv, var u, var v, support] */
GMA_float **mimc2_postprocess(GMA_float **dp, GMA_double *xyuvav, float dt);
/* :67  in place on `out` */
void GMA_float_conv2(GMA_float *in, GMA_float *kernel, GMA_float *out);

/* Optional: release device memory before exit (also registered with atexit). */
void mimc3_dropin_shutdown(void);
/* Device copies of images, nodes and results are keyed by the HOST payload pointer (img->data, xyuvav->data).  A
 * driver that rewrites such a buffer in place, or frees it and gets the same address back for other content, must
 * call this before the next module call (get_offset_image does it for everything: it opens a new image pair). */
void mimc3_dropin_invalidate(const void *host_payload);

#ifdef __cplusplus
}
#endif
#endif /* MIMC3_DROPIN_H */
