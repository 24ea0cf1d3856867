/* mimc3cu -- C ABI of the B200-native MIMC3 matching library (libmimc3cu.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch or C++ types.  It
 * replaces the reference translation unit MIMC_module.c.  Two layers:
 *
 *   1. mimc3cu_*  (this header): explicit, handle-based API used by tests, bench.py and
 *      by layer 2.  Each entry point cites the reference function it replaces.
 *   2. The five MIMC_module.h entry points themselves (get_offset_image, get_uv_pivot,
 *      matching_ncc_dlc_2, GMA_float_conv2, mimc2_postprocess), implemented over layer 1
 *      in mimc3_b200/csrc/dropin.c so that the reference's unchanged MIMC_main.c links
 *      against this library instead of MIMC_module.c.
 *
 * Conventions
 *   - Images are row-major float32 (H rows, W columns), exactly the payload of the
 *     reference's GMA_float (GMA.h:73-80).
 *   - xyuvav is row-major float64 (n, 6): x, y, u, v, vx, vy (MIMC_main.c:203, README.md:36).
 *   - Pivots are CSR: off[n+1] (int32) + piv[total][2] (int32 u, v), replacing the ragged
 *     GMA_int32** of MIMC_module.h:39.
 *   - dp ("displacement") arrays are (n, 3) float32 [du, dv, ncc] as returned by
 *     matching_ncc_dlc_2 (MIMC_module.c:805-842).
 *   - All functions return 0 on success, non-zero on error; mimc3cu_last_error() gives
 *     the message.  There is NO CPU fallback: without a CUDA device every compute entry
 *     point fails.
 *   - Calls are synchronous with respect to the host unless the name ends in _async;
 *     device work is issued on the context's stream (mimc3cu_stream).
 */
#ifndef MIMC3CU_H
#define MIMC3CU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMC3CU_VERSION 100

typedef struct mimc3cu_ctx mimc3cu_ctx;

/* Parameters the reference hard-codes in MIMC_main.c:134-168 / derives at :221-223. */
typedef struct mimc3cu_params {
    int32_t vec_ocw[4];          /* chip half-widths {7,15,30,40}          MIMC_main.c:134-137 */
    float AW_CRE;                /* 10                                      :153 */
    float AW_SF;                 /* 1.8                                     :154 */
    float mpp;                   /* metres per pixel                        :221 */
    float meter_per_spacing;     /*                                         :223 */
    float radius_neighbor_dpf1;  /* 1000/300 = 3 (integer division)         :161 */
    float radius_neighbor_ps;    /* 5                                       :162 */
    float dt;                    /* days between the images                 :114 */
    int32_t dimx, dimy;          /* node grid                               :211-219 */
    int32_t num_dp;              /* attempts per node (32)                  :74 */
    /* control-point stage, MIMC_main.c:165-168 */
    int32_t num_cp_max;          /* 500 */
    int32_t num_cp_min;          /* 50 */
    float ratio_cp;              /* 0.03 */
    float thres_spd_cp;          /* 10 m/yr */
} mimc3cu_params;

void mimc3cu_default_params(mimc3cu_params *p);

/* ---- context ------------------------------------------------------------------ */
int mimc3cu_version(void);
/* Number of CUDA devices visible (0 => every compute call will fail). */
int mimc3cu_device_count(void);
int mimc3cu_create(int device, mimc3cu_ctx **ctx);
void mimc3cu_destroy(mimc3cu_ctx *ctx);
const char *mimc3cu_last_error(const mimc3cu_ctx *ctx); /* ctx may be NULL: global error */
/* cudaStream_t of the context as an opaque pointer (for event timing by the caller). */
void *mimc3cu_stream(mimc3cu_ctx *ctx);
int mimc3cu_sync(mimc3cu_ctx *ctx);
/* Kernel launches issued by this context since creation (for the bench's gpu_launches). */
int64_t mimc3cu_launch_count(const mimc3cu_ctx *ctx);

/* ---- images (replace GMA_float payloads living in host RAM) --------------------- */
/* Allocate an (H, W) float32 image in HBM, zero-initialised (the
 * reference's conv2 output buffers behave as zero-initialised memory). */
int mimc3cu_image_create(mimc3cu_ctx *ctx, int32_t H, int32_t W, int32_t *handle);
int mimc3cu_image_destroy(mimc3cu_ctx *ctx, int32_t handle);
/* Host -> HBM (pageable or pinned host memory). */
int mimc3cu_image_upload(mimc3cu_ctx *ctx, int32_t handle, const float *host);
/* u8 / u16 -> f32 cast on the device: GMA_float_load_tiff's per-pixel loop, GMA.c:288-310. */
int mimc3cu_image_upload_u8(mimc3cu_ctx *ctx, int32_t handle, const uint8_t *host);
int mimc3cu_image_upload_u16(mimc3cu_ctx *ctx, int32_t handle, const uint16_t *host);
/* Device -> device copy from a caller-owned device buffer (e.g. a torch tensor). */
int mimc3cu_image_copy_from_device(mimc3cu_ctx *ctx, int32_t handle, const float *dev);
int mimc3cu_image_download(mimc3cu_ctx *ctx, int32_t handle, float *host);
/* Zero the payload (asynchronous on the context stream). */
int mimc3cu_image_fill_zero(mimc3cu_ctx *ctx, int32_t handle);
/* Raw device pointer of the image payload, read-only: the matcher caches statistics and a summed-area table per
 * image.  Whoever writes the payload behind the library's back (casting the const away) must call
 * mimc3cu_image_invalidate afterwards, or the exact-FP32 matcher works with the old table. */
const float *mimc3cu_image_ptr(mimc3cu_ctx *ctx, int32_t handle);
int mimc3cu_image_invalidate(mimc3cu_ctx *ctx, int32_t handle);

/* GMA_float_conv2, MIMC_module.c:2517-2585.  dst is updated IN PLACE with the
 * reference's stale-border semantics.  kernel is (kh, kw) row-major
 * float32 on the host; kh, kw in {1, 3}. */
int mimc3cu_conv2(mimc3cu_ctx *ctx, int32_t src, const float *kernel, int32_t kh, int32_t kw, int32_t dst);

/* ---- nodes and pivots ----------------------------------------------------------- */
/* get_uv_pivot, MIMC_module.c:543-602 (host code in the reference too: libm trig,
 * ragged output).  Multi-threaded host implementation, bit-identical including the
 * |cos|>|sin| normalisation quirk.  Two-call protocol: pass piv == NULL to obtain the
 * offsets and the total; then call again with piv sized total*2. Returns total (<0: error). */
int64_t mimc3cu_get_uv_pivot(const double *xyuvav, int32_t n, float dt, float mpp, float AW_SF, float AW_CRE,
                             int32_t ocw, int32_t H, int32_t W, int32_t *off, int32_t *piv);

/* Upload the node list (columns 2,3 truncated to int32 as at MIMC_module.c:822-823). */
int mimc3cu_set_nodes(mimc3cu_ctx *ctx, const double *xyuvav, int32_t n);
/* Upload a CSR pivot set into slot `slot` (0..7); the library keeps it in HBM. */
int mimc3cu_set_pivots(mimc3cu_ctx *ctx, int32_t slot, const int32_t *off, const int32_t *piv, int32_t n);

/* ---- the matcher ---------------------------------------------------------------- */
/* matching_ncc_dlc_2 (+extract_refchip, extract_sarea, investigate_valid_grid,
 * find_ncc_peak), MIMC_module.c:605-890, for all nodes set by mimc3cu_set_nodes.
 *   ref_img/search_img : image handles (i0,i1 for the forward pass; swapped for the
 *                        "swapped forward" pass, MIMC_main.c:267,284)
 *   offset[2]          : CP integer offset added to the node position in the search image
 *   pivot_slot, sign   : pivot set; sign=-1 reproduces main's in-place negation (:272-279)
 *   negate_duv         : multiply du,dv of the result by -1 (main does this to the swapped
 *                        pass on the host, :289-293)
 * Outputs are DEVICE pointers (any may be NULL except dp): dp (n,3) f32, peak (n,2) i32
 * integer peak relative to the search-area centre, ncell (n) i32 number of NCC cells the
 * reference algorithm evaluates (E in SURVEY.md 8d). Asynchronous on the context stream. */
int mimc3cu_match_async(mimc3cu_ctx *ctx, int32_t ref_img, int32_t search_img, const int32_t *offset,
                        int32_t pivot_slot, int32_t sign, int32_t ocw, int32_t negate_duv,
                        float *dp_dev, int32_t *peak_dev, int32_t *ncell_dev);
/* Same, synchronous, results copied to HOST buffers. */
int mimc3cu_match(mimc3cu_ctx *ctx, int32_t ref_img, int32_t search_img, const int32_t *offset,
                  int32_t pivot_slot, int32_t sign, int32_t ocw, int32_t negate_duv,
                  float *dp_host, int32_t *peak_host, int32_t *ncell_host);

/* Matcher selection.  The library has two CUDA implementations of the cell evaluator with
 * identical results: the general FP64-accumulating kernel (any float image) and the exact-FP32
 * kernel (images whose pixels are non-negative multiples of 1/8, i.e. everything
 * GMA_float_load_tiff and GMA_float_conv2 produce; it uses per-image summed-area tables).
 * mode 0 = automatic per call (default; env MIMC3CU_MATCHER=v1|v2 overrides at context creation),
 * 1 = always the general kernel, 2 = require the exact-FP32 kernel (calls outside its class fail). */
int mimc3cu_set_matcher(mimc3cu_ctx *ctx, int32_t mode);
/* 1 / 2: which of the two the last mimc3cu_match* call used (0: none yet). */
int mimc3cu_last_matcher(const mimc3cu_ctx *ctx);
/* Image class as seen by the matcher: exact_class (0/1), fractional bits (0 or 3), maximum. */
int mimc3cu_image_class(mimc3cu_ctx *ctx, int32_t handle, int32_t *exact_class, int32_t *frac_bits, float *max_value);

/* find_ncc_peak on explicit chips (the CP stage's call shape, MIMC_module.c:351,369):
 * `count` independent problems; refchips (count, S, S), sareas (count, D, D) float32 on
 * the HOST, one shared pivot list (P,2).  Results to host: uvncc (count,3), peak (count,2). */
int mimc3cu_find_ncc_peak_batch(mimc3cu_ctx *ctx, const float *refchips, int32_t S, const float *sareas,
                                int32_t D, int32_t count, const int32_t *piv, int32_t P,
                                float *uvncc_host, int32_t *peak_host, int32_t *ncell_host);

/* The whole multi-match of MIMC_main.c:261-350: 4 chip sizes x {forward, swapped} on the
 * raw pair, then on the three conv2-filtered pairs => 32 dp arrays, stored attempt-major
 * in dp_dev (32, n, 3) on the DEVICE.  Pivots for the 4 chip sizes must be in slots 0..3.
 * ncell_dev (32, n) optional. i0c/i1c are scratch images for the filtered pair. */
int mimc3cu_multimatch_async(mimc3cu_ctx *ctx, int32_t i0, int32_t i1, int32_t i0c, int32_t i1c,
                             const int32_t *offset, const mimc3cu_params *p, float *dp_dev, int32_t *ncell_dev);
/* Same with the integer peaks of every attempt, peak_dev (32, n, 2) on the device (may be NULL): what the reference
 * keeps internal and the parity checks compare against the oracle. */
int mimc3cu_multimatch_diag_async(mimc3cu_ctx *ctx, int32_t i0, int32_t i1, int32_t i0c, int32_t i1c,
                                  const int32_t *offset, const mimc3cu_params *p, float *dp_dev, int32_t *ncell_dev,
                                  int32_t *peak_dev);

/* ---- control points --------------------------------------------------------------- */
/* get_offset_image, MIMC_module.c:33-492: integer offset between the two images from slow
 * ("control point") nodes.  xyuvav (n,6) on the host; kernels = the three filters of
 * MIMC_main.c:175-196 (1x3, 3x1, 3x3, row-major); seed = what the reference passes to srand()
 * (time(NULL), MIMC_module.c:516) -- the candidate permutation uses glibc rand() like the
 * reference, so the same seed gives the same control points.
 * Outputs: offset[2]; flag_cp[n] (caller-zeroed; set to 1 for nodes used as CPs); *result = 1
 * (ok) or -1 (not enough control points: the reference's return value); num_cp_found optional.
 * Returns 0 unless a CUDA error occurred. */
int mimc3cu_get_offset_image(mimc3cu_ctx *ctx, int32_t i0, int32_t i1, const double *xyuvav, int32_t n,
                             const mimc3cu_params *p, const float *k1x3, const float *k3x1, const float *k3x3,
                             uint32_t seed, int32_t *offset, uint8_t *flag_cp, int32_t *result, int32_t *num_cp_found);

/* ---- postprocess ---------------------------------------------------------------- */
/* calc_mean_var_num_dp_cluster, MIMC_module.c:994-1194: dp_dev (num_dp, n, 3) ->
 * mvn_dev (n, num_dp, 5) [mean u, mean v, var u, var v, support], ncl_dev (n). */
int mimc3cu_cluster_async(mimc3cu_ctx *ctx, const float *dp_dev, int32_t n, int32_t num_dp,
                          float *mvn_dev, int32_t *ncl_dev);

/* mimc2_postprocess, MIMC_module.c:893-991 (cluster -> dpf0 -> dpf1 sweeps -> 3x3
 * smoothing -> snap -> pseudosmoothing -> pack): dp_dev (num_dp, n, 3) on the device,
 * xyuvav (n,6) on the host -> planes_dev 5 x (dimy, dimx) f32 on the device
 * [du, dv, var u, var v, support].  stats (optional, host, 4 ints): dpf1 sweeps,
 * pseudosmoothing sweeps, holes after dpf0, nodes changed by pseudosmoothing. */
int mimc3cu_postprocess(mimc3cu_ctx *ctx, const float *dp_dev, const double *xyuvav, const mimc3cu_params *p,
                        float *planes_dev, int32_t *stats);

/* ---- postprocess of one band of node rows (multi-GPU) --------------------------------------
 * The node grid is split into contiguous bands of rows, one per GPU/process.  Every iterative
 * stage reads neighbours up to mimc3cu_band_halo() rows away (3 rows for get_dpf1, 5 for the
 * pseudosmoothing with the reference's radii), so each band keeps that many halo rows of its
 * neighbours and refreshes them after every committed sweep.  The library runs the reference's
 * sweep control unchanged and calls back into the host for the three communication steps; the
 * host implements them with whatever transport it has (mimc3_b200/bands.py: torch.distributed,
 * NCCL over NVLink on GPUs).  Arrays are DEVICE pointers to the band's LOCAL arrays: local row 0 is
 * global row max(0, own_row0 - halo); rows are dimx elements of elem_bytes bytes.
 *   halo_exchange   overwrite my halo rows with the neighbours' current owned rows
 *   halo_or_reduce  uint8 flags: OR my halo rows into the owners' rows (owners' rows are final after it)
 *   allreduce_sum   int32 counters on the HOST, summed over all bands, in place
 * Each returns 0 on success.  All bands must call mimc3cu_postprocess_band collectively. */
typedef struct mimc3cu_band_comm {
    void *user;
    int (*halo_exchange)(void *user, void *const *arrays, const int32_t *elem_bytes, int32_t count);
    int (*halo_or_reduce)(void *user, void *flags_u8);
    int (*allreduce_sum)(void *user, int32_t *vals, int32_t count);
} mimc3cu_band_comm;
int mimc3cu_band_halo(const mimc3cu_params *p);
/* dp_dev (num_dp, own_rows*dimx, 3) of the OWNED nodes; xyuvav = the GLOBAL (dimy*dimx, 6) matrix on
 * the host; p->dimx/dimy = the GLOBAL grid; planes_dev 5 x (own_rows, dimx).  comm == NULL is only
 * valid for the single band [0, dimy) (== mimc3cu_postprocess) or with a communicator attached to the context
 * (mimc3cu_comm_init_*, below). stats are global. */
int mimc3cu_postprocess_band(mimc3cu_ctx *ctx, const float *dp_dev, const double *xyuvav, const mimc3cu_params *p,
                             int32_t own_row0, int32_t own_rows, const mimc3cu_band_comm *comm, float *planes_dev,
                             int32_t *stats);

/* ---- the library's own band communicator (NCCL over NVLink / NVSwitch, csrc/comm.cu) ----------------------
 * With a communicator attached to the context, mimc3cu_postprocess_band(comm = NULL) does the three
 * communication steps itself: grouped ncclSend/ncclRecv of the halo rows between band neighbours, the OR of the
 * scattered dirty flags, and ncclAllReduce of the sweep counters, all enqueued on the context's stream (one host
 * synchronisation per sweep: the read-back of the reduced counters that drives the reference's loop control).
 * Band r of the node grid lives on rank r.  NCCL is bound at run time (libnccl.so.2; MIMC3CU_NCCL_LIB overrides).
 *   one process per GPU : rank 0 calls mimc3cu_comm_unique_id, the caller distributes the 128 bytes (e.g. over
 *                         torch.distributed or MPI), every rank calls mimc3cu_comm_init_rank;
 *   one process, n GPUs : mimc3cu_comm_init_all over n contexts (the drop-in CLI, MIMC3CU_DEVICES=n); the bands'
 *                         calls must then come from n host threads. */
int mimc3cu_comm_unique_id(void *id128);
int mimc3cu_comm_init_rank(mimc3cu_ctx *ctx, const void *id128, int32_t rank, int32_t world);
int mimc3cu_comm_init_all(mimc3cu_ctx **ctxs, int32_t n);
void mimc3cu_comm_destroy(mimc3cu_ctx *ctx);
/* rank / world and the number of halo exchanges / all-reduces issued so far (any pointer may be NULL); non-zero
 * without a communicator. */
int mimc3cu_comm_info(const mimc3cu_ctx *ctx, int32_t *rank, int32_t *world, int64_t *exchanges, int64_t *allreduces);
/* Device time of the collectives (CUDA events around them on the context's stream) since the last call: summed
 * milliseconds of the halo exchanges (incl. the dirty-flag OR) and of the counter all-reduces; `on` switches the
 * bracketing on / off for what follows.  Synchronises the stream. */
int mimc3cu_comm_timing(mimc3cu_ctx *ctx, int32_t on, double *exchange_ms, double *allreduce_ms);
/* The final gather: rank r contributes bytes[r] bytes from the device buffer `send`; on `root` the device buffer
 * `recv` receives them back to back in rank order.  Asynchronous on the context's stream. */
int mimc3cu_comm_gather(mimc3cu_ctx *ctx, const void *send, const int64_t *bytes, void *recv, int32_t root);

/* Intermediate fields of the last mimc3cu_postprocess call, copied to the host for the
 * differential tests: which = 0 dpf0 (i32), 1 dpf1 ids (i32), 2 dpf1 dx (f32), 3 dpf1 dy,
 * 4 pseudosmoothing ids, 5 ps dx, 6 ps dy, 7 ncl (i32). `host` holds n 4-byte values
 * (n = the owned nodes of the band in band mode). */
int mimc3cu_postprocess_stage(mimc3cu_ctx *ctx, int32_t which, void *host);

/* du, dv -> -du, -dv on a device (n, 3) result: what main does to the swapped passes on the host (MIMC_main.c:289-293). */
int mimc3cu_dp_negate_uv_async(mimc3cu_ctx *ctx, float *dp_dev, int32_t n);

/* main()'s tail, MIMC_main.c:356-402: mean of the non-NaN du,dv removed (sequential
 * float sums), px -> m/yr, vy sign flip, sqrt of the variances. In place on planes_dev. */
int mimc3cu_finalize(mimc3cu_ctx *ctx, float *planes_dev, const mimc3cu_params *p, float *du_cp, float *dv_cp);

/* ---- measurement helper --------------------------------------------------------------- */
/* Sustained FP32 FMA throughput of this GPU's CUDA cores (TFLOP/s, best of 4 timed
 * launches of a register-resident FMA chain): the matcher's roofline denominator. */
int mimc3cu_fp32_peak(mimc3cu_ctx *ctx, double *tflops, double *ms);

/* Per-kernel-family CUDA-event timing on the context stream: family 0 = matcher kernel,
 * 1 = conv2, 2 = postprocess (whole call). timing_read synchronises, returns the summed
 * milliseconds and launch counts since the last read, and clears the timers. */
int mimc3cu_timing_enable(mimc3cu_ctx *ctx, int on);
int mimc3cu_timing_read(mimc3cu_ctx *ctx, double *ms3, int64_t *counts3);

/* ---- device memory helpers (so a plain-C host needs no CUDA headers) -------------- */
int mimc3cu_malloc(mimc3cu_ctx *ctx, size_t bytes, void **dev);
int mimc3cu_free(mimc3cu_ctx *ctx, void *dev);
int mimc3cu_memcpy_d2h(mimc3cu_ctx *ctx, void *host, const void *dev, size_t bytes);
int mimc3cu_memcpy_h2d(mimc3cu_ctx *ctx, void *dev, const void *host, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* MIMC3CU_H */
