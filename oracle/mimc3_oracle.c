/* CPU restatement of MIMC3's per-grid-node matching path -- the parity oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see mimc3_oracle.h).  New code written from the
 * behaviour of the reference; every function cites the reference lines it restates and
 * keeps the reference's evaluation order and float/double mix so results are bit-equal.
 * Compile WITHOUT -march=native / -ffast-math and with -ffp-contract=off.
 *
 * "Zero-initialised allocations" semantics are built in: memory the
 * reference leaves unwritten (last row/column of the search area and of cmap, borders of
 * the conv2 output) is 0.0 here.
 */
#include "mimc3_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#define MIN_DN 0.0000000001 /* MIMC_module.c:21 */

int orc_num_threads(void) { return omp_get_max_threads(); }

/* --------------------------------------------------------------------------------- */
/* get_uv_pivot, MIMC_module.c:543-602                                                 */
/* --------------------------------------------------------------------------------- */
static int32_t pivots_of_node(const double *row, float dt, float mpp, float AW_SF, float AW_CRE, int32_t ocw,
                              int32_t H, int32_t W, int32_t *dst /* may be NULL: count only */) {
    float u = 0.0, v = 0.0, incr_u, incr_v, norm_incr, theta;
    double length_pivot;
    int32_t num_pivot, k;
    theta = atan2(row[5], row[4]);                                   /* :560 */
    incr_u = cos(theta);
    incr_v = sin(theta);
    if (fabs(incr_u) > fabs(incr_v)) {                               /* :563 */
        incr_u = incr_u / fabs(incr_u);
        incr_v = incr_v / fabs(incr_u); /* divides by the already-normalised incr_u (quirk) */
    } else {
        incr_u = incr_u / fabs(incr_v);
        incr_v = incr_v / fabs(incr_v);
    }
    norm_incr = sqrt(incr_u * incr_u + incr_v * incr_v);             /* :572 */
    length_pivot = sqrt(row[4] * row[4] + row[5] * row[5]) / mpp / 365 * dt * AW_SF + AW_CRE + 1; /* :573 */
    num_pivot = 0;
    while (u + (float)(row[2]) - (float)ocw > 0 && u + (float)(row[2]) + (float)ocw < (float)(W - 1) &&
           v + (float)(row[3]) - (float)ocw > 0 && v + (float)(row[3]) + (float)ocw < (float)(H - 1) &&
           length_pivot > (double)(norm_incr * (double)num_pivot)) { /* :576-580 */
        num_pivot++;
        u += incr_u;
        v += incr_v;
    }
    if (dst) {
        u = 0.0; v = 0.0;
        if (num_pivot > 0) { dst[0] = 0; dst[1] = 0; }               /* :590-591 */
        for (k = 1; k < num_pivot; k++) {
            u += incr_u; v += incr_v;
            dst[2 * k] = (int32_t)(u + 0.5);                         /* :596 */
            dst[2 * k + 1] = -(int32_t)(v + 0.5);                    /* :597 */
        }
    }
    return num_pivot;
}

int64_t orc_get_uv_pivot(const double *xyuvav, int32_t n, float dt, float mpp, float AW_SF, float AW_CRE,
                         int32_t ocw, int32_t H, int32_t W, int32_t *csr_off, int32_t *piv, int64_t cap) {
    int64_t tot = 0;
    for (int32_t g = 0; g < n; g++) {
        csr_off[g] = (int32_t)tot;
        tot += pivots_of_node(xyuvav + 6 * (size_t)g, dt, mpp, AW_SF, AW_CRE, ocw, H, W, NULL);
    }
    csr_off[n] = (int32_t)tot;
    if (tot > cap) return -tot;
#pragma omp parallel for schedule(static)
    for (int32_t g = 0; g < n; g++)
        pivots_of_node(xyuvav + 6 * (size_t)g, dt, mpp, AW_SF, AW_CRE, ocw, H, W, piv + 2 * (size_t)csr_off[g]);
    return tot;
}

/* --------------------------------------------------------------------------------- */
/* find_ncc_peak, MIMC_module.c:647-801 (+ investigate_valid_grid :605-644)            */
/* --------------------------------------------------------------------------------- */
void orc_find_ncc_peak(const float *refchip, int32_t S, const float *sarea, int32_t Dy2, int32_t Dx2,
                       const int32_t *piv, int32_t P, float *uvncc, int32_t *peak, int32_t *ncell) {
    const float N_A_N = sqrt(-1.0);
    int dx2 = Dx2 / 2, dy2 = Dy2 / 2, ocw = S / 2;
    int uv_peak[2] = {dx2, dy2};
    int32_t ncells = 0;
    uvncc[0] = 0.0; uvncc[1] = 0.0; uvncc[2] = -2.0;
    float *cmap = (float *)calloc((size_t)Dy2 * Dx2, sizeof(float));
    for (int c1 = -dx2; c1 < dx2; c1++)                               /* :677-681: last row/col stay 0 */
        for (int c2 = -dy2; c2 < dy2; c2++) cmap[(size_t)(c2 + dy2) * Dx2 + (c1 + dx2)] = -2.0;

    /* investigate_valid_grid :605-644 */
    int32_t inv_ref = 0, inv_sa = 0;
    for (int i = 0; i < S * S; i++) if (refchip[i] < MIN_DN) inv_ref++;
    for (int64_t i = 0; i < (int64_t)Dy2 * Dx2; i++) if (sarea[i] < MIN_DN) inv_sa++;
    float numpx_ref = (float)(S * S), numpx_sa = (float)(Dy2 * Dx2), max_ratio = 0.8;
    if ((float)inv_ref / numpx_ref > max_ratio || (float)inv_sa / numpx_sa > max_ratio) {
        uvncc[0] = N_A_N; uvncc[1] = N_A_N; uvncc[2] = -3;            /* :684-688 */
        if (peak) { peak[0] = 0; peak[1] = 0; }
        if (ncell) *ncell = 0;
        free(cmap);
        return;
    }

    for (int32_t ip = 0; ip < P; ip++) {                              /* :691 */
        int pivot[2] = {piv[2 * ip] + dx2, piv[2 * ip + 1] + dy2};
        int duv[2] = {-1, -1};
        float nccmax = -2;
        int flag_newncc = 1;
        while ((duv[0] != 0 || duv[1] != 0) && flag_newncc != 0) {    /* :699 */
            duv[0] = 0; duv[1] = 0;
            if (pivot[0] - ocw <= 1 || pivot[0] + ocw >= Dx2 - 1 || pivot[1] - ocw <= 1 || pivot[1] + ocw >= Dy2 - 1) break;
            flag_newncc = 0;
            for (int c1 = -1; c1 <= 1; c1++) {
                for (int c2 = -1; c2 <= 1; c2++) {
                    float *cell = &cmap[(size_t)(pivot[1] + c2) * Dx2 + (pivot[0] + c1)];
                    if (*cell < -1.0) {                                /* :713 */
                        flag_newncc++; ncells++;
                        int32_t nsample = 0;
                        double sy = 0, sx = 0, sxx = 0, sxy = 0, syy = 0;
                        for (int c3 = -ocw; c3 <= ocw; c3++) {        /* column outer, row inner :719-721 */
                            for (int c4 = -ocw; c4 <= ocw; c4++) {
                                float r = refchip[(c4 + ocw) * S + (c3 + ocw)];
                                float s = sarea[(size_t)(pivot[1] + c2 + c4) * Dx2 + (pivot[0] + c1 + c3)];
                                if (r >= MIN_DN && s >= MIN_DN) {      /* null exclusion :723 */
                                    nsample++;
                                    sy += s; sx += r;
                                    sxx += r * r; syy += s * s; sxy += r * s; /* float products, double sums */
                                }
                            }
                        }
                        *cell = (float)((nsample * sxy - sx * sy) / sqrt((nsample * sxx - sx * sx) * (nsample * syy - sy * sy))); /* :734 */
                    }
                    if (*cell > nccmax) { nccmax = *cell; duv[0] = c1; duv[1] = c2; } /* :736-741 */
                }
            }
            pivot[0] += duv[0]; pivot[1] += duv[1];
        }
        if (nccmax > uvncc[2]) { uv_peak[0] = pivot[0]; uv_peak[1] = pivot[1]; uvncc[2] = nccmax; } /* :747-752 */
    }

    /* 3x3 quadratic fit :757-788 */
    float ncc9[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
        ncc9[r * 3 + c] = cmap[(size_t)(uv_peak[1] - 1 + r) * Dx2 + (uv_peak[0] - 1 + c)];
    double cp[6];
    cp[0] = 6 * ncc9[0] - 12 * ncc9[1] + 6 * ncc9[2] + 6 * ncc9[3] - 12 * ncc9[4] + 6 * ncc9[5] + 6 * ncc9[6] - 12 * ncc9[7] + 6 * ncc9[8];
    cp[1] = 9 * ncc9[0] - 9 * ncc9[2] - 9 * ncc9[6] + 9 * ncc9[8];
    cp[2] = 6 * ncc9[0] + 6 * ncc9[1] + 6 * ncc9[2] - 12 * ncc9[3] - 12 * ncc9[4] - 12 * ncc9[5] + 6 * ncc9[6] + 6 * ncc9[7] + 6 * ncc9[8];
    cp[3] = -6 * ncc9[0] + 6 * ncc9[2] - 6 * ncc9[3] + 6 * ncc9[5] - 6 * ncc9[6] + 6 * ncc9[8];
    cp[4] = -6 * ncc9[0] - 6 * ncc9[1] - 6 * ncc9[2] + 6 * ncc9[6] + 6 * ncc9[7] + 6 * ncc9[8];
    cp[5] = -4 * ncc9[0] + 8 * ncc9[1] - 4 * ncc9[2] + 8 * ncc9[3] + 20 * ncc9[4] + 8 * ncc9[5] - 4 * ncc9[6] + 8 * ncc9[7] - 4 * ncc9[8];
    for (int k = 0; k < 6; k++) cp[k] /= 36;
    uvncc[0] = -2 * cp[2] * cp[3] + cp[1] * cp[4];
    uvncc[1] = -2 * cp[0] * cp[4] + cp[1] * cp[3];
    uvncc[0] /= 4 * cp[0] * cp[2] - cp[1] * cp[1];
    uvncc[1] /= 4 * cp[0] * cp[2] - cp[1] * cp[1];
    uvncc[0] += (float)(uv_peak[0] - dx2);
    uvncc[1] += (float)(uv_peak[1] - dy2);
    if (peak) { peak[0] = uv_peak[0] - dx2; peak[1] = uv_peak[1] - dy2; }
    if (ncell) *ncell = ncells;
    free(cmap);
}

/* --------------------------------------------------------------------------------- */
/* matching_ncc_dlc_2 :805-842, extract_refchip :845-855, extract_sarea :857-890       */
/* --------------------------------------------------------------------------------- */
void orc_match(const float *i0, const float *i1, int32_t H, int32_t W, const double *xyuvav, int32_t n,
               const int32_t *offset, const int32_t *csr_off, const int32_t *piv, int32_t sign, int32_t ocw,
               float *out, int32_t *peak, int32_t *ncell) {
    const int S = 2 * ocw + 1;
#pragma omp parallel
    {
        float *refchip = (float *)calloc((size_t)S * S, sizeof(float));
#pragma omp for schedule(dynamic)
        for (int32_t g = 0; g < n; g++) {
            int32_t uv0[2] = {(int32_t)xyuvav[6 * (size_t)g + 2], (int32_t)xyuvav[6 * (size_t)g + 3]};
            for (int c2 = -ocw; c2 <= ocw; c2++)                      /* no bounds check :852 */
                for (int c1 = -ocw; c1 <= ocw; c1++)
                    refchip[(c2 + ocw) * S + (c1 + ocw)] = i0[(size_t)(uv0[1] + c2) * W + (uv0[0] + c1)];
            uv0[0] += offset[0]; uv0[1] += offset[1];
            int32_t P = csr_off[g + 1] - csr_off[g];
            float uvncc[3]; int32_t pk[2] = {0, 0}, nc = 0;
            if (P <= 0) { /* reference: undefined behaviour (:589-591); defined here as "nothing evaluable" */
                uvncc[0] = sqrt(-1.0); uvncc[1] = uvncc[0]; uvncc[2] = -2.0;
            } else {
                int32_t *pv = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)P);
                for (int32_t k = 0; k < 2 * P; k++) pv[k] = sign * piv[2 * (size_t)csr_off[g] + k];
                int dx2 = abs(pv[2 * (P - 1)]) + ocw + 2, dy2 = abs(pv[2 * (P - 1) + 1]) + ocw + 2; /* :863-866 */
                int Dx2 = dx2 * 2 + 1, Dy2 = dy2 * 2 + 1;
                float *sarea = (float *)calloc((size_t)Dx2 * Dy2, sizeof(float));
                for (int c2 = -dy2; c2 < dy2; c2++) {                 /* `<`: last row/col unwritten = 0 (H1) */
                    int cv = uv0[1] + c2;
                    if (cv < 0 || cv >= H) continue;
                    for (int c1 = -dx2; c1 < dx2; c1++) {
                        int cu = uv0[0] + c1;
                        if (cu >= 0 && cu < W) sarea[(size_t)(c2 + dy2) * Dx2 + (c1 + dx2)] = i1[(size_t)cv * W + cu];
                    }
                }
                orc_find_ncc_peak(refchip, S, sarea, Dy2, Dx2, pv, P, uvncc, pk, &nc);
                free(sarea); free(pv);
            }
            out[3 * (size_t)g] = uvncc[0]; out[3 * (size_t)g + 1] = uvncc[1]; out[3 * (size_t)g + 2] = uvncc[2];
            if (peak) { peak[2 * (size_t)g] = pk[0]; peak[2 * (size_t)g + 1] = pk[1]; }
            if (ncell) ncell[g] = nc;
        }
        free(refchip);
    }
}

/* --------------------------------------------------------------------------------- */
/* GMA_float_conv2, MIMC_module.c:2517-2585                                            */
/* --------------------------------------------------------------------------------- */
void orc_conv2(const float *in, int32_t H, int32_t W, const float *kernel, int32_t kh, int32_t kw, float *out) {
    const float N_A_N = sqrt(-1);
    int32_t ocwx = kw / 2, ocwy = kh / 2;
    for (int32_t r = ocwy; r < H - ocwy; r++) {
        for (int32_t c = ocwx; c < W - ocwx; c++) {
            float sum_dn = 0;
            for (int32_t a = 0; a < kh; a++)
                for (int32_t b = 0; b < kw; b++) {
                    float px = in[(size_t)(r + a - ocwy) * W + (c + b - ocwx)];
                    float dn_in = (int32_t)(px + 0.5) ? px : N_A_N;    /* :2548 */
                    sum_dn += dn_in * kernel[a * kw + b];
                }
            out[(size_t)r * W + c] = sum_dn;
        }
    }
    float dn_min = 1e+37;
    for (size_t i = 0; i < (size_t)H * W; i++) if (out[i] < dn_min) dn_min = out[i]; /* whole buffer :2558 */
    for (int32_t r = ocwy; r < H - ocwy; r++)
        for (int32_t c = ocwx; c < W; c++) {                          /* right border included :2572 */
            float *o = &out[(size_t)r * W + c];
            if (isnan(*o)) *o = 0; else *o -= dn_min - 1;
        }
}

/* --------------------------------------------------------------------------------- */
/* calc_mean_var_num_dp_cluster :994-1130, cluster_euclidian :1133-1180, mark_row :1183 */
/* --------------------------------------------------------------------------------- */
static void mark_row(int32_t row, uint8_t id, uint8_t *ids, const uint8_t *dist, int32_t k) {
    for (int32_t c = 0; c < k; c++)
        if (dist[row * k + c] && ids[c] == 0) { ids[c] = id; mark_row(c, id, ids, dist, k); }
}

void orc_cluster(const float *dp, int32_t n, int32_t num_dpoi, float *mvn, int32_t *ncl) {
    const float min_ncc = 0.1, min_dist = 0.5;
    const float min_dist_sq = min_dist * min_dist;
#pragma omp parallel
    {
        float *stk = (float *)malloc(sizeof(float) * 2 * (size_t)num_dpoi);
        uint8_t *ids = (uint8_t *)malloc((size_t)num_dpoi);
        uint8_t *dist = (uint8_t *)malloc((size_t)num_dpoi * num_dpoi);
        float *sx = (float *)malloc(sizeof(float) * 4 * (size_t)num_dpoi);
        float *sy = sx + num_dpoi, *sxx = sy + num_dpoi, *syy = sxx + num_dpoi;
        int *ns = (int *)malloc(sizeof(int) * (size_t)num_dpoi);
#pragma omp for schedule(static)
        for (int32_t g = 0; g < n; g++) {
            int32_t k = 0;
            for (int32_t a = 0; a < num_dpoi; a++) {
                const float *d = dp + ((size_t)a * n + g) * 3;
                if (d[2] > min_ncc) { stk[2 * k] = d[0]; stk[2 * k + 1] = d[1]; k++; } /* :1048 */
            }
            for (int32_t i = 0; i < k; i++) ids[i] = 0;
            for (int32_t i = 0; i < k; i++)
                for (int32_t j = i; j < k; j++) {
                    float dx = stk[2 * j] - stk[2 * i], dy = stk[2 * j + 1] - stk[2 * i + 1];
                    uint8_t near = (dx * dx + dy * dy < min_dist_sq) ? 1 : 0; /* :1156 */
                    dist[i * k + j] = near; dist[j * k + i] = near;
                }
            uint8_t id_curr = 0;
            for (int32_t i = 0; i < k; i++)
                if (ids[i] == 0) { id_curr++; mark_row(i, id_curr, ids, dist, k); } /* :1170-1176 */
            int32_t max_id = 0;
            for (int32_t i = 0; i < k; i++) if (max_id < ids[i]) max_id = ids[i];
            for (int32_t c = 0; c < max_id; c++) { sx[c] = 0; sy[c] = 0; sxx[c] = 0; syy[c] = 0; ns[c] = 0; }
            for (int32_t i = 0; i < k; i++) {
                /* A candidate with a NaN displacement is never labelled (its self-distance is
                 * NaN, :1154-1156) and the reference then writes sx[-1] (H11, out-of-bounds).
                 * Defined here as: it consumes a cluster id and contributes to no cluster. */
                if (ids[i] == 0) continue;
                int c = ids[i] - 1;
                sx[c] += stk[2 * i]; sy[c] += stk[2 * i + 1];
                sxx[c] += stk[2 * i] * stk[2 * i]; syy[c] += stk[2 * i + 1] * stk[2 * i + 1];
                ns[c] += 1;
            }
            float *o = mvn + (size_t)g * num_dpoi * 5;
            for (int32_t c = 0; c < max_id; c++) {                     /* :1097-1104 */
                o[5 * c] = sx[c] / (float)ns[c];
                o[5 * c + 1] = sy[c] / (float)ns[c];
                o[5 * c + 2] = sxx[c] / (float)ns[c] - o[5 * c] * o[5 * c];
                o[5 * c + 3] = syy[c] / (float)ns[c] - o[5 * c + 1] * o[5 * c + 1];
                o[5 * c + 4] = (float)ns[c] / (float)num_dpoi;
            }
            ncl[g] = max_id;
        }
        free(stk); free(ids); free(dist); free(sx); free(ns);
    }
}

/* --------------------------------------------------------------------------------- */
/* get_ruv_neighbor :1266-1327                                                         */
/* --------------------------------------------------------------------------------- */
int32_t orc_ruv_neighbor(const double *xyuvav, const orc_post_params *p, float radius, int32_t *ruv, int32_t cap) {
    int32_t cu = p->dimx / 2, cv = p->dimy / 2, k = 0;
    float cx = (float)xyuvav[6 * (size_t)cu + 0];
    float cy = (float)xyuvav[6 * (size_t)cv * p->dimx + 1];
    for (int32_t v = 0; v < p->dimy; v++)
        for (int32_t u = 0; u < p->dimx; u++) {
            float fx = (float)xyuvav[6 * (size_t)u + 0], fy = (float)xyuvav[6 * (size_t)v * p->dimx + 1];
            float dxy0 = fx - cx, dxy1 = fy - cy;
            float sq = dxy0 * dxy0 + dxy1 * dxy1;
            if (sq <= (radius * p->meter_per_spacing) * (radius * p->meter_per_spacing)) {
                if (k < cap) { ruv[2 * k] = u - cu; ruv[2 * k + 1] = v - cv; }
                k++;
            }
        }
    return k;
}

/* get_dpf0 :1224-1263 */
void orc_dpf0(const float *mvn, const int32_t *ncl, const orc_post_params *p, float min_ratio, int32_t *dpf0) {
    int32_t n = p->dimx * p->dimy;
    for (int32_t g = 0; g < n; g++) {
        dpf0[g] = -1;
        for (int32_t c = 0; c < ncl[g]; c++)
            if (mvn[((size_t)g * p->num_dp + c) * 5 + 4] > min_ratio) { dpf0[g] = c; break; }
    }
}

/* --------------------------------------------------------------------------------- */
/* get_dpf1 :1330-1718                                                                 */
/* --------------------------------------------------------------------------------- */
int32_t orc_dpf1(int32_t *dpf0, float *dpf_dx, float *dpf_dy, const int32_t *ruv, int32_t nruv,
                 const float *mvn, const int32_t *ncl, const double *xyuvav, const orc_post_params *p) {
    const int32_t dimx = p->dimx, dimy = p->dimy, n = dimx * dimy, K = p->num_dp;
    const float N_A_N = sqrt(-1.0);
    float *dxb = (float *)malloc(sizeof(float) * (size_t)n), *dyb = (float *)malloc(sizeof(float) * (size_t)n);
    float *noi = (float *)malloc(sizeof(float) * (size_t)n);
    float *vec = (float *)malloc(sizeof(float) * 7 * (size_t)nruv);
    for (int32_t g = 0; g < n; g++) {
        noi[g] = 1.0;
        if (dpf0[g] >= 0) {
            dpf_dx[g] = mvn[((size_t)g * K + dpf0[g]) * 5 + 0];
            dpf_dy[g] = mvn[((size_t)g * K + dpf0[g]) * 5 + 1];
        } else { dpf_dx[g] = N_A_N; dpf_dy[g] = N_A_N; }
        dxb[g] = N_A_N; dyb[g] = N_A_N;
    }
    int32_t NOI = 0, num_unprocessed = 1;
    for (int32_t thres_n = nruv - 1; thres_n >= 3; thres_n--) {        /* :1390 */
        float thres_weight = 0.5;
        float factor = 1.0 / 365.0 * p->dt / p->mpp;                   /* :1393 */
        while (num_unprocessed != 0 && thres_weight >= 0.5) {
            thres_weight -= 0.02;
            int32_t num_processed = 1;
            while (num_processed != 0) {
                NOI++;
                num_processed = 0;
                for (int32_t cv = 0; cv < dimy; cv++) for (int32_t cu = 0; cu < dimx; cu++) {
                    int32_t g = cv * dimx + cu;
                    if (!(isnan(dpf_dx[g] + dpf_dy[g]) && ncl[g] != 0)) continue; /* :1412 */
                    int32_t nn = 0;
                    float dpe[2];
                    dpe[0] = xyuvav[6 * (size_t)g + 4] * factor;
                    dpe[1] = -xyuvav[6 * (size_t)g + 5] * factor;
                    float mag_dpe = sqrt(dpe[0] * dpe[0] + dpe[1] * dpe[1]);
                    for (int32_t k = 0; k < nruv; k++) {
                        int32_t u = cu + ruv[2 * k], v = cv + ruv[2 * k + 1];
                        if (u < 0 || u >= dimx || v < 0 || v >= dimy) continue;
                        float dn0 = dpf_dx[v * dimx + u], dn1 = dpf_dy[v * dimx + u];
                        if (isnan(dn0 + dn1)) continue;                 /* :1430 */
                        float apv0 = (float)(xyuvav[6 * ((size_t)v * dimx + u) + 4]) * factor;
                        float apv1 = -(float)(xyuvav[6 * ((size_t)v * dimx + u) + 5]) * factor;
                        float *r = vec + 7 * nn;
                        r[0] = (float)ruv[2 * k]; r[1] = (float)ruv[2 * k + 1];
                        r[4] = sqrt(dn0 * dn0 + dn1 * dn1);
                        r[5] = sqrt(apv0 * apv0 + apv1 * apv1);
                        r[6] = noi[v * dimx + u];
                        r[3] = r[4] / sqrt(apv0 * apv0 + apv1 * apv1); /* float / double -> float :1444 */
                        nn++;
                    }
                    if (nn < thres_n) continue;                        /* :1453 */
                    float w_min = 1E+37, w_max = -1E+37, max_noi = 1.0;
                    int32_t id_w_max = 0, id_w_min = 0;
                    for (int32_t k = 0; k < nn; k++) {
                        float *r = vec + 7 * k;
                        float d0 = r[0], d1 = r[1];
                        float mag_dxy = sqrt(d0 * d0 + d1 * d1);
                        float w = (dpe[0] * d0 + dpe[1] * d1) / (mag_dpe * mag_dxy);
                        w = w > 0 ? w : -w;
                        if (w >= thres_weight) {
                            r[2] = w;
                            if (r[3] > w_max) { w_max = r[3]; id_w_max = k; }
                            if (r[3] < w_min) { w_min = r[3]; id_w_min = k; }
                        } else r[2] = 0.0;
                    }
                    vec[7 * id_w_max + 2] = 0.0; vec[7 * id_w_min + 2] = 0.0; /* :1496-1497 */
                    float sum_w = 0.0, sum_w_dp = 0.0, sum_w_dpe = 0.0, sum_noi = 0.0;
                    for (int32_t k = 0; k < nn; k++) {
                        float *r = vec + 7 * k;
                        float w2 = 1 / (1 + expf(-r[5] + 5)) / max_noi; /* :1514 */
                        sum_w += r[2];
                        sum_w_dp += r[2] * w2 * r[4] / r[6];
                        sum_w_dpe += r[2] * w2 * r[5] / r[6];
                        sum_noi += r[6];
                    }
                    if (sum_w >= 1.0) {                                 /* :1547 */
                        float factor_mag = sum_w_dp / sum_w_dpe;
                        dxb[g] = dpe[0] * factor_mag;
                        dyb[g] = dpe[1] * factor_mag;
                        noi[g] = sum_noi / nn + 1;
                        num_processed++;
                    }
                }
                for (int32_t g = 0; g < n; g++)                         /* Jacobi commit :1577-1589 */
                    if (!isnan(dxb[g]) && !isnan(dyb[g])) {
                        dpf_dx[g] = dxb[g]; dpf_dy[g] = dyb[g]; dxb[g] = N_A_N; dyb[g] = N_A_N;
                    }
            }
            num_unprocessed = 0;
            for (int32_t g = 0; g < n; g++)
                if ((isnan(dpf_dx[g]) || isnan(dpf_dy[g])) && ncl[g] != 0) num_unprocessed++;
        }
    }
    /* 3x3 box smoothing of the filled nodes :1623-1666 */
    for (int32_t cv = 1; cv < dimy - 1; cv++) for (int32_t cu = 1; cu < dimx - 1; cu++) {
        int32_t g = cv * dimx + cu;
        if (dpf0[g] < 0 && !isnan(dpf_dx[g] + dpf_dy[g])) {
            float num = 0.0, sdx = 0.0, sdy = 0.0;
            for (int dv = -1; dv <= 1; dv++) for (int du = -1; du <= 1; du++) {
                int32_t h = (cv + dv) * dimx + cu + du;
                if (!isnan(dpf_dx[h] + dpf_dy[h])) { sdx += dpf_dx[h]; sdy += dpf_dy[h]; num = num + 1; }
            }
            dxb[g] = sdx / num; dyb[g] = sdy / num;
        } else { dxb[g] = dpf_dx[g]; dyb[g] = dpf_dy[g]; }
    }
    for (int32_t cv = 1; cv < dimy - 1; cv++) for (int32_t cu = 1; cu < dimx - 1; cu++) {
        int32_t g = cv * dimx + cu;
        dpf_dx[g] = dxb[g]; dpf_dy[g] = dyb[g];
    }
    /* snap to the nearest cluster :1680-1706 */
    for (int32_t g = 0; g < n; g++) {
        if (!(dpf0[g] < 0 && ncl[g] != 0)) continue;
        float best = 1E+37; int32_t id = 0;
        for (int32_t c = 0; c < ncl[g]; c++) {
            float d0 = dpf_dx[g] - mvn[((size_t)g * K + c) * 5 + 0];
            float d1 = dpf_dy[g] - mvn[((size_t)g * K + c) * 5 + 1];
            float sq = d0 * d0 + d1 * d1;
            if (sq < best) { best = sq; id = c; }
        }
        dpf0[g] = id;
        dpf_dx[g] = mvn[((size_t)g * K + id) * 5 + 0];
        dpf_dy[g] = mvn[((size_t)g * K + id) * 5 + 1];
    }
    free(dxb); free(dyb); free(noi); free(vec);
    return NOI;
}

/* --------------------------------------------------------------------------------- */
/* GMA_double_inv :2430-2496 (Gauss-Jordan, no pivoting) and quadfit2 :2314-2409        */
/* --------------------------------------------------------------------------------- */
static void inv6(const double a[6][6], double I[6][6]) {
    double b[6][6];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { b[i][j] = a[i][j]; I[i][j] = (i == j) ? 1 : 0; }
    for (int c1 = 0; c1 < 5; c1++) {
        double pivot = b[c1][c1];
        for (int c2 = c1 + 1; c2 < 6; c2++) {
            double coeff = b[c2][c1] / pivot;
            for (int c3 = 0; c3 < 6; c3++) { b[c2][c3] -= b[c1][c3] * coeff; I[c2][c3] -= I[c1][c3] * coeff; }
        }
    }
    for (int c1 = 5; c1 >= 0; c1--) {
        double pivot = b[c1][c1];
        for (int c2 = c1 - 1; c2 >= 0; c2--) {
            double coeff = b[c2][c1] / pivot;
            for (int c3 = 5; c3 >= 0; c3--) { b[c2][c3] -= b[c1][c3] * coeff; I[c2][c3] -= I[c1][c3] * coeff; }
        }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) I[i][j] /= b[i][i];
}

/* value of the weighted quadratic LSQ surface at (0,0), for both displacement components */
static void quadfit2_at_origin(const int32_t *xy, const double *z, const double *w, int32_t nobs, double out[2]) {
    double (*A)[6] = (double (*)[6])malloc(sizeof(double) * 6 * (size_t)nobs);
    double N[6][6], IN[6][6], atwb[6], coeff[6];
    for (int32_t o = 0; o < nobs; o++) {
        double x = (double)xy[2 * o], y = (double)xy[2 * o + 1];
        A[o][0] = x * x; A[o][1] = x * y; A[o][2] = y * y; A[o][3] = x; A[o][4] = y; A[o][5] = 1;
    }
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
        N[r][c] = 0;
        for (int32_t o = 0; o < nobs; o++) N[r][c] += A[o][r] * w[o] * A[o][c];
    }
    inv6(N, IN);
    const double terms[6] = {0, 0, 0, 0, 0, 1};                        /* xyi = (0,0) :2384-2389 */
    for (int oc = 0; oc < 2; oc++) {
        for (int r = 0; r < 6; r++) {
            atwb[r] = 0;
            for (int32_t o = 0; o < nobs; o++) atwb[r] += A[o][r] * w[o] * z[2 * o + oc];
        }
        for (int r = 0; r < 6; r++) { coeff[r] = 0; for (int c = 0; c < 6; c++) coeff[r] += IN[r][c] * atwb[c]; }
        out[oc] = 0;
        for (int c = 0; c < 6; c++) out[oc] += terms[c] * coeff[c];
    }
    free(A);
}

/* --------------------------------------------------------------------------------- */
/* get_dpf_pseudosmoothing :1986-2312                                                  */
/* --------------------------------------------------------------------------------- */
int32_t orc_pseudosmooth(int32_t *dpf, float *dpf_dx, float *dpf_dy, const int32_t *ruv, int32_t nruv,
                         const float *mvn, const int32_t *ncl, const double *xyuvav, const orc_post_params *p) {
    const int32_t dimx = p->dimx, dimy = p->dimy, n = dimx * dimy, K = p->num_dp;
    const float N_A_N = sqrt(-1.0);
    enum { MAXS = 104 };
    uint8_t *layer[2] = {(uint8_t *)calloc((size_t)n, 1), (uint8_t *)calloc((size_t)n, 1)};
    uint8_t *stack[MAXS]; int32_t nstack = 0;
    float *bx = (float *)malloc(sizeof(float) * (size_t)n), *by = (float *)malloc(sizeof(float) * (size_t)n);
    int32_t *bid = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *uvn = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nruv);
    double *duvn = (double *)malloc(sizeof(double) * 2 * (size_t)nruv), *w = (double *)malloc(sizeof(double) * (size_t)nruv);
    double eigvel[2], ITM[4], duv_grid[2] = {0, 0}, duv_candidate[2] = {0, 0}, duv_interp[2];
    eigvel[0] = 1500.0 / 300.0; eigvel[1] = eigvel[0] / 3.0;
    int32_t id_cluster_closest = 0;
    for (int32_t g = 0; g < n; g++) {                                 /* :2031-2056 */
        int32_t id = dpf[g];
        layer[0][g] = (id < 0) ? 0 : (mvn[((size_t)g * K + id) * 5 + 4] >= 0.6 ? 0 : 1);
        bx[g] = N_A_N; by[g] = N_A_N; bid[g] = -1;
    }
    stack[0] = (uint8_t *)malloc((size_t)n); memcpy(stack[0], layer[0], (size_t)n); nstack = 1;
    int32_t NOI = 0; char any = 1, fluct = 0;
    while (NOI <= 100 && any) {                                       /* :2077 */
        uint8_t *mask = layer[NOI % 2], *next = layer[(NOI + 1) % 2];
        any = 0; NOI++;
        memset(next, 0, (size_t)n);
        for (int32_t cv = 0; cv < dimy; cv++) for (int32_t cu = 0; cu < dimx; cu++) {
            int32_t g = cv * dimx + cu;
            if (!mask[g]) continue;
            int32_t nn = 0;
            for (int32_t k = 0; k < nruv; k++) {
                int32_t u = cu + ruv[2 * k], v = cv + ruv[2 * k + 1];
                if (u >= 0 && u < dimx && v >= 0 && v < dimy && !isnan(dpf_dx[v * dimx + u]) && !isnan(dpf_dy[v * dimx + u])) {
                    uvn[2 * nn] = ruv[2 * k]; uvn[2 * nn + 1] = ruv[2 * k + 1];
                    duvn[2 * nn] = dpf_dx[v * dimx + u]; duvn[2 * nn + 1] = dpf_dy[v * dimx + u];
                    nn++;
                }
            }
            if (nn < 10) continue;                                    /* :2131 */
            double vx = xyuvav[6 * (size_t)g + 4], vy = xyuvav[6 * (size_t)g + 5];
            ITM[0] = (eigvel[1] * vx * vx + eigvel[0] * vy * vy) / ((eigvel[0] * eigvel[1]) * (vx * vx + vy * vy));
            ITM[1] = ((eigvel[0] - eigvel[1]) * vx * vy) / ((eigvel[0] * eigvel[1]) * (vx * vx + vy * vy));
            ITM[3] = (eigvel[1] * vy * vy + eigvel[0] * vx * vx) / ((eigvel[0] * eigvel[1]) * (vx * vx + vy * vy));
            for (int32_t k = 0; k < nn; k++)
                w[k] = exp(-(ITM[0] * uvn[2 * k] * uvn[2 * k] + 2 * ITM[1] * uvn[2 * k] * uvn[2 * k + 1] + ITM[3] * uvn[2 * k + 1] * uvn[2 * k + 1]));
            quadfit2_at_origin(uvn, duvn, w, nn, duv_interp);
            int32_t id = dpf[g], nc = ncl[g];
            double sq_min = 1E+37; int8_t upd = 0;
            for (int32_t c = 0; c < nc; c++) {
                double c0 = mvn[((size_t)g * K + c) * 5 + 0], c1 = mvn[((size_t)g * K + c) * 5 + 1];
                double sq = (duv_interp[0] - c0) * (duv_interp[0] - c0) + (duv_interp[1] - c1) * (duv_interp[1] - c1);
                if (sq < sq_min) { upd = 1; sq_min = sq; id_cluster_closest = c; }
            }
            if (upd) {  /* otherwise the reference reuses the previous node's values (:2182-2188) */
                duv_grid[0] = mvn[((size_t)g * K + id) * 5 + 0]; duv_grid[1] = mvn[((size_t)g * K + id) * 5 + 1];
                duv_candidate[0] = mvn[((size_t)g * K + id_cluster_closest) * 5 + 0];
                duv_candidate[1] = mvn[((size_t)g * K + id_cluster_closest) * 5 + 1];
            }
            if ((duv_grid[0] - duv_candidate[0]) * (duv_grid[0] - duv_candidate[0]) +
                (duv_grid[1] - duv_candidate[1]) * (duv_grid[1] - duv_candidate[1]) < 0.0001) continue;
            bx[g] = duv_candidate[0]; by[g] = duv_candidate[1]; bid[g] = id_cluster_closest;
            any = 1;
            for (int32_t k = 0; k < nn; k++) {
                int32_t h = (cv + uvn[2 * k + 1]) * dimx + cu + uvn[2 * k];
                if (stack[0][h]) next[h] = 1;                          /* :2205 */
            }
        }
        for (int32_t g = 0; g < n; g++)                                /* commit :2218-2233 */
            if (bid[g] >= 0) { dpf_dx[g] = bx[g]; dpf_dy[g] = by[g]; dpf[g] = bid[g]; bx[g] = N_A_N; by[g] = N_A_N; bid[g] = -1; }
        for (int32_t k = NOI - 1; k >= 0; k--)                          /* fluctuation check :2237-2260 */
            if (memcmp(stack[k], next, (size_t)n) == 0) { fluct = 1; break; }
        if (fluct) { NOI--; break; }
        stack[nstack] = (uint8_t *)malloc((size_t)n); memcpy(stack[nstack], next, (size_t)n); nstack++;
    }
    for (int32_t k = 0; k < nstack; k++) free(stack[k]);
    free(layer[0]); free(layer[1]); free(bx); free(by); free(bid); free(uvn); free(duvn); free(w);
    return NOI;
}

/* --------------------------------------------------------------------------------- */
/* mimc2_postprocess :893-991 and main()'s tail MIMC_main.c:356-402                    */
/* --------------------------------------------------------------------------------- */
void orc_postprocess(const float *dp, const double *xyuvav, const orc_post_params *p, float *planes) {
    const int32_t n = p->dimx * p->dimy, K = p->num_dp;
    const float N_A_N = sqrt(-1.0);
    float *mvn = (float *)calloc((size_t)n * K * 5, sizeof(float));
    int32_t *ncl = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *dpf0 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    float *dx = (float *)malloc(sizeof(float) * (size_t)n), *dy = (float *)malloc(sizeof(float) * (size_t)n);
    int32_t cap = n, *ruv = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cap);
    orc_cluster(dp, n, K, mvn, ncl);
    orc_dpf0(mvn, ncl, p, 0.6, dpf0);
    int32_t nruv = orc_ruv_neighbor(xyuvav, p, p->radius_neighbor_dpf1, ruv, cap);
    orc_dpf1(dpf0, dx, dy, ruv, nruv, mvn, ncl, xyuvav, p);
    nruv = orc_ruv_neighbor(xyuvav, p, p->radius_neighbor_ps, ruv, cap);
    orc_pseudosmooth(dpf0, dx, dy, ruv, nruv, mvn, ncl, xyuvav, p);
    for (int32_t g = 0; g < n; g++)
        for (int k = 0; k < 5; k++)
            planes[(size_t)k * n + g] = dpf0[g] >= 0 ? mvn[((size_t)g * K + dpf0[g]) * 5 + k] : N_A_N;
    free(mvn); free(ncl); free(dpf0); free(dx); free(dy); free(ruv);
}

void orc_finalize(float *planes, const orc_post_params *p, float *du_cp_out, float *dv_cp_out) {
    const int32_t n = p->dimx * p->dimy;
    float sdu = 0.0, sdv = 0.0; int32_t num_cp = 0;
    for (int32_t g = 0; g < n; g++) {
        float a = planes[g], b = planes[(size_t)n + g];
        if (!isnan(a) && !isnan(b)) { sdu += a; sdv += b; num_cp++; }
    }
    float du_cp = sdu / (float)num_cp, dv_cp = sdv / (float)num_cp;
    float f = p->mpp / p->dt * 365;
    for (int32_t g = 0; g < n; g++) {
        planes[g] = (planes[g] - du_cp) * f;
        planes[(size_t)n + g] = -(planes[(size_t)n + g] - dv_cp) * f;
        planes[2 * (size_t)n + g] = sqrt(planes[2 * (size_t)n + g]) * f;
        planes[3 * (size_t)n + g] = sqrt(planes[3 * (size_t)n + g]) * f;
    }
    if (du_cp_out) *du_cp_out = du_cp;
    if (dv_cp_out) *dv_cp_out = dv_cp;
}

/* --------------------------------------------------------------------------------- */
/* get_offset_image :33-492, GMA_double_randperm_row :494-540                          */
/* --------------------------------------------------------------------------------- */
static float img_at(const float *img, int32_t H, int32_t W, int32_t v, int32_t u) {
    /* the reference does no bounds check here (:100-103, :264, :288); outside pixels read as 0 */
    return (u >= 0 && u < W && v >= 0 && v < H) ? img[(size_t)v * W + u] : 0.0f;
}

int orc_get_offset_image(const float *i0, const float *i1, int32_t H, int32_t W, const double *xyuvav, int32_t n,
                         const orc_cp_params *p, unsigned int seed, int32_t *offset, uint8_t *flag_cp) {
    static const float K0[3] = {-1, 0, 1}, K1[3] = {-1, 0, 1};                  /* MIMC_main.c:175-196 */
    static const float K2[9] = {-0.125f, -0.125f, -0.125f, -0.125f, 1.0f, -0.125f, -0.125f, -0.125f, -0.125f};
    const float *kern[3] = {K0, K1, K2};
    const int kh[3] = {1, 3, 3}, kw[3] = {3, 1, 3};
    const int32_t ocw2 = p->vec_ocw[2];
    const int32_t ocw_chip = (int32_t)(ocw2 + p->AW_CRE + 2);                    /* :48 */
    const int32_t T = 2 * ocw_chip + 1, Tw = T + 2;
    int32_t num_cp;
    if (n * p->ratio_cp > p->num_cp_max) num_cp = p->num_cp_max;                 /* :57-64 */
    else num_cp = (int32_t)(n * p->ratio_cp);

    /* candidates :71-121 */
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t num_cand = 0;
    const int32_t thres_numpx = (ocw2 * 2 + 1) * (ocw2 * 2 + 1) / 2;
    for (int32_t g = 0; g < n; g++) {
        float spd_sq = xyuvav[6 * (size_t)g + 4] * xyuvav[6 * (size_t)g + 4] + xyuvav[6 * (size_t)g + 5] * xyuvav[6 * (size_t)g + 5];
        if (!(spd_sq < p->thres_spd_cp * p->thres_spd_cp)) continue;
        int32_t u = (int32_t)xyuvav[6 * (size_t)g + 2], v = (int32_t)xyuvav[6 * (size_t)g + 3];
        int32_t inv0 = 0;
        for (int32_t c1 = -ocw2; c1 <= ocw2; c1++)
            for (int32_t c2 = -ocw2; c2 <= ocw2; c2++) {
                int32_t vv = c1 + v, uu = c2 + u;
                if (uu >= 0 && uu < W && vv >= 0 && vv < H && i0[(size_t)vv * W + uu] < 0.00001) inv0++;
            }
        if (inv0 > thres_numpx) continue;                                        /* :111 (tests i0 twice) */
        cand[num_cand++] = g;
    }
    if (num_cand < p->num_cp_min) { free(cand); return -1; }                     /* :130-135 */
    if (num_cp > num_cand) num_cp = (int32_t)((float)num_cand * 0.75);           /* :137-141 */
    if (num_cp < 1) num_cp = 1;                                                  /* (the reference divides by zero here) */

    /* GMA_double_randperm_row :494-540 on the candidate rows */
    {
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)num_cand), *out = (int32_t *)malloc(sizeof(int32_t) * (size_t)num_cand);
        memcpy(tmp, cand, sizeof(int32_t) * (size_t)num_cand);
        srand(seed);                                                             /* :516 */
        for (int32_t lim = num_cand - 1; lim >= 0; lim--) {
            int32_t idx = lim != 0 ? (int32_t)(rand() % lim) : 0;
            out[lim] = tmp[idx]; tmp[idx] = tmp[0]; tmp[0] = tmp[lim];           /* :531-537 */
        }
        memcpy(cand, out, sizeof(int32_t) * (size_t)num_cand);
        free(tmp); free(out);
    }
    int32_t num_segment = (num_cand < p->num_cp_min) ? 1 : num_cand / num_cp;   /* :186 */
    int32_t *seg = (int32_t *)malloc(sizeof(int32_t) * (size_t)(num_segment + 1));
    seg[0] = 0;
    for (int32_t c = 1; c <= num_segment; c++) seg[c] = (int32_t)(num_cand * ((float)c / (float)num_segment));   /* :193 */

    /* shared rectangular pivot set :165-176 */
    const int32_t R = (int32_t)p->AW_CRE, P = (2 * R + 1) * (2 * R + 1);
    int32_t *piv = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)P);
    { int32_t k = 0; for (int32_t a = -R; a <= R; a++) for (int32_t b = -R; b <= R; b++) { piv[2 * k] = a; piv[2 * k + 1] = b; k++; } }

    float sduv[2] = {0.0f, 0.0f};
    int32_t num_cp_current = 0, ok = 0;
    for (int32_t c = 0; c < num_segment; c++) {
        const int32_t nsub = seg[c + 1] - seg[c];
        if (nsub <= 0) continue;
        float *dp = (float *)calloc((size_t)16 * nsub * 3, sizeof(float));
        float *tile0 = (float *)malloc(sizeof(float) * (size_t)nsub * T * T), *tile1 = (float *)malloc(sizeof(float) * (size_t)nsub * T * T);
        for (int variant = -1; variant <= 2; variant++) {                        /* :248 */
            /* ONE output buffer per image, reused for every node of the segment (:273-276): the
             * stale-border behaviour of GMA_float_conv2 chains the nodes together */
            float *tin = (float *)calloc((size_t)Tw * Tw, sizeof(float));
            float *out0 = (float *)calloc((size_t)Tw * Tw, sizeof(float)), *out1 = (float *)calloc((size_t)Tw * Tw, sizeof(float));
            for (int32_t k = 0; k < nsub; k++) {
                const int32_t g = cand[seg[c] + k];
                const int32_t u = (int32_t)xyuvav[6 * (size_t)g + 2], v = (int32_t)xyuvav[6 * (size_t)g + 3];
                for (int im = 0; im < 2; im++) {
                    const float *img = im ? i1 : i0;
                    float *tile = (im ? tile1 : tile0) + (size_t)k * T * T;
                    if (variant < 0) {                                           /* :250-268 */
                        for (int32_t a = -ocw_chip; a <= ocw_chip; a++)
                            for (int32_t b = -ocw_chip; b <= ocw_chip; b++)
                                tile[(size_t)(a + ocw_chip) * T + (b + ocw_chip)] = img_at(img, H, W, v + a, u + b);
                    } else {                                                     /* :273-308 */
                        float *out = im ? out1 : out0;
                        for (int32_t a = -ocw_chip - 1; a <= ocw_chip + 1; a++)
                            for (int32_t b = -ocw_chip - 1; b <= ocw_chip + 1; b++)
                                tin[(size_t)(a + ocw_chip + 1) * Tw + (b + ocw_chip + 1)] = img_at(img, H, W, v + a, u + b);
                        orc_conv2(tin, Tw, Tw, kern[variant], kh[variant], kw[variant], out);
                        for (int32_t a = 0; a < T; a++)
                            for (int32_t b = 0; b < T; b++) tile[(size_t)a * T + b] = out[(size_t)(a + 1) * Tw + (b + 1)];
                    }
                }
            }
            free(tin); free(out0); free(out1);
            for (int c3 = 1; c3 < 3; c3++) {                                     /* :325 */
                const int32_t ocw = p->vec_ocw[c3], S = 2 * ocw + 1;
                const int32_t slot = (c3 - 1) * 8 + (variant + 1) * 2;
#pragma omp parallel
                {
                    float *chip = (float *)malloc(sizeof(float) * (size_t)S * S);
#pragma omp for schedule(dynamic)
                    for (int32_t k = 0; k < nsub; k++) {
                        for (int dir = 0; dir < 2; dir++) {
                            const float *rt = (dir ? tile1 : tile0) + (size_t)k * T * T, *st = (dir ? tile0 : tile1) + (size_t)k * T * T;
                            for (int32_t a = -ocw; a <= ocw; a++)
                                for (int32_t b = -ocw; b <= ocw; b++)
                                    chip[(size_t)(a + ocw) * S + (b + ocw)] = rt[(size_t)(ocw_chip + a) * T + (ocw_chip + b)];
                            float uvncc[3];
                            orc_find_ncc_peak(chip, S, st, T, T, piv, P, uvncc, NULL, NULL);     /* :351, :369 */
                            float *d = dp + ((size_t)(slot + dir) * nsub + k) * 3;
                            d[0] = dir ? -uvncc[0] : uvncc[0]; d[1] = dir ? -uvncc[1] : uvncc[1]; d[2] = uvncc[2];
                        }
                    }
                    free(chip);
                }
            }
        }
        float *mvn = (float *)calloc((size_t)nsub * 16 * 5, sizeof(float));
        int32_t *ncl = (int32_t *)calloc((size_t)nsub, sizeof(int32_t));
        orc_cluster(dp, nsub, 16, mvn, ncl);                                     /* :389 */
        for (int32_t k = 0; k < nsub; k++)
            for (int32_t q = 0; q < ncl[k]; q++)
                if (mvn[((size_t)k * 16 + q) * 5 + 4] >= 0.6) {                  /* :400 */
                    sduv[0] += mvn[((size_t)k * 16 + q) * 5]; sduv[1] += mvn[((size_t)k * 16 + q) * 5 + 1];
                    flag_cp[cand[seg[c] + k]] = 1;
                    num_cp_current++;
                }
        free(mvn); free(ncl); free(dp); free(tile0); free(tile1);
        if (num_cp <= num_cp_current) { ok = 1; break; }                         /* :419 */
    }
    free(seg); free(piv); free(cand);
    if (!ok && num_cp_current >= p->num_cp_min) ok = 1;                          /* :437-441 */
    if (!ok) return -1;
    float du_cp = sduv[0] / (float)num_cp_current, dv_cp = sduv[1] / (float)num_cp_current;
    offset[0] = du_cp > 0 ? (int32_t)(du_cp + 0.5) : (int32_t)(du_cp - 0.5);    /* :455-473 */
    offset[1] = dv_cp > 0 ? (int32_t)(dv_cp + 0.5) : (int32_t)(dv_cp - 0.5);
    return 1;
}
