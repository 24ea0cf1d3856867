/* CPU model of the CUDA matcher's "explore, then replay" schedule -- TEST INFRASTRUCTURE.
 *
 * The reference's hill climb (find_ncc_peak, MIMC_module.c:691-753) evaluates NCC cells
 * lazily, one 3x3 probe at a time, pivot after pivot.  The CUDA kernel (csrc/match2.cu)
 * cannot afford one evaluation round per probe, so it splits the work:
 *
 *   explore  every pivot walks its climb path on the values known so far, ignoring the
 *            reference's "nothing new in this probe => stop" rule (that rule only ever
 *            SHORTENS a path, so the walked cells are a superset of the reference's); a
 *            pivot whose next 3x3 probe has unknown cells is blocked and asks for them;
 *            all requests of a round are evaluated together (at most `maxj` per round);
 *   replay   once no pivot is blocked, the reference's state machine runs verbatim on
 *            the now-known values and decides visibility, counts, peak and ncc.
 *
 * This file restates that schedule on the CPU so that (i) its equivalence with
 * orc_find_ncc_peak is asserted by the CPU test-suite on thousands of nodes before the
 * kernel is trusted with it, and (ii) the number of rounds and of extra cells can be
 * predicted without a GPU.  Cell arithmetic follows MIMC_module.c:713-735.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mimc3_oracle.h"

#define MIN_DN 0.0000000001

static float ncc_cell(const float *refchip, int S, const float *sarea, int Dx2, int px, int py) {
    const int ocw = S / 2;
    int32_t nsample = 0;
    double sy = 0, sx = 0, sxx = 0, sxy = 0, syy = 0;
    for (int c3 = -ocw; c3 <= ocw; c3++)
        for (int c4 = -ocw; c4 <= ocw; c4++) {
            float r = refchip[(c4 + ocw) * S + (c3 + ocw)];
            float s = sarea[(size_t)(py + c4) * Dx2 + (px + c3)];
            if (r >= MIN_DN && s >= MIN_DN) {
                nsample++;
                sy += s; sx += r;
                sxx += r * r; syy += s * s; sxy += r * s;
            }
        }
    return (float)((nsample * sxy - sx * sy) / sqrt((nsample * sxx - sx * sx) * (nsample * syy - sy * sy)));
}

/* stats[0] rounds, [1] cells computed, [2] explore steps, [3] replay steps */
void orc_model_find_ncc_peak(const float *refchip, int32_t S, const float *sarea, int32_t Dy2, int32_t Dx2,
                             const int32_t *piv, int32_t P, int32_t maxj, float *uvncc, int32_t *peak,
                             int32_t *ncell, int32_t *stats) {
    const float N_A_N = sqrt(-1.0);
    const int dx2 = Dx2 / 2, dy2 = Dy2 / 2, ocw = S / 2;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    uvncc[0] = 0; uvncc[1] = 0; uvncc[2] = -2.0;
    int32_t inv_ref = 0, inv_sa = 0;
    for (int i = 0; i < S * S; i++) if (refchip[i] < MIN_DN) inv_ref++;
    for (int64_t i = 0; i < (int64_t)Dy2 * Dx2; i++) if (sarea[i] < MIN_DN) inv_sa++;
    if ((float)inv_ref / (float)(S * S) > 0.8f || (float)inv_sa / (float)(Dy2 * Dx2) > 0.8f) {
        uvncc[0] = N_A_N; uvncc[1] = N_A_N; uvncc[2] = -3;
        peak[0] = 0; peak[1] = 0; *ncell = 0;
        return;
    }
    const size_t nc = (size_t)Dy2 * Dx2;
    float *val = (float *)calloc(nc, sizeof(float));
    uint8_t *known = (uint8_t *)calloc(nc, 1), *listed = (uint8_t *)calloc(nc, 1), *visible = (uint8_t *)calloc(nc, 1);
    int *wx = (int *)malloc(sizeof(int) * P), *wy = (int *)malloc(sizeof(int) * P), *wst = (int *)malloc(sizeof(int) * P);
    float *wmax = (float *)malloc(sizeof(float) * P);
    int *jobs = (int *)malloc(sizeof(int) * (maxj > 0 ? maxj : 1));
    for (int ip = 0; ip < P; ip++) { wx[ip] = piv[2 * ip] + dx2; wy[ip] = piv[2 * ip + 1] + dy2; wmax[ip] = -2; wst[ip] = 0; }
#define EVALUABLE(x, y) (!((x) - ocw <= 1 || (x) + ocw >= Dx2 - 1 || (y) - ocw <= 1 || (y) + ocw >= Dy2 - 1))
    for (;;) {
        /* ---- explore: every pivot walks as far as the known values allow ---- */
        for (int ip = 0; ip < P; ip++) {
            while (wst[ip] != 2) {
                if (!EVALUABLE(wx[ip], wy[ip])) { wst[ip] = 2; break; }
                int complete = 1;
                for (int c1 = -1; c1 <= 1; c1++)
                    for (int c2 = -1; c2 <= 1; c2++)
                        if (!known[(size_t)(wy[ip] + c2) * Dx2 + (wx[ip] + c1)]) complete = 0;
                if (!complete) { wst[ip] = 1; break; }
                stats[2]++;
                int d0 = 0, d1 = 0;
                for (int c1 = -1; c1 <= 1; c1++)
                    for (int c2 = -1; c2 <= 1; c2++) {
                        float v = val[(size_t)(wy[ip] + c2) * Dx2 + (wx[ip] + c1)];
                        if (v > wmax[ip]) { wmax[ip] = v; d0 = c1; d1 = c2; }
                    }
                if (d0 == 0 && d1 == 0) { wst[ip] = 2; break; }
                wx[ip] += d0; wy[ip] += d1; wst[ip] = 0;
            }
        }
        /* ---- requests of the blocked pivots, at most maxj per round ---- */
        int m = 0;
        for (int ip = 0; ip < P && m < maxj; ip++) {
            if (wst[ip] != 1) continue;
            for (int c1 = -1; c1 <= 1; c1++)
                for (int c2 = -1; c2 <= 1; c2++) {
                    size_t cell = (size_t)(wy[ip] + c2) * Dx2 + (wx[ip] + c1);
                    if (!known[cell] && !listed[cell] && m < maxj) { listed[cell] = 1; jobs[m++] = (int)cell; }
                }
        }
        if (m == 0) break;
        for (int j = 0; j < m; j++) {
            int cx = jobs[j] % Dx2, cy = jobs[j] / Dx2;
            val[jobs[j]] = ncc_cell(refchip, S, sarea, Dx2, cx, cy);
            known[jobs[j]] = 1;
        }
        stats[0]++; stats[1] += m;
    }
    /* ---- replay: the reference's state machine on the known values ---- */
    int uv_peak[2] = {dx2, dy2};
    int32_t ncells = 0, missing = 0;
    for (int ip = 0; ip < P; ip++) {
        int px = piv[2 * ip] + dx2, py = piv[2 * ip + 1] + dy2;
        int duv[2] = {-1, -1};
        float nccmax = -2;
        int flag_new = 1;
        while ((duv[0] != 0 || duv[1] != 0) && flag_new != 0) {
            duv[0] = 0; duv[1] = 0;
            if (!EVALUABLE(px, py)) break;
            stats[3]++;
            flag_new = 0;
            for (int c1 = -1; c1 <= 1; c1++)
                for (int c2 = -1; c2 <= 1; c2++) {
                    size_t cell = (size_t)(py + c2) * Dx2 + (px + c1);
                    if (!known[cell]) missing++;
                    /* `cmap < -1.0` (:713): not yet evaluated, or evaluated to a value below -1 */
                    if (!visible[cell] || val[cell] < -1.0f) { flag_new++; ncells++; visible[cell] = 1; }
                    if (val[cell] > nccmax) { nccmax = val[cell]; duv[0] = c1; duv[1] = c2; }
                }
            px += duv[0]; py += duv[1];
        }
        if (nccmax > uvncc[2]) { uv_peak[0] = px; uv_peak[1] = py; uvncc[2] = nccmax; }
    }
    if (missing) stats[0] = -missing;   /* the schedule left a needed cell unevaluated: a bug */
    float ncc9[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            int x = uv_peak[0] - 1 + c, y = uv_peak[1] - 1 + r;
            float v = -2.0f;
            /* cmap's last row / column are never initialised to -2 by the reference (:677-681): 0 there */
            if (x == Dx2 - 1 || y == Dy2 - 1) v = 0.0f;
            size_t cell = (size_t)y * Dx2 + x;
            if (x >= 0 && x < Dx2 && y >= 0 && y < Dy2 && visible[cell]) v = val[cell];
            ncc9[r * 3 + c] = v;
        }
    double cp[6];
    cp[0] = 6 * ncc9[0] - 12 * ncc9[1] + 6 * ncc9[2] + 6 * ncc9[3] - 12 * ncc9[4] + 6 * ncc9[5] + 6 * ncc9[6] - 12 * ncc9[7] + 6 * ncc9[8];
    cp[1] = 9 * ncc9[0] - 9 * ncc9[2] - 9 * ncc9[6] + 9 * ncc9[8];
    cp[2] = 6 * ncc9[0] + 6 * ncc9[1] + 6 * ncc9[2] - 12 * ncc9[3] - 12 * ncc9[4] - 12 * ncc9[5] + 6 * ncc9[6] + 6 * ncc9[7] + 6 * ncc9[8];
    cp[3] = -6 * ncc9[0] + 6 * ncc9[2] - 6 * ncc9[3] + 6 * ncc9[5] - 6 * ncc9[6] + 6 * ncc9[8];
    cp[4] = -6 * ncc9[0] - 6 * ncc9[1] - 6 * ncc9[2] + 6 * ncc9[6] + 6 * ncc9[7] + 6 * ncc9[8];
    for (int k = 0; k < 5; k++) cp[k] /= 36;
    uvncc[0] = -2 * cp[2] * cp[3] + cp[1] * cp[4];
    uvncc[1] = -2 * cp[0] * cp[4] + cp[1] * cp[3];
    uvncc[0] /= 4 * cp[0] * cp[2] - cp[1] * cp[1];
    uvncc[1] /= 4 * cp[0] * cp[2] - cp[1] * cp[1];
    uvncc[0] += (float)(uv_peak[0] - dx2);
    uvncc[1] += (float)(uv_peak[1] - dy2);
    peak[0] = uv_peak[0] - dx2; peak[1] = uv_peak[1] - dy2;
    *ncell = ncells;
    free(val); free(known); free(listed); free(visible); free(wx); free(wy); free(wst); free(wmax); free(jobs);
}

/* orc_match with the model schedule; stats (n,4). */
void orc_model_match(const float *i0, const float *i1, int32_t H, int32_t W, const double *xyuvav, int32_t n,
                     const int32_t *offset, const int32_t *csr_off, const int32_t *piv, int32_t sign, int32_t ocw,
                     int32_t maxj, float *out, int32_t *peak, int32_t *ncell, int32_t *stats) {
    const int S = 2 * ocw + 1;
#pragma omp parallel for schedule(dynamic)
    for (int32_t g = 0; g < n; g++) {
        float *refchip = (float *)calloc((size_t)S * S, sizeof(float));
        int32_t uv0[2] = {(int32_t)xyuvav[6 * (size_t)g + 2], (int32_t)xyuvav[6 * (size_t)g + 3]};
        for (int c2 = -ocw; c2 <= ocw; c2++)
            for (int c1 = -ocw; c1 <= ocw; c1++)
                refchip[(c2 + ocw) * S + (c1 + ocw)] = i0[(size_t)(uv0[1] + c2) * W + (uv0[0] + c1)];
        uv0[0] += offset[0]; uv0[1] += offset[1];
        int32_t P = csr_off[g + 1] - csr_off[g];
        float uvncc[3] = {sqrtf(-1.0f), sqrtf(-1.0f), -2.0f};
        int32_t pk[2] = {0, 0}, ncl = 0, st[4] = {0, 0, 0, 0};
        if (P > 0) {
            int32_t *pv = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)P);
            for (int32_t k = 0; k < 2 * P; k++) pv[k] = sign * piv[2 * (size_t)csr_off[g] + k];
            int dx2 = abs(pv[2 * (P - 1)]) + ocw + 2, dy2 = abs(pv[2 * (P - 1) + 1]) + ocw + 2;
            int Dx2 = dx2 * 2 + 1, Dy2 = dy2 * 2 + 1;
            float *sarea = (float *)calloc((size_t)Dx2 * Dy2, sizeof(float));
            for (int c2 = -dy2; c2 < dy2; c2++) {
                int cv = uv0[1] + c2;
                if (cv < 0 || cv >= H) continue;
                for (int c1 = -dx2; c1 < dx2; c1++) {
                    int cu = uv0[0] + c1;
                    if (cu >= 0 && cu < W) sarea[(size_t)(c2 + dy2) * Dx2 + (c1 + dx2)] = i1[(size_t)cv * W + cu];
                }
            }
            orc_model_find_ncc_peak(refchip, S, sarea, Dy2, Dx2, pv, P, maxj, uvncc, pk, &ncl, st);
            free(sarea); free(pv);
        }
        out[3 * (size_t)g] = uvncc[0]; out[3 * (size_t)g + 1] = uvncc[1]; out[3 * (size_t)g + 2] = uvncc[2];
        peak[2 * (size_t)g] = pk[0]; peak[2 * (size_t)g + 1] = pk[1];
        ncell[g] = ncl;
        for (int k = 0; k < 4; k++) stats[4 * (size_t)g + k] = st[k];
        free(refchip);
    }
}
