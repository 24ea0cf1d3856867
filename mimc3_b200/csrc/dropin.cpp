// The reference's module interface (MIMC_module.h:34-67) over the mimc3cu C ABI: the drop-in
// boundary of SURVEY.md 8(b).  Host C++ only; all compute happens in libmimc3cu.so.
//
// State kept between calls (the driver is single-threaded and calls strictly in sequence):
//   * device copies of host images, keyed by the host payload pointer (i0/i1 are loaded once;
//     i0c/i1c are only ever written through GMA_float_conv2, i.e. through this file);
//   * the node list, keyed by the xyuvav payload pointer;
//   * one CSR pivot set per get_uv_pivot call, recognised again by the returned pointer; the
//     driver negates the pivots in place between the forward and the swapped pass
//     (MIMC_main.c:272-279), which is detected by comparing the host arrays with the CSR copy.
// Errors: only get_offset_image has a return channel; everything else prints and exit(2)s.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <map>
#include <thread>
#include <vector>

#include "../../include/mimc3_dropin.h"
#include "../../include/mimc3cu.h"

namespace {

mimc3cu_ctx *g_ctx = nullptr;

struct DevImage { int32_t handle; int32_t H, W; };
std::map<const float *, DevImage> g_images;
const double *g_nodes_ptr = nullptr;
int32_t g_nodes_n = 0;

struct PivotCache {
    GMA_int32 **host = nullptr;
    int32_t n = 0, ocw = 0, slot = -1;
    std::vector<int32_t> off, piv;
};
PivotCache g_piv[8];
int g_piv_next = 0;

[[noreturn]] void die(const char *where) {
    fprintf(stderr, "mimc3cu drop-in: %s failed: %s\n", where, mimc3cu_last_error(g_ctx));
    exit(2);
}
#define CK(call, where) do { if (call) die(where); } while (0)

bool env_on(const char *name) { const char *v = getenv(name); return v && *v && strcmp(v, "0"); }

mimc3cu_ctx *ctx() {
    if (!g_ctx) {
        const char *dev = getenv("MIMC3CU_DEVICE");
        if (mimc3cu_create(dev ? atoi(dev) : 0, &g_ctx)) {
            fprintf(stderr, "mimc3cu drop-in: %s\n", mimc3cu_last_error(nullptr));
            exit(2);
        }
        atexit(mimc3_dropin_shutdown);
    }
    return g_ctx;
}

// Device copy of a host image.  `fresh_upload`: copy the host payload now (first sight, or forced).
DevImage &device_image(GMA_float *img, bool upload_if_new) {
    auto it = g_images.find(img->data);
    if (it != g_images.end() && (it->second.H != img->nrows || it->second.W != img->ncols)) {
        mimc3cu_image_destroy(ctx(), it->second.handle);
        g_images.erase(it);
        it = g_images.end();
    }
    bool is_new = it == g_images.end();
    if (is_new) {
        DevImage d{0, img->nrows, img->ncols};
        CK(mimc3cu_image_create(ctx(), img->nrows, img->ncols, &d.handle), "image_create");
        it = g_images.emplace(img->data, d).first;
    }
    if ((is_new && upload_if_new) || env_on("MIMC3CU_DROPIN_REVALIDATE"))
        CK(mimc3cu_image_upload(ctx(), it->second.handle, img->data), "image_upload");
    return it->second;
}

void ensure_nodes(GMA_double *xyuvav) {
    if (xyuvav->ncols < 6) { fprintf(stderr, "mimc3cu drop-in: xyuvav needs 6 columns\n"); exit(2); }
    if (g_nodes_ptr == xyuvav->data && g_nodes_n == xyuvav->nrows && !env_on("MIMC3CU_DROPIN_REVALIDATE")) return;
    std::vector<double> flat;
    const double *src = xyuvav->data;
    if (xyuvav->ncols != 6) {   // e.g. the 7-column CP matrix: repack
        flat.resize((size_t)xyuvav->nrows * 6);
        for (int32_t g = 0; g < xyuvav->nrows; g++) memcpy(&flat[6 * (size_t)g], xyuvav->val[g], 6 * sizeof(double));
        src = flat.data();
    }
    CK(mimc3cu_set_nodes(ctx(), src, xyuvav->nrows), "set_nodes");
    g_nodes_ptr = xyuvav->data; g_nodes_n = xyuvav->nrows;
}

const double *xyuvav_flat(GMA_double *xyuvav, std::vector<double> &tmp) {
    if (xyuvav->ncols == 6) return xyuvav->data;
    tmp.resize((size_t)xyuvav->nrows * 6);
    for (int32_t g = 0; g < xyuvav->nrows; g++) memcpy(&tmp[6 * (size_t)g], xyuvav->val[g], 6 * sizeof(double));
    return tmp.data();
}

void fill_params(mimc3cu_params *p) {
    mimc3cu_default_params(p);
    for (int k = 0; k < 4; k++) p->vec_ocw[k] = param_mimc2.vec_ocw[k];
    p->AW_CRE = param_mimc2.AW_CRE; p->AW_SF = param_mimc2.AW_SF;
    p->mpp = param_mimc2.mpp; p->meter_per_spacing = param_mimc2.meter_per_spacing;
    p->radius_neighbor_dpf1 = param_mimc2.radius_neighbor_dpf1; p->radius_neighbor_ps = param_mimc2.radius_neighbor_ps;
    p->dt = dt; p->dimx = dimx_vmap; p->dimy = dimy_vmap; p->num_dp = num_dp;
    p->num_cp_max = param_mimc2.num_cp_max; p->num_cp_min = param_mimc2.num_cp_min;
    p->ratio_cp = param_mimc2.ratio_cp; p->thres_spd_cp = param_mimc2.thres_spd_cp;
}

template <typename F>
void parallel_for(int32_t n, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw ? (hw > 32 ? 32 : hw) : 4);
    if (n < 8192) nt = 1;
    if (nt == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const int32_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int32_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
        if (b >= e) break;
        th.emplace_back([=] { f(b, e); });
    }
    for (auto &t : th) t.join();
}

// +1: host pivots equal the CSR copy, -1: they are its negation, 0: anything else
int compare_pivots(const PivotCache &pc, GMA_int32 **host) {
    std::atomic<int> verdict(3);   // bit0: may be equal, bit1: may be negated
    parallel_for(pc.n, [&](int32_t b, int32_t e) {
        int ok = 3;
        for (int32_t g = b; g < e && ok; g++) {
            const int32_t P = pc.off[g + 1] - pc.off[g];
            const GMA_int32 *h = host[g];
            if (P <= 0) continue;
            if (h->nrows != P) { ok = 0; break; }
            const int32_t *c = &pc.piv[2 * (size_t)pc.off[g]];
            for (int32_t k = 0; k < P; k++) {
                const int32_t hu = h->val[k][0], hv = h->val[k][1];
                if (hu != c[2 * k] || hv != c[2 * k + 1]) ok &= ~1;
                if (hu != -c[2 * k] || hv != -c[2 * k + 1]) ok &= ~2;
            }
        }
        verdict.fetch_and(ok);
    });
    const int ok = verdict.load();
    if (ok & 1) return 1;
    if (ok & 2) return -1;
    return 0;
}

void flatten_pivots(GMA_int32 **host, int32_t n, std::vector<int32_t> &off, std::vector<int32_t> &piv) {
    off.assign((size_t)n + 1, 0);
    for (int32_t g = 0; g < n; g++) off[g + 1] = off[g] + (host[g]->nrows > 0 ? host[g]->nrows : 0);
    piv.resize(2 * (size_t)off[n] + 2);
    parallel_for(n, [&](int32_t b, int32_t e) {
        for (int32_t g = b; g < e; g++) {
            int32_t *c = &piv[2 * (size_t)off[g]];
            for (int32_t k = 0; k < off[g + 1] - off[g]; k++) { c[2 * k] = host[g]->val[k][0]; c[2 * k + 1] = host[g]->val[k][1]; }
        }
    });
}

PivotCache &new_pivot_slot() {
    PivotCache &pc = g_piv[g_piv_next];
    pc.slot = g_piv_next;
    g_piv_next = (g_piv_next + 1) % 8;
    return pc;
}

}  // namespace

extern "C" {

void mimc3_dropin_shutdown(void) {
    if (g_ctx) { mimc3cu_destroy(g_ctx); g_ctx = nullptr; }
    g_images.clear();
}

int get_offset_image(GMA_float *i0, GMA_float *i1, GMA_float **kern, GMA_double *xyuvav, int32_t *offset, GMA_uint8 *flag_cp) {
    DevImage &a = device_image(i0, true), &b = device_image(i1, true);
    mimc3cu_params p;
    fill_params(&p);
    std::vector<double> tmp;
    const double *xy = xyuvav_flat(xyuvav, tmp);
    const int32_t n = xyuvav->nrows;
    std::vector<uint8_t> flag((size_t)n, 0);
    int32_t result = -1, found = 0;
    const float *k0 = kern ? kern[0]->data : nullptr, *k1 = kern ? kern[1]->data : nullptr, *k2 = kern ? kern[2]->data : nullptr;
    CK(mimc3cu_get_offset_image(ctx(), a.handle, b.handle, xy, n, &p, k0, k1, k2, (uint32_t)time(NULL), offset, flag.data(),
                                &result, &found), "get_offset_image");
    for (int32_t g = 0; g < n; g++) if (flag[g]) flag_cp->val[g][0] = 1;
    if (result == 1) printf("Sufficient # of CP found: %d, offset = [%d, %d]\n", found, offset[0], offset[1]);
    else printf("Not enough # of successful CP measurement (%d<%d)\n", found, param_mimc2.num_cp_min);
    return result;
}

GMA_int32 **get_uv_pivot(GMA_double *xyuvav, float dt_, param prm, int32_t ocw, GMA_float *i1) {
    const int32_t n = xyuvav->nrows;
    std::vector<double> tmp;
    const double *xy = xyuvav_flat(xyuvav, tmp);
    PivotCache &pc = new_pivot_slot();
    pc.off.assign((size_t)n + 1, 0);
    int64_t tot = mimc3cu_get_uv_pivot(xy, n, dt_, prm.mpp, prm.AW_SF, prm.AW_CRE, ocw, i1->nrows, i1->ncols, pc.off.data(), nullptr);
    if (tot < 0) die("get_uv_pivot");
    pc.piv.assign(2 * (size_t)tot + 2, 0);
    if (mimc3cu_get_uv_pivot(xy, n, dt_, prm.mpp, prm.AW_SF, prm.AW_CRE, ocw, i1->nrows, i1->ncols, pc.off.data(), pc.piv.data()) < 0)
        die("get_uv_pivot");
    // ragged host copy in the driver's own allocation layout: it negates the entries in place and
    // frees every list with GMA_int32_destroy (MIMC_main.c:272-279, 295-298)
    GMA_int32 **out = (GMA_int32 **)malloc(sizeof(GMA_int32 *) * (size_t)n);
    for (int32_t g = 0; g < n; g++) {
        const int32_t P = pc.off[g + 1] - pc.off[g];
        // a node without pivots is undefined behaviour in the reference (it writes val[0][0] of a
        // 0-row matrix, MIMC_module.c:589-591); keep one addressable row
        GMA_int32 *m = GMA_int32_create(P > 0 ? P : 1, 2);
        if (P <= 0) { m->val[0][0] = 0; m->val[0][1] = 0; m->nrows = 0; }
        else memcpy(m->data, &pc.piv[2 * (size_t)pc.off[g]], sizeof(int32_t) * 2 * (size_t)P);
        out[g] = m;
    }
    CK(mimc3cu_set_pivots(ctx(), pc.slot, pc.off.data(), pc.piv.data(), n), "set_pivots");
    pc.host = out; pc.n = n; pc.ocw = ocw;
    return out;
}

GMA_float *matching_ncc_dlc_2(GMA_float *i0, GMA_float *i1, GMA_double *xyuvav, int32_t *offset, GMA_int32 **uv_pivot, int32_t ocw,
                              float AW_CRE, float AW_SF) {
    (void)AW_CRE; (void)AW_SF;   // unused in the reference body as well (MIMC_module.c:805-842)
    DevImage &a = device_image(i0, true), &b = device_image(i1, true);
    ensure_nodes(xyuvav);
    const int32_t n = xyuvav->nrows;
    PivotCache *pc = nullptr;
    for (auto &c : g_piv) if (c.host == uv_pivot && c.n == n) pc = &c;
    int sign = 0;
    if (pc) sign = compare_pivots(*pc, uv_pivot);
    if (!pc || sign == 0) {   // pivots the library has not produced (or edited beyond a sign flip): take them as they are
        pc = &new_pivot_slot();
        flatten_pivots(uv_pivot, n, pc->off, pc->piv);
        CK(mimc3cu_set_pivots(ctx(), pc->slot, pc->off.data(), pc->piv.data(), n), "set_pivots");
        pc->host = uv_pivot; pc->n = n; pc->ocw = ocw;
        sign = 1;
    }
    GMA_float *out = GMA_float_create(n, 3);
    CK(mimc3cu_match(ctx(), a.handle, b.handle, offset, pc->slot, sign, ocw, 0, out->data, nullptr, nullptr), "matching_ncc_dlc_2");
    return out;
}

void GMA_float_conv2(GMA_float *in, GMA_float *kern, GMA_float *out) {
    DevImage &a = device_image(in, true);
    // `out` keeps whatever it holds outside the interior (stale-border semantics):
    // on first sight its host content is taken over, afterwards the device copy is authoritative
    DevImage &o = device_image(out, true);
    CK(mimc3cu_conv2(ctx(), a.handle, kern->data, kern->nrows, kern->ncols, o.handle), "GMA_float_conv2");
    if (!env_on("MIMC3CU_DROPIN_NO_WRITEBACK")) CK(mimc3cu_image_download(ctx(), o.handle, out->data), "image_download");
}

GMA_float **mimc2_postprocess(GMA_float **dp, GMA_double *xyuvav, float dt_) {
    mimc3cu_params p;
    fill_params(&p);
    p.dt = dt_;
    const int32_t n = xyuvav->nrows, K = num_dp;
    std::vector<double> tmp;
    const double *xy = xyuvav_flat(xyuvav, tmp);
    void *d_dp = nullptr, *d_planes = nullptr;
    CK(mimc3cu_malloc(ctx(), sizeof(float) * 3 * (size_t)n * K, &d_dp), "malloc");
    CK(mimc3cu_malloc(ctx(), sizeof(float) * 5 * (size_t)n, &d_planes), "malloc");
    for (int32_t a = 0; a < K; a++)
        CK(mimc3cu_memcpy_h2d(ctx(), (float *)d_dp + (size_t)a * n * 3, dp[a]->data, sizeof(float) * 3 * (size_t)n), "memcpy_h2d");
    CK(mimc3cu_postprocess(ctx(), (const float *)d_dp, xy, &p, (float *)d_planes, nullptr), "mimc2_postprocess");
    GMA_float **out = (GMA_float **)malloc(sizeof(GMA_float *) * 5);
    for (int k = 0; k < 5; k++) {
        out[k] = GMA_float_create(dimy_vmap, dimx_vmap);
        CK(mimc3cu_memcpy_d2h(ctx(), out[k]->data, (float *)d_planes + (size_t)k * n, sizeof(float) * (size_t)n), "memcpy_d2h");
    }
    mimc3cu_free(ctx(), d_dp); mimc3cu_free(ctx(), d_planes);
    return out;
}

}  // extern "C"
