// The reference's module interface (MIMC_module.h:34-67) over the mimc3cu C ABI: the drop-in
// boundary of SURVEY.md 8(b).  Host C++ only; all compute happens in libmimc3cu.so.
//
// What the driver does around these calls (MIMC_main.c:229-402) shapes the state kept here:
//   * i0/i1 are loaded once; i0c/i1c are only ever written through GMA_float_conv2.  Device copies
//     are keyed by the host payload pointer; get_offset_image (the first call of a pair) refreshes
//     i0/i1 and drops every other cached object, so a driver that reuses buffers for the next pair
//     never sees stale device data (mimc3_dropin_invalidate does the same for one pointer).
//   * get_uv_pivot is called 16 times for 4 distinct pivot sets: the CSR copy, its device upload and
//     the matcher's shared-memory bins are kept per chip half-width and only the ragged host copy
//     (which the driver negates in place and frees node by node, :272-298) is rebuilt.  The negation
//     is recognised by probing a few nodes, not by walking every pivot.
//   * every matching_ncc_dlc_2 result stays on the device as well (attempt-major, in call order), so
//     mimc2_postprocess skips the 32 host-to-device copies when the driver's arrays still hold what was
//     returned to it -- apart from the sign flip it applies to the swapped passes (:289-293), which is
//     detected on a sample and redone on the device.
//   * GMA_float_conv2 does not copy the filtered image back (the driver never reads i0c/i1c;
//     MIMC3CU_DROPIN_WRITEBACK=1 restores the copy).
//   * MIMC3CU_DEVICES=n shards the node rows over n GPUs of the box: one context per GPU, both images
//     on every GPU, a contiguous band of node rows each (balanced by the pivot counts), one host thread
//     per GPU inside every entry point, and the banded postprocess over the library's NCCL
//     communicator (mimc3cu_comm_init_all).
// Errors: only get_offset_image has a return channel; everything else prints and exit(2)s.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <chrono>
#include <map>
#include <thread>
#include <vector>

#include "../../include/mimc3_dropin.h"
#include "../../include/mimc3cu.h"

namespace {

struct DevImage { int32_t handle; int32_t H, W; };

struct Dev {                                  // one GPU: its context, its band of node rows, its caches
    mimc3cu_ctx *ctx = nullptr;
    std::map<const float *, DevImage> images;
    int32_t row0 = 0, rows = 0;               // band of node rows [row0, row0 + rows)
    int32_t g0 = 0, n = 0;                    // first node and node count of the band
    float *dp = nullptr;                      // (num_dp, n, 3) results of the attempts, in call order
    size_t dp_cap = 0;
    float *planes = nullptr;                  // (5, n)
    size_t planes_cap = 0;
    bool nodes_set = false;
};
std::vector<Dev> g_devs;
bool g_multi_comm = false;

// one pivot set per chip half-width
struct PivotSet {
    int32_t ocw = 0, n = 0, H = 0, W = 0, slot = -1;
    float dt = 0, mpp = 0, sf = 0, cre = 0;
    const double *xy = nullptr;
    std::vector<int32_t> off, piv;            // CSR over all nodes
    bool uploaded = false;
    GMA_int32 **host = nullptr;               // the ragged copy currently in the driver's hands
};
PivotSet g_piv[8];
int g_piv_count = 0;

// the attempts returned so far: host array -> device slot, and a fingerprint to recognise it again
constexpr int kProbe = 64;
struct Attempt { const GMA_float *host = nullptr; float probe[kProbe][3]; };
std::vector<Attempt> g_attempts;
const double *g_xy = nullptr;
int32_t g_n = 0, g_dimx = 0, g_dimy = 0;
bool g_timing = false;

[[noreturn]] void die(const char *where, mimc3cu_ctx *c = nullptr) {
    fprintf(stderr, "mimc3cu drop-in: %s failed: %s\n", where, mimc3cu_last_error(c));
    exit(2);
}
#define CK(call, where, ctxp) do { if (call) die(where, ctxp); } while (0)

bool env_on(const char *name) { const char *v = getenv(name); return v && *v && strcmp(v, "0"); }
int env_int(const char *name, int dflt) { const char *v = getenv(name); return v && *v ? atoi(v) : dflt; }

struct Stopwatch {
    const char *what; std::chrono::steady_clock::time_point t0;
    explicit Stopwatch(const char *w) : what(w), t0(std::chrono::steady_clock::now()) {}
    ~Stopwatch() {
        if (g_timing) fprintf(stderr, "[mimc3cu drop-in] %-22s %9.3f ms\n", what,
                              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

void init() {
    if (!g_devs.empty()) return;
    g_timing = env_on("MIMC3CU_DROPIN_TIMING");
    int want = env_int("MIMC3CU_DEVICES", 1), first = env_int("MIMC3CU_DEVICE", 0);
    const int have = mimc3cu_device_count();
    if (want < 1) want = 1;
    if (first + want > have && have > 0) {
        fprintf(stderr, "mimc3cu drop-in: MIMC3CU_DEVICE=%d + MIMC3CU_DEVICES=%d exceeds the %d visible GPUs\n", first, want, have);
        exit(2);
    }
    g_devs.resize(want);
    for (int d = 0; d < want; d++)
        if (mimc3cu_create(first + d, &g_devs[d].ctx)) {
            fprintf(stderr, "mimc3cu drop-in: %s\n", mimc3cu_last_error(nullptr));
            exit(2);
        }
    if (want > 1) {
        std::vector<mimc3cu_ctx *> cs;
        for (auto &d : g_devs) cs.push_back(d.ctx);
        CK(mimc3cu_comm_init_all(cs.data(), want), "comm_init_all", cs[0]);
        g_multi_comm = true;
    }
    atexit(mimc3_dropin_shutdown);
}

// f(d) for every GPU, one host thread each
template <typename F>
void for_devs(F f) {
    init();
    if (g_devs.size() == 1) { f(0); return; }
    std::vector<std::thread> th;
    for (size_t d = 0; d < g_devs.size(); d++) th.emplace_back([=] { f((int)d); });
    for (auto &t : th) t.join();
}

template <typename F>
void parallel_for(int32_t n, F f) {
    static const int cap = [] { int v = env_int("MIMC3CU_HOST_THREADS", 0); return v > 0 ? v : 32; }();
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw ? hw : 4);
    if (nt > cap) nt = cap;
    if (n < 8192) nt = 1;
    if (nt <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const int32_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int32_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
        if (b >= e) break;
        th.emplace_back([=] { f(b, e); });
    }
    for (auto &t : th) t.join();
}

// Device copy of a host image on GPU d.  mode 0: take the host content on first sight; 1: always upload;
// 2: conv2 output -- zero-initialised on first sight, the device copy is authoritative afterwards.
DevImage &device_image(int d, GMA_float *img, int mode) {
    Dev &D = g_devs[d];
    auto it = D.images.find(img->data);
    if (it != D.images.end() && (it->second.H != img->nrows || it->second.W != img->ncols)) {
        mimc3cu_image_destroy(D.ctx, it->second.handle);
        D.images.erase(it);
        it = D.images.end();
    }
    const bool is_new = it == D.images.end();
    if (is_new) {
        DevImage di{0, img->nrows, img->ncols};
        CK(mimc3cu_image_create(D.ctx, img->nrows, img->ncols, &di.handle), "image_create", D.ctx);   // zero-filled
        it = D.images.emplace(img->data, di).first;
    }
    const bool revalidate = env_on("MIMC3CU_DROPIN_REVALIDATE");
    if (mode == 1 || revalidate || (is_new && mode == 0))
        CK(mimc3cu_image_upload(D.ctx, it->second.handle, img->data), "image_upload", D.ctx);
    return it->second;
}

const double *xyuvav_flat(GMA_double *xyuvav, std::vector<double> &tmp) {
    if (xyuvav->ncols == 6) return xyuvav->data;
    if (xyuvav->ncols < 6) { fprintf(stderr, "mimc3cu drop-in: xyuvav needs 6 columns\n"); exit(2); }
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

void drop_run_state() {
    for (auto &ps : g_piv) ps = PivotSet();
    g_piv_count = 0;
    g_attempts.clear();
    g_xy = nullptr; g_n = 0;
    for (auto &D : g_devs) D.nodes_set = false;
}

// Bands of node rows, one per GPU, balanced by the matching cost of the rows (sum of P over the row's nodes;
// calibrated on the 8-GPU run of one scene, see bands.row_work);
// every band keeps at least `halo` rows for the banded postprocess.
void assign_bands(const double *xy, int32_t n, const std::vector<int32_t> *off) {
    const int nd = (int)g_devs.size();
    const bool grid = dimx_vmap > 0 && dimy_vmap > 0 && (int64_t)dimx_vmap * dimy_vmap == n;
    g_xy = xy; g_n = n; g_dimx = grid ? dimx_vmap : n; g_dimy = grid ? dimy_vmap : 1;
    mimc3cu_params p;
    fill_params(&p);
    const int halo = mimc3cu_band_halo(&p);
    if (nd > 1 && (!grid || g_dimy < nd * halo)) {
        fprintf(stderr, "mimc3cu drop-in: %d nodes do not form a grid of >= %d rows for %d GPUs\n", n, nd * halo, nd);
        exit(2);
    }
    std::vector<double> cum((size_t)g_dimy + 1, 0.0);
    for (int32_t r = 0; r < g_dimy; r++) {
        double w = g_dimx;
        if (off) w = (double)((*off)[(size_t)(r + 1) * g_dimx] - (*off)[(size_t)r * g_dimx]);
        cum[r + 1] = cum[r] + w;
    }
    int32_t prev = 0;
    for (int d = 0; d < nd; d++) {
        int32_t cut = g_dimy;
        if (d + 1 < nd) {
            const double target = cum[g_dimy] * (d + 1) / nd;
            cut = prev;
            while (cut < g_dimy && cum[cut] < target) cut++;
            if (cut < prev + halo) cut = prev + halo;
            if (cut > g_dimy - (nd - 1 - d) * halo) cut = g_dimy - (nd - 1 - d) * halo;
        }
        Dev &D = g_devs[d];
        D.row0 = prev; D.rows = cut - prev; D.g0 = prev * g_dimx; D.n = D.rows * g_dimx; D.nodes_set = false;
        prev = cut;
    }
}

void ensure_nodes(int d, const double *xy) {
    Dev &D = g_devs[d];
    if (D.nodes_set) return;
    CK(mimc3cu_set_nodes(D.ctx, xy + 6 * (size_t)D.g0, D.n), "set_nodes", D.ctx);
    D.nodes_set = true;
}

void upload_pivots(PivotSet &ps) {
    for_devs([&](int d) {
        Dev &D = g_devs[d];
        // the band's CSR: offsets rebased to the band's first pivot
        std::vector<int32_t> off((size_t)D.n + 1);
        const int32_t base = ps.off[D.g0];
        for (int32_t g = 0; g <= D.n; g++) off[g] = ps.off[(size_t)D.g0 + g] - base;
        CK(mimc3cu_set_pivots(D.ctx, ps.slot, off.data(), ps.piv.data() + 2 * (size_t)base, D.n), "set_pivots", D.ctx);
    });
    ps.uploaded = true;
}

// +1: the driver's lists equal the CSR copy, -1: they are its negation, 0: neither.  Probes kProbe nodes
// (a list of one pivot is (0,0) either way and says nothing); `full` walks everything.
int pivot_sign(const PivotSet &ps, GMA_int32 **host, bool full) {
    const int32_t n = ps.n;
    std::atomic<int> verdict(3);   // bit0: may be equal, bit1: may be negated
    auto check = [&](int32_t g, int &ok) {
        const int32_t P = ps.off[g + 1] - ps.off[g];
        if (P <= 0) return;
        const GMA_int32 *h = host[g];
        if (h->nrows != P) { ok = 0; return; }
        const int32_t *c = &ps.piv[2 * (size_t)ps.off[g]];
        for (int32_t k = 0; k < P; k++) {
            const int32_t hu = h->val[k][0], hv = h->val[k][1];
            if (hu != c[2 * k] || hv != c[2 * k + 1]) ok &= ~1;
            if (hu != -c[2 * k] || hv != -c[2 * k + 1]) ok &= ~2;
        }
    };
    if (full) {
        parallel_for(n, [&](int32_t b, int32_t e) {
            int ok = 3;
            for (int32_t g = b; g < e && ok; g++) check(g, ok);
            verdict.fetch_and(ok);
        });
    } else {
        int ok = 3;
        for (int s = 0; s < kProbe && ok; s++) check((int32_t)(((int64_t)s * (n - 1)) / (kProbe - 1 > 0 ? kProbe - 1 : 1)), ok);
        verdict = ok;
    }
    const int ok = verdict.load();
    if (ok == 3) return full ? 1 : pivot_sign(ps, host, true);   // only one-pivot lists probed: look at all of them
    if (ok & 1) return 1;
    if (ok & 2) return -1;
    return 0;
}

int32_t probe_node(int s, int32_t n) { return (int32_t)(((int64_t)s * (n - 1)) / (kProbe - 1)); }

bool same_float(float a, float b) { return (a != a && b != b) || memcmp(&a, &b, 4) == 0; }

}  // namespace

extern "C" {

void mimc3_dropin_shutdown(void) {
    for (auto &D : g_devs) {
        if (!D.ctx) continue;
        if (D.dp) mimc3cu_free(D.ctx, D.dp);
        if (D.planes) mimc3cu_free(D.ctx, D.planes);
    }
    for (auto &D : g_devs) if (D.ctx) mimc3cu_comm_destroy(D.ctx);
    for (auto &D : g_devs) if (D.ctx) { mimc3cu_destroy(D.ctx); D.ctx = nullptr; }
    g_devs.clear();
}

void mimc3_dropin_invalidate(const void *host_payload) {
    for (auto &D : g_devs) {
        auto it = D.images.find((const float *)host_payload);
        if (it != D.images.end()) { mimc3cu_image_destroy(D.ctx, it->second.handle); D.images.erase(it); }
    }
    if (host_payload == (const void *)g_xy) drop_run_state();
    for (auto &a : g_attempts) if (a.host && (const void *)a.host->data == host_payload) a.host = nullptr;
}

int get_offset_image(GMA_float *i0, GMA_float *i1, GMA_float **kern, GMA_double *xyuvav, int32_t *offset, GMA_uint8 *flag_cp) {
    init();
    Stopwatch sw("get_offset_image");
    // first call of an image pair: whatever was cached belongs to the previous pair
    drop_run_state();
    for (auto &D : g_devs) {
        for (auto &kv : D.images) mimc3cu_image_destroy(D.ctx, kv.second.handle);
        D.images.clear();
    }
    for_devs([&](int d) { device_image(d, i0, 1); device_image(d, i1, 1); });
    Dev &D = g_devs[0];   // <= 500 control points: one GPU
    DevImage &a = device_image(0, i0, 0), &b = device_image(0, i1, 0);
    mimc3cu_params p;
    fill_params(&p);
    std::vector<double> tmp;
    const double *xy = xyuvav_flat(xyuvav, tmp);
    const int32_t n = xyuvav->nrows;
    std::vector<uint8_t> flag((size_t)n, 0);
    int32_t result = -1, found = 0;
    const float *k0 = kern ? kern[0]->data : nullptr, *k1 = kern ? kern[1]->data : nullptr, *k2 = kern ? kern[2]->data : nullptr;
    CK(mimc3cu_get_offset_image(D.ctx, a.handle, b.handle, xy, n, &p, k0, k1, k2, (uint32_t)time(NULL), offset, flag.data(),
                                &result, &found), "get_offset_image", D.ctx);
    for (int32_t g = 0; g < n; g++) if (flag[g]) flag_cp->val[g][0] = 1;
    if (result == 1) printf("Sufficient # of CP found: %d, offset = [%d, %d]\n", found, offset[0], offset[1]);
    else printf("Not enough # of successful CP measurement (%d<%d)\n", found, param_mimc2.num_cp_min);
    return result;
}

GMA_int32 **get_uv_pivot(GMA_double *xyuvav, float dt_, param prm, int32_t ocw, GMA_float *i1) {
    init();
    Stopwatch sw("get_uv_pivot");
    const int32_t n = xyuvav->nrows;
    std::vector<double> tmp;
    const double *xy = xyuvav_flat(xyuvav, tmp);
    if (xyuvav->ncols == 6 && (xy != g_xy || n != g_n)) drop_run_state();   // another node list: nothing cached applies
    PivotSet *ps = nullptr;
    for (int k = 0; k < g_piv_count; k++) {
        PivotSet &c = g_piv[k];
        if (c.ocw == ocw && c.n == n && c.xy == xy && c.H == i1->nrows && c.W == i1->ncols && c.dt == dt_ && c.mpp == prm.mpp &&
            c.sf == prm.AW_SF && c.cre == prm.AW_CRE && xyuvav->ncols == 6 && !env_on("MIMC3CU_DROPIN_REVALIDATE")) ps = &c;
    }
    if (!ps) {
        if (g_piv_count == 8) { for (auto &c : g_piv) c = PivotSet(); g_piv_count = 0; }
        ps = &g_piv[g_piv_count];
        *ps = PivotSet();
        ps->slot = g_piv_count++;
        ps->ocw = ocw; ps->n = n; ps->xy = xy; ps->H = i1->nrows; ps->W = i1->ncols;
        ps->dt = dt_; ps->mpp = prm.mpp; ps->sf = prm.AW_SF; ps->cre = prm.AW_CRE;
        ps->off.assign((size_t)n + 1, 0);
        int64_t tot = mimc3cu_get_uv_pivot(xy, n, dt_, prm.mpp, prm.AW_SF, prm.AW_CRE, ocw, i1->nrows, i1->ncols, ps->off.data(), nullptr);
        if (tot < 0) die("get_uv_pivot");
        ps->piv.assign(2 * (size_t)tot + 2, 0);
        if (mimc3cu_get_uv_pivot(xy, n, dt_, prm.mpp, prm.AW_SF, prm.AW_CRE, ocw, i1->nrows, i1->ncols, ps->off.data(), ps->piv.data()) < 0)
            die("get_uv_pivot");
        if (xyuvav->ncols != 6) ps->xy = nullptr;   // repacked copy: do not recognise it again
    }
    // ragged host copy in the driver's own allocation layout: it negates the entries in place and frees every
    // list with GMA_int32_destroy (MIMC_main.c:272-279, 295-298)
    GMA_int32 **out = (GMA_int32 **)malloc(sizeof(GMA_int32 *) * (size_t)n);
    const std::vector<int32_t> &off = ps->off, &piv = ps->piv;
    parallel_for(n, [&](int32_t b, int32_t e) {
        for (int32_t g = b; g < e; g++) {
            const int32_t P = off[g + 1] - off[g];
            // a node without pivots is undefined behaviour in the reference (it writes val[0][0] of a
            // 0-row matrix, MIMC_module.c:589-591); keep one addressable row
            GMA_int32 *m = GMA_int32_create(P > 0 ? P : 1, 2);
            if (P <= 0) { m->val[0][0] = 0; m->val[0][1] = 0; m->nrows = 0; }
            else memcpy(m->data, &piv[2 * (size_t)off[g]], sizeof(int32_t) * 2 * (size_t)P);
            out[g] = m;
        }
    });
    ps->host = out;
    return out;
}

GMA_float *matching_ncc_dlc_2(GMA_float *i0, GMA_float *i1, GMA_double *xyuvav, int32_t *offset, GMA_int32 **uv_pivot, int32_t ocw,
                              float AW_CRE, float AW_SF) {
    (void)AW_CRE; (void)AW_SF;   // unused in the reference body as well (MIMC_module.c:805-842)
    init();
    Stopwatch sw("matching_ncc_dlc_2");
    const int32_t n = xyuvav->nrows;
    std::vector<double> tmp;
    const double *xy = xyuvav_flat(xyuvav, tmp);
    PivotSet *ps = nullptr;
    for (int k = 0; k < g_piv_count; k++) if (g_piv[k].host == uv_pivot && g_piv[k].n == n && g_piv[k].ocw == ocw) ps = &g_piv[k];
    int sign = ps ? pivot_sign(*ps, uv_pivot, false) : 0;
    if (!ps || sign == 0) {
        // pivots the library has not produced (or edited beyond a sign flip): take them as they are
        if (g_piv_count == 8) { for (auto &c : g_piv) c = PivotSet(); g_piv_count = 0; }
        ps = &g_piv[g_piv_count];
        *ps = PivotSet();
        ps->slot = g_piv_count++;
        ps->ocw = -ocw; ps->n = n; ps->host = uv_pivot;    // negative ocw: never matched by get_uv_pivot's cache
        ps->off.assign((size_t)n + 1, 0);
        for (int32_t g = 0; g < n; g++) ps->off[g + 1] = ps->off[g] + (uv_pivot[g]->nrows > 0 ? uv_pivot[g]->nrows : 0);
        ps->piv.resize(2 * (size_t)ps->off[n] + 2);
        parallel_for(n, [&](int32_t b, int32_t e) {
            for (int32_t g = b; g < e; g++) {
                int32_t *c = &ps->piv[2 * (size_t)ps->off[g]];
                for (int32_t k = 0; k < ps->off[g + 1] - ps->off[g]; k++) { c[2 * k] = uv_pivot[g]->val[k][0]; c[2 * k + 1] = uv_pivot[g]->val[k][1]; }
            }
        });
        sign = 1;
    }
    if (xy != g_xy || n != g_n || g_devs[0].n == 0) assign_bands(xy, n, &ps->off);
    if (!ps->uploaded) upload_pivots(*ps);

    const int32_t K = num_dp > 0 ? num_dp : 32;
    if ((int32_t)g_attempts.size() >= K) g_attempts.clear();   // a new run of attempts over the same pair
    const int32_t slot = (int32_t)g_attempts.size();
    // the kernels are enqueued first; the host array the driver gets back is allocated (and its pages touched) while
    // they run -- at 16.6 M nodes the allocation and the page faults of a 200 MB array cost ~0.1 s per attempt
    for_devs([&](int d) {
        Dev &D = g_devs[d];
        DevImage &a = device_image(d, i0, 0), &b = device_image(d, i1, 0);
        ensure_nodes(d, xy);
        const size_t need = (size_t)K * D.n * 3;
        if (need > D.dp_cap) {
            // first attempt of a run (or a bigger grid): nothing worth keeping
            if (D.dp) CK(mimc3cu_free(D.ctx, D.dp), "free", D.ctx);
            CK(mimc3cu_malloc(D.ctx, need * sizeof(float), (void **)&D.dp), "malloc", D.ctx);
            D.dp_cap = need;
        }
        float *dst = D.dp + (size_t)slot * D.n * 3;
        CK(mimc3cu_match_async(D.ctx, a.handle, b.handle, offset, ps->slot, sign, ocw, 0, dst, nullptr, nullptr), "matching_ncc_dlc_2", D.ctx);
    });
    GMA_float *out = GMA_float_create(n, 3);
    parallel_for(n, [&](int32_t b, int32_t e) { memset(out->data + 3 * (size_t)b, 0, sizeof(float) * 3 * (size_t)(e - b)); });
    for_devs([&](int d) {
        Dev &D = g_devs[d];
        const float *src = D.dp + (size_t)slot * D.n * 3;
        CK(mimc3cu_memcpy_d2h(D.ctx, out->data + 3 * (size_t)D.g0, src, sizeof(float) * 3 * (size_t)D.n), "memcpy_d2h", D.ctx);
    });
    Attempt at;
    at.host = out;
    for (int s = 0; s < kProbe; s++) memcpy(at.probe[s], out->data + 3 * (size_t)probe_node(s, n), 3 * sizeof(float));
    g_attempts.push_back(at);
    return out;
}

void GMA_float_conv2(GMA_float *in, GMA_float *kern, GMA_float *out) {
    init();
    Stopwatch sw("GMA_float_conv2");
    const bool writeback = env_on("MIMC3CU_DROPIN_WRITEBACK");
    for_devs([&](int d) {
        Dev &D = g_devs[d];
        DevImage &a = device_image(d, in, 0);
        // `out` keeps whatever it holds outside the interior (stale-border semantics): zero-initialised
        // on first sight like the driver's allocation (H1), afterwards the device copy is authoritative
        DevImage &o = device_image(d, out, writeback ? 0 : 2);
        CK(mimc3cu_conv2(D.ctx, a.handle, kern->data, kern->nrows, kern->ncols, o.handle), "GMA_float_conv2", D.ctx);
        if (writeback && d == 0) CK(mimc3cu_image_download(D.ctx, o.handle, out->data), "image_download", D.ctx);
    });
}

GMA_float **mimc2_postprocess(GMA_float **dp, GMA_double *xyuvav, float dt_) {
    init();
    Stopwatch sw("mimc2_postprocess");
    mimc3cu_params p;
    fill_params(&p);
    p.dt = dt_;
    const int32_t n = xyuvav->nrows, K = num_dp;
    if ((int64_t)dimx_vmap * dimy_vmap != n) {
        // the reference indexes its grids with dimx*dimy as well; a node count that is not a full grid would make the
        // attempts' stride and the planes' size disagree
        fprintf(stderr, "mimc3cu drop-in: mimc2_postprocess needs nrows (%d) == dimx_vmap * dimy_vmap (%d x %d)\n", n, dimx_vmap, dimy_vmap);
        exit(2);
    }
    std::vector<double> tmp;
    const double *xy = xyuvav_flat(xyuvav, tmp);
    if (xy != g_xy || n != g_n || g_devs[0].n == 0) { assign_bands(xy, n, nullptr); g_attempts.clear(); }

    // which of the driver's arrays are still what was returned to it (apart from the sign flip of du, dv)?
    std::vector<int> action((size_t)K, 2);   // 0: device copy is current, 1: negate du, dv on the device, 2: upload
    for (int32_t a = 0; a < K; a++) {
        if (a >= (int32_t)g_attempts.size() || g_attempts[a].host != dp[a] || dp[a]->nrows != n) continue;
        bool same = true, neg = true;
        for (int s = 0; s < kProbe; s++) {
            const float *h = dp[a]->data + 3 * (size_t)probe_node(s, n), *q = g_attempts[a].probe[s];
            same = same && same_float(h[0], q[0]) && same_float(h[1], q[1]) && same_float(h[2], q[2]);
            neg = neg && same_float(h[0], -q[0]) && same_float(h[1], -q[1]) && same_float(h[2], q[2]);
        }
        action[a] = same ? 0 : (neg ? 1 : 2);
    }
    GMA_float **out = (GMA_float **)malloc(sizeof(GMA_float *) * 5);
    for (int k = 0; k < 5; k++) out[k] = GMA_float_create(dimy_vmap, dimx_vmap);
    for_devs([&](int d) {
        Dev &D = g_devs[d];
        const size_t need = (size_t)K * D.n * 3;
        if (need > D.dp_cap) {
            if (D.dp) CK(mimc3cu_free(D.ctx, D.dp), "free", D.ctx);
            CK(mimc3cu_malloc(D.ctx, need * sizeof(float), (void **)&D.dp), "malloc", D.ctx);
            D.dp_cap = need;
            for (auto &x : action) x = 2;
        }
        if ((size_t)5 * D.n > D.planes_cap) {
            if (D.planes) CK(mimc3cu_free(D.ctx, D.planes), "free", D.ctx);
            CK(mimc3cu_malloc(D.ctx, sizeof(float) * 5 * (size_t)D.n, (void **)&D.planes), "malloc", D.ctx);
            D.planes_cap = (size_t)5 * D.n;
        }
        for (int32_t a = 0; a < K; a++) {
            float *slot = D.dp + (size_t)a * D.n * 3;
            if (action[a] == 2) CK(mimc3cu_memcpy_h2d(D.ctx, slot, dp[a]->data + 3 * (size_t)D.g0, sizeof(float) * 3 * (size_t)D.n), "memcpy_h2d", D.ctx);
            else if (action[a] == 1) CK(mimc3cu_dp_negate_uv_async(D.ctx, slot, D.n), "dp_negate", D.ctx);
        }
        if (g_devs.size() == 1) CK(mimc3cu_postprocess(D.ctx, D.dp, xy, &p, D.planes, nullptr), "mimc2_postprocess", D.ctx);
        else CK(mimc3cu_postprocess_band(D.ctx, D.dp, xy, &p, D.row0, D.rows, nullptr, D.planes, nullptr), "mimc2_postprocess", D.ctx);
        for (int k = 0; k < 5; k++)
            CK(mimc3cu_memcpy_d2h(D.ctx, out[k]->data + (size_t)D.g0, D.planes + (size_t)k * D.n, sizeof(float) * (size_t)D.n), "memcpy_d2h", D.ctx);
    });
    // the device copies now hold the driver's signs: keep the fingerprints in step
    for (int32_t a = 0; a < K && a < (int32_t)g_attempts.size(); a++)
        if (action[a] == 1) for (int s = 0; s < kProbe; s++) { g_attempts[a].probe[s][0] = -g_attempts[a].probe[s][0]; g_attempts[a].probe[s][1] = -g_attempts[a].probe[s][1]; }
    return out;
}

}  // extern "C"
