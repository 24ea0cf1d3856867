// Band communicator: the NCCL side of the multi-GPU postprocess.
//
// Node rows are sharded over GPUs in contiguous bands, band r on communicator rank r.  Matching needs no
// collective.  The iterative stages of mimc2_postprocess (get_dpf1, MIMC_module.c:1387-1612; the
// pseudosmoothing, :2077-2288) read neighbours up to `halo` node rows away, so after every committed sweep
// a band refreshes the halo rows it keeps of its two neighbours (grouped ncclSend/ncclRecv straight out of /
// into the field arrays, no packing), ORs the dirty flags it scattered into its neighbours' rows into their
// owners, and all bands sum their int32 sweep counters (ncclAllReduce on the device counters).  Everything is
// enqueued on the context's stream: the only host synchronisation per sweep is the read-back of the reduced
// counters that the reference's loop control needs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) instead of at link time: inside a PyTorch process the
// library must use the NCCL PyTorch has already loaded (two NCCL copies in one process do not mix), and the
// single-GPU product must load on a host without NCCL.  Either one process per GPU (mimc3cu_comm_init_rank
// with an id distributed by the caller, e.g. over torch.distributed) or one process driving all GPUs
// (mimc3cu_comm_init_all, what the drop-in CLI does).
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("MIMC3CU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
#define BIND(field, sym)                                                             \
    api.field = (decltype(api.field))dlsym(api.handle, sym);                         \
    if (!api.field) { api.error = std::string("libnccl lacks ") + sym; return; }
        BIND(GetUniqueId, "ncclGetUniqueId") BIND(CommInitRank, "ncclCommInitRank") BIND(CommInitAll, "ncclCommInitAll")
        BIND(CommDestroy, "ncclCommDestroy") BIND(GroupStart, "ncclGroupStart") BIND(GroupEnd, "ncclGroupEnd")
        BIND(Send, "ncclSend") BIND(Recv, "ncclRecv") BIND(AllReduce, "ncclAllReduce") BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    });
    return &api;
}

#define NCCL_CHECK(ctx, call)                                                                                   \
    do {                                                                                                        \
        ncclResult_t r__ = (call);                                                                              \
        if (r__ != ncclSuccess)                                                                                 \
            return mimc3cu_fail(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(r__)); \
    } while (0)

__global__ void or_rows_kernel(uint8_t *dst, const uint8_t *src, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] |= src[i];
}

int ensure_tmp(mimc3cu_ctx *ctx, BandComm *bc, size_t bytes) {
    if (bytes <= bc->tmp_bytes) return 0;
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (bc->tmp) CU_CHECK(ctx, cudaFree(bc->tmp));
    bc->tmp = nullptr; bc->tmp_bytes = 0;
    CU_CHECK(ctx, cudaMalloc(&bc->tmp, bytes));
    bc->tmp_bytes = bytes;
    return 0;
}

// CUDA events around a collective on the context stream (only when mimc3cu_comm_timing was switched on)
struct CommTimer {
    mimc3cu_ctx *ctx; BandComm *bc; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *list; cudaEvent_t stop = nullptr;
    CommTimer(mimc3cu_ctx *c, std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *l) : ctx(c), bc(c->comm), list(l) {
        if (!bc->timing) return;
        auto take = [&] { cudaEvent_t e = nullptr; if (!bc->ev_pool.empty()) { e = bc->ev_pool.back(); bc->ev_pool.pop_back(); } else cudaEventCreate(&e); return e; };
        cudaEvent_t start = take();
        stop = take();
        cudaEventRecord(start, ctx->stream);
        list->push_back({start, stop});
    }
    ~CommTimer() { if (stop) cudaEventRecord(stop, ctx->stream); }
};

}  // namespace

// ---- used by post.cu --------------------------------------------------------------------------------------
int bandcomm_halo_exchange(mimc3cu_ctx *ctx, void *const *arrays, const int32_t *elem_bytes, int32_t count, int dimx, int rows,
                           int own0, int own1, int halo) {
    BandComm *bc = ctx->comm;
    NcclApi *N = nccl_api();
    const bool up = own0 > 0, down = own1 < rows;   // a halo exists <=> a neighbour exists
    if (!up && !down) return 0;
    CommTimer tm(ctx, &bc->ev_exchange);
    NCCL_CHECK(ctx, N->GroupStart());
    for (int32_t k = 0; k < count; k++) {
        char *base = (char *)arrays[k];
        const size_t row = (size_t)dimx * elem_bytes[k], blk = row * halo;
        if (up) {    // my first owned rows -> rank-1's bottom halo; rank-1's last owned rows -> my top halo
            NCCL_CHECK(ctx, N->Send(base + row * own0, blk, ncclChar, bc->rank - 1, (ncclComm_t)bc->comm, ctx->stream));
            NCCL_CHECK(ctx, N->Recv(base + row * (own0 - halo), blk, ncclChar, bc->rank - 1, (ncclComm_t)bc->comm, ctx->stream));
        }
        if (down) {
            NCCL_CHECK(ctx, N->Send(base + row * (own1 - halo), blk, ncclChar, bc->rank + 1, (ncclComm_t)bc->comm, ctx->stream));
            NCCL_CHECK(ctx, N->Recv(base + row * own1, blk, ncclChar, bc->rank + 1, (ncclComm_t)bc->comm, ctx->stream));
        }
    }
    NCCL_CHECK(ctx, N->GroupEnd());
    bc->n_exchanges++;
    return 0;
}

// Flags this band scattered into its halo rows belong to the neighbours: send them there and OR what the neighbours
// scattered into our rows.
int bandcomm_halo_or_reduce(mimc3cu_ctx *ctx, uint8_t *flags, int dimx, int rows, int own0, int own1, int halo) {
    BandComm *bc = ctx->comm;
    NcclApi *N = nccl_api();
    const bool up = own0 > 0, down = own1 < rows;
    if (!up && !down) return 0;
    const size_t blk = (size_t)dimx * halo;
    if (int rc = ensure_tmp(ctx, bc, 2 * blk)) return rc;
    uint8_t *from_up = (uint8_t *)bc->tmp, *from_down = from_up + blk;
    CommTimer tm(ctx, &bc->ev_exchange);
    NCCL_CHECK(ctx, N->GroupStart());
    if (up) {
        NCCL_CHECK(ctx, N->Send(flags + (size_t)dimx * (own0 - halo), blk, ncclChar, bc->rank - 1, (ncclComm_t)bc->comm, ctx->stream));
        NCCL_CHECK(ctx, N->Recv(from_up, blk, ncclChar, bc->rank - 1, (ncclComm_t)bc->comm, ctx->stream));
    }
    if (down) {
        NCCL_CHECK(ctx, N->Send(flags + (size_t)dimx * own1, blk, ncclChar, bc->rank + 1, (ncclComm_t)bc->comm, ctx->stream));
        NCCL_CHECK(ctx, N->Recv(from_down, blk, ncclChar, bc->rank + 1, (ncclComm_t)bc->comm, ctx->stream));
    }
    NCCL_CHECK(ctx, N->GroupEnd());
    const int nb = (int)((blk + 255) / 256);
    if (up) or_rows_kernel<<<nb, 256, 0, ctx->stream>>>(flags + (size_t)dimx * own0, from_up, blk);
    if (down) or_rows_kernel<<<nb, 256, 0, ctx->stream>>>(flags + (size_t)dimx * (own1 - halo), from_down, blk);
    ctx->launches += (up ? 1 : 0) + (down ? 1 : 0);
    CU_CHECK(ctx, cudaGetLastError());
    bc->n_exchanges++;
    return 0;
}

int bandcomm_allreduce_sum(mimc3cu_ctx *ctx, int32_t *dev_vals, int32_t count) {
    BandComm *bc = ctx->comm;
    CommTimer tm(ctx, &bc->ev_allreduce);
    NCCL_CHECK(ctx, nccl_api()->AllReduce(dev_vals, dev_vals, (size_t)count, ncclInt32, ncclSum, (ncclComm_t)bc->comm, ctx->stream));
    bc->n_allreduce++;
    return 0;
}

static int attach(mimc3cu_ctx *ctx, ncclComm_t comm, int rank, int world) {
    if (ctx->comm) return mimc3cu_fail(ctx, "comm: this context already has a communicator");
    BandComm *bc = new BandComm();
    bc->comm = (void *)comm; bc->rank = rank; bc->world = world;
    ctx->comm = bc;
    return 0;
}

extern "C" {

int mimc3cu_comm_unique_id(void *id128) {
    NcclApi *N = nccl_api();
    if (!N->handle || !N->error.empty()) return mimc3cu_fail(nullptr, "comm: %s", N->error.c_str());
    ncclUniqueId id;
    NCCL_CHECK(nullptr, N->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int mimc3cu_comm_init_rank(mimc3cu_ctx *ctx, const void *id128, int32_t rank, int32_t world) {
    NcclApi *N = nccl_api();
    if (!N->handle || !N->error.empty()) return mimc3cu_fail(ctx, "comm: %s", N->error.c_str());
    if (world < 1 || rank < 0 || rank >= world) return mimc3cu_fail(ctx, "comm: bad rank %d of %d", rank, world);
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    NCCL_CHECK(ctx, N->CommInitRank(&comm, world, id, rank));
    return attach(ctx, comm, rank, world);
}

int mimc3cu_comm_init_all(mimc3cu_ctx **ctxs, int32_t n) {
    NcclApi *N = nccl_api();
    if (n < 1 || !ctxs) return mimc3cu_fail(nullptr, "comm: no contexts");
    if (!N->handle || !N->error.empty()) return mimc3cu_fail(ctxs[0], "comm: %s", N->error.c_str());
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    NCCL_CHECK(ctxs[0], N->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++)
        if (int rc = attach(ctxs[i], comms[i], i, n)) return rc;
    return 0;
}

void mimc3cu_comm_destroy(mimc3cu_ctx *ctx) {
    if (!ctx || !ctx->comm) return;
    BandComm *bc = ctx->comm;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (bc->comm) nccl_api()->CommDestroy((ncclComm_t)bc->comm);
    if (bc->tmp) cudaFree(bc->tmp);
    for (auto *l : {&bc->ev_exchange, &bc->ev_allreduce}) for (auto &pr : *l) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : bc->ev_pool) cudaEventDestroy(e);
    delete bc;
    ctx->comm = nullptr;
}

int mimc3cu_comm_info(const mimc3cu_ctx *ctx, int32_t *rank, int32_t *world, int64_t *exchanges, int64_t *allreduces) {
    if (!ctx->comm) return 1;
    if (rank) *rank = ctx->comm->rank;
    if (world) *world = ctx->comm->world;
    if (exchanges) *exchanges = ctx->comm->n_exchanges;
    if (allreduces) *allreduces = ctx->comm->n_allreduce;
    return 0;
}

// Device time of the collectives since the last call: switches the event bracketing on (on = 1) / off (on = 0) and returns
// the summed milliseconds of the halo exchanges (incl. the dirty-flag OR) and of the counter all-reduces.  Synchronises.
int mimc3cu_comm_timing(mimc3cu_ctx *ctx, int32_t on, double *exchange_ms, double *allreduce_ms) {
    BandComm *bc = ctx->comm;
    if (!bc) return mimc3cu_fail(ctx, "comm_timing: no communicator");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    double tot[2] = {0.0, 0.0};
    int k = 0;
    for (auto *l : {&bc->ev_exchange, &bc->ev_allreduce}) {
        for (auto &pr : *l) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) tot[k] += t;
            bc->ev_pool.push_back(pr.first); bc->ev_pool.push_back(pr.second);
        }
        l->clear();
        k++;
    }
    if (exchange_ms) *exchange_ms = tot[0];
    if (allreduce_ms) *allreduce_ms = tot[1];
    bc->timing = on != 0;
    return 0;
}

// Gather of per-band device buffers on rank `root` (the final gather of the five planes): rank r contributes
// bytes[r] bytes from `send`; on root `recv` receives them back to back in rank order.  Asynchronous on the stream.
int mimc3cu_comm_gather(mimc3cu_ctx *ctx, const void *send, const int64_t *bytes, void *recv, int32_t root) {
    BandComm *bc = ctx->comm;
    if (!bc) return mimc3cu_fail(ctx, "comm_gather: no communicator");
    NcclApi *N = nccl_api();
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    NCCL_CHECK(ctx, N->GroupStart());
    if (bc->rank == root) {
        size_t off = 0;
        for (int r = 0; r < bc->world; r++) {
            if (r == root) CU_CHECK(ctx, cudaMemcpyAsync((char *)recv + off, send, (size_t)bytes[r], cudaMemcpyDeviceToDevice, ctx->stream));
            else NCCL_CHECK(ctx, N->Recv((char *)recv + off, (size_t)bytes[r], ncclChar, r, (ncclComm_t)bc->comm, ctx->stream));
            off += (size_t)bytes[r];
        }
    } else {
        NCCL_CHECK(ctx, N->Send(send, (size_t)bytes[bc->rank], ncclChar, root, (ncclComm_t)bc->comm, ctx->stream));
    }
    NCCL_CHECK(ctx, N->GroupEnd());
    return 0;
}

}  // extern "C"
