// comm.cu -- multi-GPU result gather of the C-ABI: match lists of all ranks to rank 0 over NCCL.
//
// The reference's only parallelism is "independent file blocks on a thread pool"
// (/root/reference/src/core/search_engine.cpp:104-172): nothing is exchanged between blocks, the
// results are concatenated and sorted (:193-197).  Across GPUs the same holds -- ranks scan
// disjoint block ranges with no data-path collective -- so the one communication step is this
// gather.  It is ONE grouped NCCL operation per call: every rank packs the lists of a step into a
// fixed-capacity buffer [counts | offsets ...] (+ a parallel buffer of table base values) and
// sends it to rank 0; a rank whose lists do not fit sends the remainder point-to-point, which
// rank 0 learns from the gathered header.  NCCL is loaded with dlopen so that single-GPU users
// (the GUI drop-in) need no NCCL installation.
#include "../../include/mmoore_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

// from capi.cu
struct mmg_results_view { uint64_t count; const uint64_t *d_off; const uint32_t *d_val; };
extern "C" int mmg_internal_results_view(const mmg_results *r, mmg_results_view *out);
extern "C" void *mmg_internal_stream(void);
extern "C" void mmg_internal_set_error(const char *msg);

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // prefer a libnccl that is already loaded (torch bundles one), then the system one
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) return api;
#define LOAD(name) *(void **)(&api.name) = dlsym(api.lib, "nccl" #name)
    LOAD(GetUniqueId); LOAD(CommInitRank); LOAD(CommDestroy); LOAD(GroupStart); LOAD(GroupEnd);
    LOAD(Send); LOAD(Recv); LOAD(GetErrorString);
#undef LOAD
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
             api.Recv && api.GetErrorString;
    return api;
}

int err(int code, const std::string &msg) {
    mmg_internal_set_error(msg.c_str());
    return code;
}

#define NC(call)                                                                                  \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess) return err(MMG_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r_)); \
    } while (0)
#define CUC(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return err(MMG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

}  // namespace

struct mmg_gathered;

struct mmg_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uint64_t cap = 0;            // entries per packed buffer (header included)
    uint64_t *pack_off = nullptr;
    uint32_t *pack_val = nullptr;
    uint64_t *recv_off = nullptr;   // rank 0: world * cap
    uint32_t *recv_val = nullptr;
    uint64_t *hdr_host = nullptr;   // pinned, rank 0: the headers of all ranks (64 entries per rank)
    mmg_gathered *inflight = nullptr;   // rank 0: a gather whose headers have not been read yet
};

struct mmg_gathered {
    int nlists = 0;
    std::vector<uint64_t> counts;                 // per list: total over ranks
    // device pieces per list in rank order: (pointer, n) pairs into recv buffers / spill buffers
    struct Piece { const uint64_t *off; const uint32_t *val; uint64_t n; };
    std::vector<std::vector<Piece>> pieces;
    std::vector<void *> owned;                    // spill buffers to free
    // deferred completion (rank 0): headers are read and spills received on first use
    bool pending = false;
    mmg_comm *comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr;                  // headers have landed in hdr_host
    std::vector<mmg_results_view> own;            // rank 0's own lists (their overflow is copied on completion)
    int error = MMG_OK;
};

namespace {

// rank 0: read the gathered headers, build the piece lists, receive what did not fit into the packed buffers
int finish_gather(mmg_gathered *g) {
    if (!g->pending) return g->error;
    g->pending = false;
    mmg_comm *c = g->comm;
    if (c->inflight == g) c->inflight = nullptr;
    cudaStream_t stream = g->stream;
    const int nlists = g->nlists;
    const uint64_t room = c->cap - (uint64_t)nlists;
    auto bail = [&](int code) { g->error = code; return code; };
    if (cudaEventSynchronize(g->ready) != cudaSuccess) return bail(err(MMG_ERR_CUDA, "waiting for the gathered headers failed"));
    cudaEventDestroy(g->ready);
    g->ready = nullptr;
    const uint64_t *all = c->hdr_host;
    bool spilled = false;
    for (int r = 0; r < c->world; r++) {
        uint64_t pos = nlists, used_r = 0;
        for (int k = 0; k < nlists; k++) {
            const uint64_t n = all[(size_t)r * 64 + k];
            const uint64_t f = std::min<uint64_t>(n, room - used_r);
            g->counts[k] += n;
            if (f) g->pieces[k].push_back({c->recv_off + (size_t)r * c->cap + pos, c->recv_val + (size_t)r * c->cap + pos, f});
            pos += f; used_r += f;
            if (f < n) {
                const uint64_t rest = n - f;
                uint64_t *so; uint32_t *sv;
                if (cudaMalloc((void **)&so, rest * sizeof(uint64_t)) != cudaSuccess ||
                    cudaMalloc((void **)&sv, rest * sizeof(uint32_t)) != cudaSuccess)
                    return bail(err(MMG_ERR_NOMEM, "spill buffer allocation failed"));
                g->owned.push_back(so); g->owned.push_back(sv);
                if (r == 0) {
                    if (cudaMemcpyAsync(so, g->own[k].d_off + f, rest * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
                        cudaMemcpyAsync(sv, g->own[k].d_val + f, rest * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
                        return bail(err(MMG_ERR_CUDA, "copying rank 0's overflow failed"));
                } else {
                    if (nccl().Recv(so, rest, ncclUint64, r, c->comm, stream) != ncclSuccess ||
                        nccl().Recv(sv, rest, ncclUint32, r, c->comm, stream) != ncclSuccess)
                        return bail(err(MMG_ERR_CUDA, "receiving a spilled list failed"));
                }
                g->pieces[k].push_back({so, sv, rest});
                spilled = true;
            }
        }
    }
    if (spilled && cudaStreamSynchronize(stream) != cudaSuccess) return bail(err(MMG_ERR_CUDA, "spill transfer failed"));
    return MMG_OK;
}

}  // namespace

extern "C" {

int mmg_comm_unique_id(void *out128) {
    if (!nccl().ok) return err(MMG_ERR_CUDA, "NCCL library not available (dlopen libnccl.so.2 failed)");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(nccl().GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
    return MMG_OK;
}

int mmg_comm_create(const void *id128, int rank, int world, uint64_t capacity, mmg_comm **out) {
    if (!out || !id128 || rank < 0 || rank >= world) return err(MMG_ERR_ARG, "bad communicator arguments");
    if (!nccl().ok) return err(MMG_ERR_CUDA, "NCCL library not available (dlopen libnccl.so.2 failed)");
    mmg_comm *c = new mmg_comm();
    c->rank = rank; c->world = world; c->cap = std::max<uint64_t>(capacity, 64);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    NC(nccl().CommInitRank(&c->comm, world, id, rank));
    CUC(cudaMalloc((void **)&c->pack_off, c->cap * sizeof(uint64_t)));
    CUC(cudaMalloc((void **)&c->pack_val, c->cap * sizeof(uint32_t)));
    if (rank == 0) {
        CUC(cudaMalloc((void **)&c->recv_off, (size_t)world * c->cap * sizeof(uint64_t)));
        CUC(cudaMalloc((void **)&c->recv_val, (size_t)world * c->cap * sizeof(uint32_t)));
    }
    CUC(cudaHostAlloc((void **)&c->hdr_host, (size_t)world * 64 * sizeof(uint64_t), cudaHostAllocDefault));
    *out = c;
    return MMG_OK;
}

void mmg_comm_destroy(mmg_comm *c) {
    if (!c) return;
    if (c->inflight) finish_gather(c->inflight);
    if (c->comm) nccl().CommDestroy(c->comm);
    cudaFree(c->pack_off); cudaFree(c->pack_val); cudaFree(c->recv_off); cudaFree(c->recv_val);
    cudaFreeHost(c->hdr_host);
    delete c;
}

// Gathers `nlists` result lists (the searches of one step) of every rank to rank 0.
// On rank 0 *out receives the gathered lists (rank order == ascending file offsets); elsewhere NULL.
// The call only ENQUEUES work on the scan stream; rank 0 completes it (header read, spill receives) when
// the gathered object is first used or freed.  If rank 0's own lists may overflow the packed buffer they
// must stay alive until then.
int mmg_comm_gather(mmg_comm *c, const mmg_results *const *lists, int nlists, mmg_gathered **out) {
    if (!c || !lists || nlists <= 0 || nlists > 60 || !out) return err(MMG_ERR_ARG, "bad gather arguments");
    *out = nullptr;
    if (c->inflight) finish_gather(c->inflight);          // its headers live in the buffer we are about to reuse
    cudaStream_t stream = static_cast<cudaStream_t>(mmg_internal_stream());
    const uint64_t room = c->cap - (uint64_t)nlists;
    std::vector<mmg_results_view> v(nlists);
    for (int k = 0; k < nlists; k++) mmg_internal_results_view(lists[k], &v[k]);

    // pack: header (counts) + as much of every list as fits
    uint64_t hdr[64];
    uint64_t at = nlists, used = 0;
    std::vector<uint64_t> fit(nlists);
    for (int k = 0; k < nlists; k++) {
        hdr[k] = v[k].count;
        fit[k] = std::min<uint64_t>(v[k].count, room - used);
        if (fit[k]) {
            CUC(cudaMemcpyAsync(c->pack_off + at, v[k].d_off, fit[k] * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
            CUC(cudaMemcpyAsync(c->pack_val + at, v[k].d_val, fit[k] * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
        }
        at += fit[k]; used += fit[k];
    }
    // pageable source: the runtime stages it before returning, so `hdr` may go out of scope
    CUC(cudaMemcpyAsync(c->pack_off, hdr, (size_t)nlists * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));

    // one grouped NCCL operation moves every rank's packed buffers to rank 0
    NC(nccl().GroupStart());
    if (c->rank != 0) {
        NC(nccl().Send(c->pack_off, c->cap, ncclUint64, 0, c->comm, stream));
        NC(nccl().Send(c->pack_val, c->cap, ncclUint32, 0, c->comm, stream));
    } else {
        for (int r = 1; r < c->world; r++) {
            NC(nccl().Recv(c->recv_off + (size_t)r * c->cap, c->cap, ncclUint64, r, c->comm, stream));
            NC(nccl().Recv(c->recv_val + (size_t)r * c->cap, c->cap, ncclUint32, r, c->comm, stream));
        }
    }
    NC(nccl().GroupEnd());

    if (c->rank != 0) {
        // what did not fit travels point-to-point; rank 0 posts the matching receives when it completes the gather.
        // Stream order keeps the lists alive: their cudaFreeAsync is queued behind these sends.
        for (int k = 0; k < nlists; k++) {
            if (fit[k] < v[k].count) {
                NC(nccl().Send(v[k].d_off + fit[k], v[k].count - fit[k], ncclUint64, 0, c->comm, stream));
                NC(nccl().Send(v[k].d_val + fit[k], v[k].count - fit[k], ncclUint32, 0, c->comm, stream));
            }
        }
        return MMG_OK;
    }

    // rank 0: own buffers + everybody's headers (read back asynchronously)
    CUC(cudaMemcpyAsync(c->recv_off, c->pack_off, c->cap * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
    CUC(cudaMemcpyAsync(c->recv_val, c->pack_val, c->cap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    CUC(cudaMemcpy2DAsync(c->hdr_host, 64 * sizeof(uint64_t), c->recv_off, c->cap * sizeof(uint64_t),
                          (size_t)nlists * sizeof(uint64_t), c->world, cudaMemcpyDeviceToHost, stream));
    mmg_gathered *g = new mmg_gathered();
    g->nlists = nlists;
    g->counts.assign(nlists, 0);
    g->pieces.resize(nlists);
    g->pending = true;
    g->comm = c;
    g->stream = stream;
    g->own = v;
    CUC(cudaEventCreateWithFlags(&g->ready, cudaEventDisableTiming));
    CUC(cudaEventRecord(g->ready, stream));
    c->inflight = g;
    *out = g;
    return MMG_OK;
}

uint64_t mmg_gathered_count(const mmg_gathered *g, int list) {
    if (!g || list < 0 || list >= g->nlists) return 0;
    finish_gather(const_cast<mmg_gathered *>(g));
    return g->counts[list];
}

// Copies list `list` (all ranks, ascending file offsets) to host buffers: offsets[count], values[2*count].
// Valid until the next mmg_comm_gather on the same communicator.
int mmg_gathered_copy(const mmg_gathered *g, int list, uint64_t *offsets, uint32_t *values) {
    if (!g || list < 0 || list >= g->nlists) return err(MMG_ERR_ARG, "bad gathered list");
    if (int rc = finish_gather(const_cast<mmg_gathered *>(g)); rc != MMG_OK) return rc;
    uint64_t at = 0;
    for (const auto &p : g->pieces[list]) {
        if (offsets) CUC(cudaMemcpy(offsets + at, p.off, p.n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (values) {
            std::vector<uint32_t> packed(p.n);
            CUC(cudaMemcpy(packed.data(), p.val, p.n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < p.n; i++) { values[2 * (at + i)] = packed[i] & 0xFFFFu; values[2 * (at + i) + 1] = packed[i] >> 16; }
        }
        at += p.n;
    }
    return MMG_OK;
}

void mmg_gathered_free(mmg_gathered *g) {
    if (!g) return;
    finish_gather(g);
    if (g->ready) cudaEventDestroy(g->ready);
    for (void *p : g->owned) cudaFree(p);
    delete g;
}

}  // extern "C"
