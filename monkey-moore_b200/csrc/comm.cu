// comm.cu -- multi-GPU result gather of the C-ABI: match lists of all ranks to rank 0 over NCCL.
//
// The reference's only parallelism is "independent file blocks on a thread pool"
// (/root/reference/src/core/search_engine.cpp:104-172): nothing is exchanged between blocks, the
// results are concatenated and sorted (:193-197).  Across GPUs the same holds -- ranks scan
// disjoint block ranges with no data-path collective -- so the one communication step is this
// gather.  Protocol of one mmg_comm_gather call (a collective):
//
//   1. every rank packs the lists of the step into a fixed-capacity buffer [counts | offsets ...]
//      (+ a parallel buffer of table base values) and ONE grouped NCCL send/recv moves the packed
//      buffers to rank 0;
//   2. a rank whose lists do not fit copies the remainder into buffers of its own and sends it at the
//      START OF ITS NEXT CALL into this communicator (the next gather, or mmg_comm_wait);
//   3. rank 0 learns from the gathered headers what did not fit and posts the matching receives, also at
//      the start of its next call (or when the gathered object is used first).  Sends and receives of
//      consecutive gathers therefore pair up in order, nobody reads headers on the critical path of a
//      pipelined scan loop, and -- the point of deferring the sends -- NO SEND EVER WAITS for a peer
//      that is not inside the matching collective call: NCCL serialises the kernels of different
//      communicators on a device, so a send left waiting for a lazily posted receive would stall the
//      collectives of every other communicator (torch.distributed's, say) and deadlock the job.
//
// Everything runs on a dedicated, process-wide gather stream of the device: the scan streams never wait for NCCL, and a
// list that is still being read by a send is released in gather-stream order (the results object is
// told to free itself on this stream), so the stream-ordered allocator cannot recycle it early.
// NCCL is loaded with dlopen so that single-GPU users (the GUI drop-in) need no NCCL installation.
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
extern "C" void mmg_internal_results_free_on(const mmg_results *r, void *stream);
extern "C" void *mmg_internal_gather_stream(void);
extern "C" void mmg_internal_set_error(const char *msg);
extern "C" void mmg_internal_comm_alive(int delta);

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
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    LOAD(Send); LOAD(Recv); LOAD(AllGather); LOAD(GetErrorString);
#undef LOAD
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
             api.Recv && api.AllGather && api.GetErrorString;
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

const int CHAIN_MAP = 128;   // bytes of a slice map on the wire (mmg_program_max_jump <= 128 entry phases)
const int HDR = 64;      // header entries per rank (list counts); nlists <= 60

}  // namespace

struct mmg_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uint64_t cap = 0;               // entries per packed buffer (header included)
    cudaStream_t stream = nullptr;  // every copy and NCCL call of the gather runs here (the device's process-wide gather stream)
    cudaEvent_t t0 = nullptr, t1 = nullptr, hdr_ready = nullptr;   // device time of the last gather; headers landed
    bool timed = false;
    uint64_t *pack_off = nullptr;
    uint32_t *pack_val = nullptr;
    uint64_t *recv_off = nullptr;   // rank 0: world * cap
    uint32_t *recv_val = nullptr;
    uint64_t *hdr_host = nullptr;   // pinned, rank 0: the headers of all ranks (HDR entries per rank)
    mmg_gathered *inflight = nullptr;   // rank 0: the gather whose headers have not been read yet
    // ranks != 0: what the last gather could not pack, copied into owned buffers, sent at the start of the next call
    struct Deferred { uint64_t *off; uint32_t *val; uint64_t n; };
    std::vector<Deferred> deferred;
    // mmg_comm_search: the slice maps of all ranks (CHAIN_MAP bytes each); [world] is this rank's own map
    uint8_t *chain_host = nullptr;  // pinned
    uint8_t *chain_dev = nullptr;
};

struct mmg_gathered {
    int nlists = 0;
    std::vector<uint64_t> counts;                 // per list: total over ranks
    // device pieces per list in rank order: (pointer, n) pairs into recv buffers / spill buffers
    struct Piece { const uint64_t *off; const uint32_t *val; uint64_t n; };
    std::vector<std::vector<Piece>> pieces;
    std::vector<void *> owned;                    // spill buffers (stream-ordered allocations on the gather stream)
    cudaStream_t stream = nullptr;
    cudaEvent_t landed = nullptr;                 // every piece has arrived on rank 0
    bool waited = false;
    // deferred part (rank 0): the other ranks' headers are read on first use; the receives of their overflow are posted
    // at the next collective point of the communicator (next gather / mmg_comm_wait) or by a blocking accessor
    bool pending = false;                         // headers not read yet
    bool posted = true;                           // receives of the overflow posted (nothing to post until the headers say so)
    bool abandoned = false;                       // freed by the caller before it was complete: the communicator finishes and deletes it
    struct Spill { int rank; uint64_t *off; uint32_t *val; uint64_t n; };
    std::vector<Spill> spills;
    mmg_comm *comm = nullptr;
    int error = MMG_OK;
};

namespace {

// rank 0, stage 1: read the gathered headers of ranks 1.. and build their piece lists (host wait for the packed buffers;
// no NCCL call, safe at any time)
int read_headers(mmg_gathered *g) {
    if (!g->pending) return g->error;
    g->pending = false;
    mmg_comm *c = g->comm;
    cudaStream_t stream = g->stream;
    const int nlists = g->nlists;
    const uint64_t room = c->cap - (uint64_t)nlists;
    auto bail = [&](int code) { g->error = code; return code; };
    if (cudaEventSynchronize(c->hdr_ready) != cudaSuccess) return bail(err(MMG_ERR_CUDA, "waiting for the gathered headers failed"));
    const uint64_t *all = c->hdr_host;
    for (int r = 1; r < c->world; r++) {
        uint64_t pos = nlists, used_r = 0;
        for (int k = 0; k < nlists; k++) {
            const uint64_t n = all[(size_t)r * HDR + k];
            const uint64_t f = std::min<uint64_t>(n, room - used_r);
            g->counts[k] += n;
            if (f) g->pieces[k].push_back({c->recv_off + (size_t)r * c->cap + pos, c->recv_val + (size_t)r * c->cap + pos, f});
            pos += f; used_r += f;
            if (f < n) {
                const uint64_t rest = n - f;
                uint64_t *so = nullptr; uint32_t *sv = nullptr;
                if (cudaMallocAsync((void **)&so, rest * sizeof(uint64_t), stream) != cudaSuccess) return bail(err(MMG_ERR_NOMEM, "spill buffer allocation failed"));
                g->owned.push_back(so);
                if (cudaMallocAsync((void **)&sv, rest * sizeof(uint32_t), stream) != cudaSuccess) return bail(err(MMG_ERR_NOMEM, "spill buffer allocation failed"));
                g->owned.push_back(sv);
                g->spills.push_back({r, so, sv, rest});
                g->pieces[k].push_back({so, sv, rest});
            }
        }
    }
    g->posted = false;
    return MMG_OK;
}

void delete_gathered(mmg_gathered *g) {
    for (void *p : g->owned) cudaFreeAsync(p, g->stream);     // behind the receives that fill them
    if (g->landed) cudaEventDestroy(g->landed);
    delete g;
}

// rank 0, stage 2: post the receives of what the other ranks set aside (they send it at the start of their next call).
// Only at a collective point of the communicator, or from an accessor that then blocks until the data has landed.
int complete_gather(mmg_gathered *g) {
    if (int rc = read_headers(g); rc != MMG_OK) return rc;
    if (g->posted) return g->error;
    g->posted = true;
    mmg_comm *c = g->comm;
    cudaStream_t stream = g->stream;
    auto bail = [&](int code) { g->error = code; return code; };
    if (c->inflight == g) c->inflight = nullptr;
    if (!g->spills.empty()) {
        // same order per peer as the sends (list by list, offsets then values); one group so that all peers stream at once
        if (nccl().GroupStart() != ncclSuccess) return bail(err(MMG_ERR_CUDA, "ncclGroupStart failed"));
        for (const auto &sp : g->spills) {
            if (nccl().Recv(sp.off, sp.n, ncclUint64, sp.rank, c->comm, stream) != ncclSuccess ||
                nccl().Recv(sp.val, sp.n, ncclUint32, sp.rank, c->comm, stream) != ncclSuccess)
                return bail(err(MMG_ERR_CUDA, "receiving a spilled list failed"));
        }
        if (nccl().GroupEnd() != ncclSuccess) return bail(err(MMG_ERR_CUDA, "ncclGroupEnd failed"));
    }
    if (cudaEventRecord(c->t1, stream) != cudaSuccess) return bail(err(MMG_ERR_CUDA, "cudaEventRecord failed"));
    c->timed = true;
    if (cudaEventRecord(g->landed, stream) != cudaSuccess) return bail(err(MMG_ERR_CUDA, "cudaEventRecord failed"));
    return MMG_OK;
}

// collective point of the communicator on rank 0: finish the gather that is still open
void finish_inflight(mmg_comm *c) {
    mmg_gathered *g = c->inflight;
    if (!g) return;
    complete_gather(g);
    c->inflight = nullptr;
    if (g->abandoned) delete_gathered(g);
}

// ranks != 0: send what the previous gather deferred (rank 0 posts the matching receives at the same point of ITS call)
int flush_deferred(mmg_comm *c) {
    if (c->deferred.empty()) return MMG_OK;
    cudaStream_t stream = c->stream;
    NC(nccl().GroupStart());
    for (const auto &d : c->deferred) {
        NC(nccl().Send(d.off, d.n, ncclUint64, 0, c->comm, stream));
        NC(nccl().Send(d.val, d.n, ncclUint32, 0, c->comm, stream));
    }
    NC(nccl().GroupEnd());
    for (const auto &d : c->deferred) { cudaFreeAsync(d.off, stream); cudaFreeAsync(d.val, stream); }
    c->deferred.clear();
    return MMG_OK;
}

void destroy_comm(mmg_comm *c) {
    if (!c) return;
    for (const auto &d : c->deferred) { cudaFreeAsync(d.off, c->stream); cudaFreeAsync(d.val, c->stream); }   // never sent: dropped
    c->deferred.clear();
    if (c->inflight) {                     // nobody will send its overflow any more: read the headers, drop the rest
        mmg_gathered *g = c->inflight;
        read_headers(g);
        g->posted = true;
        g->comm = nullptr;
        c->inflight = nullptr;
        if (g->abandoned) delete_gathered(g);
    }
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    cudaFree(c->pack_off); cudaFree(c->pack_val); cudaFree(c->recv_off); cudaFree(c->recv_val);
    if (c->hdr_host) cudaFreeHost(c->hdr_host);
    if (c->chain_host) cudaFreeHost(c->chain_host);
    cudaFree(c->chain_dev);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->hdr_ready) cudaEventDestroy(c->hdr_ready);
    delete c;       // the stream belongs to the device state: lists released after this point still name it
}

int wait_landed(mmg_gathered *g) {
    if (int rc = complete_gather(g); rc != MMG_OK) return rc;
    if (g->waited) return MMG_OK;
    if (cudaEventSynchronize(g->landed) != cudaSuccess) return err(MMG_ERR_CUDA, "waiting for the gathered lists failed");
    g->waited = true;
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
    *out = nullptr;
    if (!nccl().ok) return err(MMG_ERR_CUDA, "NCCL library not available (dlopen libnccl.so.2 failed)");
    mmg_comm *c = new mmg_comm();
    c->rank = rank; c->world = world; c->cap = std::max<uint64_t>(capacity, 2 * HDR);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    auto build = [&]() -> int {
        NC(nccl().CommInitRank(&c->comm, world, id, rank));
        c->stream = static_cast<cudaStream_t>(mmg_internal_gather_stream());
        if (!c->stream) return err(MMG_ERR_CUDA, "no usable CUDA device");
        CUC(cudaEventCreate(&c->t0));
        CUC(cudaEventCreate(&c->t1));
        CUC(cudaEventCreateWithFlags(&c->hdr_ready, cudaEventDisableTiming));
        CUC(cudaMalloc((void **)&c->pack_off, c->cap * sizeof(uint64_t)));
        CUC(cudaMalloc((void **)&c->pack_val, c->cap * sizeof(uint32_t)));
        if (rank == 0) {
            CUC(cudaMalloc((void **)&c->recv_off, (size_t)world * c->cap * sizeof(uint64_t)));
            CUC(cudaMalloc((void **)&c->recv_val, (size_t)world * c->cap * sizeof(uint32_t)));
        }
        CUC(cudaHostAlloc((void **)&c->hdr_host, (size_t)world * HDR * sizeof(uint64_t), cudaHostAllocDefault));
        return MMG_OK;
    };
    const int rc = build();
    if (rc != MMG_OK) { destroy_comm(c); return rc; }
    mmg_internal_comm_alive(+1);
    *out = c;
    return MMG_OK;
}

void mmg_comm_destroy(mmg_comm *c) {
    if (c) mmg_internal_comm_alive(-1);
    destroy_comm(c);
}

// Gathers `nlists` result lists (the searches of one step) of every rank to rank 0.
// On rank 0 *out receives the gathered lists (rank order == ascending file offsets); elsewhere NULL.
// Every rank only ENQUEUES work on the gather stream (rank 0 first completes its previous gather, whose headers have
// long arrived by then); rank 0 reads the other ranks' headers and posts the receives of whatever did not fit when the
// gathered object is first used.  The scan streams are never involved.
int mmg_comm_gather(mmg_comm *c, const mmg_results *const *lists, int nlists, mmg_gathered **out) {
    if (!c || !lists || nlists <= 0 || nlists > HDR - 4 || !out) return err(MMG_ERR_ARG, "bad gather arguments");
    *out = nullptr;
    // the previous gather's headers live in the buffer this one reuses, and its overflow receives must be posted before
    // this gather's receives (sends and receives of a pair of ranks match in order)
    finish_inflight(c);
    if (int rc = flush_deferred(c); rc != MMG_OK) return rc;
    cudaStream_t stream = c->stream;
    const uint64_t room = c->cap - (uint64_t)nlists;
    std::vector<mmg_results_view> v(nlists);
    for (int k = 0; k < nlists; k++) mmg_internal_results_view(lists[k], &v[k]);      // completes the scans: counts are final

    CUC(cudaEventRecord(c->t0, stream));
    // pack: header (counts) + as much of every list as fits.  Rank 0 packs straight into its receive slot.
    uint64_t *dst_off = c->rank == 0 ? c->recv_off : c->pack_off;
    uint32_t *dst_val = c->rank == 0 ? c->recv_val : c->pack_val;
    uint64_t hdr[HDR];
    uint64_t at = nlists, used = 0;
    std::vector<uint64_t> fit(nlists);
    for (int k = 0; k < nlists; k++) {
        hdr[k] = v[k].count;
        fit[k] = std::min<uint64_t>(v[k].count, room - used);
        if (fit[k]) {
            CUC(cudaMemcpyAsync(dst_off + at, v[k].d_off, fit[k] * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
            CUC(cudaMemcpyAsync(dst_val + at, v[k].d_val, fit[k] * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
        }
        at += fit[k]; used += fit[k];
    }
    // pageable source: the runtime stages it before returning, so `hdr` may go out of scope
    CUC(cudaMemcpyAsync(dst_off, hdr, (size_t)nlists * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));

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
        // what did not fit is copied aside (the lists may be freed right after this call) and sent at the next call
        for (int k = 0; k < nlists; k++) {
            if (fit[k] < v[k].count) {
                const uint64_t rest = v[k].count - fit[k];
                uint64_t *so = nullptr; uint32_t *sv = nullptr;
                CUC(cudaMallocAsync((void **)&so, rest * sizeof(uint64_t), stream));
                if (cudaMallocAsync((void **)&sv, rest * sizeof(uint32_t), stream) != cudaSuccess) {
                    cudaFreeAsync(so, stream);
                    return err(MMG_ERR_NOMEM, "spill buffer allocation failed");
                }
                c->deferred.push_back({so, sv, rest});
                CUC(cudaMemcpyAsync(so, v[k].d_off + fit[k], rest * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
                CUC(cudaMemcpyAsync(sv, v[k].d_val + fit[k], rest * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
            }
        }
        CUC(cudaEventRecord(c->t1, stream));
        c->timed = true;
        // the copies and sends above read the lists on THIS stream: release them in this stream's order
        for (int k = 0; k < nlists; k++) mmg_internal_results_free_on(lists[k], stream);
        return MMG_OK;
    }

    // ---- rank 0: its own lists are complete here (what did not fit is copied now, while the lists are known to be alive);
    // the other ranks' headers travel to the host and are read when the gathered object is first used
    mmg_gathered *g = new mmg_gathered();
    g->nlists = nlists;
    g->counts.assign(nlists, 0);
    g->pieces.resize(nlists);
    g->stream = stream;
    g->comm = c;
    auto bail = [&](int code) { mmg_gathered_free(g); return code; };
    {
        uint64_t pos = nlists;
        for (int k = 0; k < nlists; k++) {
            const uint64_t n = v[k].count, f = fit[k];
            g->counts[k] += n;
            if (f) g->pieces[k].push_back({c->recv_off + pos, c->recv_val + pos, f});
            pos += f;
            if (f < n) {
                const uint64_t rest = n - f;
                uint64_t *so = nullptr; uint32_t *sv = nullptr;
                if (cudaMallocAsync((void **)&so, rest * sizeof(uint64_t), stream) != cudaSuccess) return bail(err(MMG_ERR_NOMEM, "spill buffer allocation failed"));
                g->owned.push_back(so);
                if (cudaMallocAsync((void **)&sv, rest * sizeof(uint32_t), stream) != cudaSuccess) return bail(err(MMG_ERR_NOMEM, "spill buffer allocation failed"));
                g->owned.push_back(sv);
                if (cudaMemcpyAsync(so, v[k].d_off + f, rest * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
                    cudaMemcpyAsync(sv, v[k].d_val + f, rest * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
                    return bail(err(MMG_ERR_CUDA, "copying rank 0's overflow failed"));
                g->pieces[k].push_back({so, sv, rest});
            }
        }
    }
    if (cudaEventCreateWithFlags(&g->landed, cudaEventDisableTiming) != cudaSuccess) return bail(err(MMG_ERR_CUDA, "cudaEventCreate failed"));
    if (c->world > 1) {
        if (cudaMemcpy2DAsync(c->hdr_host, HDR * sizeof(uint64_t), c->recv_off, c->cap * sizeof(uint64_t),
                              (size_t)nlists * sizeof(uint64_t), c->world, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaEventRecord(c->hdr_ready, stream) != cudaSuccess)
            return bail(err(MMG_ERR_CUDA, "reading the gathered headers back failed"));
        g->pending = true;
        c->inflight = g;
    } else {
        if (cudaEventRecord(c->t1, stream) != cudaSuccess || cudaEventRecord(g->landed, stream) != cudaSuccess)
            return bail(err(MMG_ERR_CUDA, "cudaEventRecord failed"));
        c->timed = true;
    }
    for (int k = 0; k < nlists; k++) mmg_internal_results_free_on(lists[k], stream);
    *out = g;
    return MMG_OK;
}

// Blocks until everything this rank enqueued for its gathers has executed (sends delivered / lists landed).
// *ms_last (may be NULL) receives the device time of the last gather on this rank's gather stream.
int mmg_comm_wait(mmg_comm *c, float *ms_last) {
    if (!c) return err(MMG_ERR_ARG, "null communicator");
    finish_inflight(c);
    if (int rc = flush_deferred(c); rc != MMG_OK) return rc;
    CUC(cudaStreamSynchronize(c->stream));
    if (ms_last) {
        *ms_last = 0.f;
        if (c->timed && cudaEventElapsedTime(ms_last, c->t0, c->t1) != cudaSuccess) { cudaGetLastError(); *ms_last = 0.f; }
    }
    return MMG_OK;
}

// One chain over a buffer spread over the ranks: see include/mmoore_b200.h.  The all-gather of the slice maps runs on
// the gather stream like every other NCCL call of the communicator; the host waits for it (the entry phase is needed
// to enqueue the second half of the scan), which costs one round trip of a few dozen microseconds per search.
int mmg_comm_search(mmg_comm *c, const mmg_program *p, const void *data, uint64_t owned_len, uint64_t avail_len,
                    int mem, uint64_t first_element, mmg_results **out) {
    if (!c || !p || !out) return err(MMG_ERR_ARG, "bad search arguments");
    *out = nullptr;
    finish_inflight(c);                                   // same entry discipline as a gather: pending receives and
    if (int rc = flush_deferred(c); rc != MMG_OK) return rc;     // deferred sends go first, in the same order on every rank
    const size_t M = CHAIN_MAP;
    if (!c->chain_host) {
        CUC(cudaHostAlloc((void **)&c->chain_host, (size_t)(c->world + 1) * M, cudaHostAllocDefault));
        CUC(cudaMalloc((void **)&c->chain_dev, (size_t)(c->world + 1) * M));
    }
    mmg_results *r = nullptr;
    int rc = mmg_chain_begin(p, data, owned_len, avail_len, mem, first_element, &r);
    uint8_t *mine = c->chain_host + (size_t)c->world * M;
    std::memset(mine, 0xFF, M);                           // 0xFF in entry 0 tells the others that this rank failed
    int n = 0;
    if (rc == MMG_OK) rc = mmg_chain_map(r, mine, (int)M, &n);
    if (rc != MMG_OK) std::memset(mine, 0xFF, M);
    // every rank takes part in the exchange even after a local failure, so that nobody is left waiting
    cudaStream_t stream = c->stream;
    auto exchange = [&]() -> int {
        CUC(cudaMemcpyAsync(c->chain_dev + (size_t)c->world * M, mine, M, cudaMemcpyHostToDevice, stream));
        NC(nccl().AllGather(c->chain_dev + (size_t)c->world * M, c->chain_dev, M, ncclUint8, c->comm, stream));
        CUC(cudaMemcpyAsync(c->chain_host, c->chain_dev, (size_t)c->world * M, cudaMemcpyDeviceToHost, stream));
        CUC(cudaStreamSynchronize(stream));
        return MMG_OK;
    };
    const int xrc = exchange();
    if (rc == MMG_OK && xrc != MMG_OK) rc = xrc;
    if (rc == MMG_OK)
        for (int k = 0; k < c->world; k++)
            if (c->chain_host[(size_t)k * M] == 0xFF) { rc = err(MMG_ERR_CUDA, "the slice scan of another rank failed"); break; }
    if (rc == MMG_OK) rc = mmg_chain_finish(r, mmg_chain_entry(c->chain_host, (int)M, c->rank));
    if (rc != MMG_OK) { if (r) mmg_results_free(r); return rc; }
    *out = r;
    return MMG_OK;
}

uint64_t mmg_gathered_count(const mmg_gathered *g, int list) {
    if (!g || list < 0 || list >= g->nlists) return 0;
    read_headers(const_cast<mmg_gathered *>(g));            // counts need the headers only
    return g->counts[list];
}

// Copies list `list` (all ranks, ascending file offsets) to host buffers: offsets[count], values[2*count].
// Valid until the next mmg_comm_gather on the same communicator.
int mmg_gathered_copy(const mmg_gathered *g, int list, uint64_t *offsets, uint32_t *values) {
    if (!g || list < 0 || list >= g->nlists) return err(MMG_ERR_ARG, "bad gathered list");
    if (int rc = wait_landed(const_cast<mmg_gathered *>(g)); rc != MMG_OK) return rc;
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

// Device-resident pieces of list `list` in file order (rank 0): up to `cap` (pointer, pointer, n) triples;
// returns the number of pieces.  Valid until the next mmg_comm_gather on the same communicator.
int mmg_gathered_pieces(const mmg_gathered *g, int list, const uint64_t **offs, const uint32_t **vals, uint64_t *ns, int cap) {
    if (!g || list < 0 || list >= g->nlists) return 0;
    if (wait_landed(const_cast<mmg_gathered *>(g)) != MMG_OK) return 0;
    int n = 0;
    for (const auto &p : g->pieces[list]) {
        if (n < cap) { if (offs) offs[n] = p.off; if (vals) vals[n] = p.val; if (ns) ns[n] = p.n; }
        n++;
    }
    return n;
}

void mmg_gathered_free(mmg_gathered *g) {
    if (!g) return;
    // still open (headers unread or overflow not received): the senders will send their overflow at their next call all
    // the same, so the communicator finishes this gather at ITS next collective point and deletes the object then
    if (g->comm && g->comm->inflight == g && (g->pending || !g->posted)) { g->abandoned = true; return; }
    delete_gathered(g);
}

}  // extern "C"
