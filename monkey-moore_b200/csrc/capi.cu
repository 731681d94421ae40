// capi.cu -- the C-ABI (include/mmoore_b200.h): host orchestration of the scan kernels.
//
// Host-side counterparts in the reference (file:line under /root/reference/):
//   MonkeyMoore<Ty>::search                      src/core/monkey_moore.cpp:41-49
//   SearchEngine<T>::compute_search_blocks       src/core/search_engine.cpp:218-253
//   SearchEngine<T>::run worker + merge + sort   src/core/search_engine.cpp:104-172, 193-197
// There is no CPU implementation of the scan in this library: without a CUDA device every
// scan entry point returns MMG_ERR_CUDA.
#include "../../include/mmoore_b200.h"
#include "launch.h"
#include "pattern.hpp"

#include <algorithm>
#include <atomic>
#include <map>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            throw ScanError{e_ == cudaErrorMemoryAllocation ? MMG_ERR_NOMEM : MMG_ERR_CUDA};       \
        }                                                                                          \
    } while (0)

struct ScanError { int code; };

std::atomic<int> g_live_comms{0};   // result-gather communicators alive in this process (comm.cu)
int g_path_override = 0;   // 0 auto, 1 force generic, 2 force evaluate-everything tiles (testing)
std::atomic<int> g_complete{0};   // mmg_set_complete_matches: report every matching window, not only the chain's
thread_local cudaStream_t g_user_stream = nullptr;   // set by mmg_set_stream: scans run on the caller's stream
thread_local bool g_use_user_stream = false;

// Per (device, lane) scan workspace, reused by every tiled scan: nothing is allocated, zeroed or copied per scan in
// the steady state.  `zero` holds the state that must be all-zero when a scan starts (status words, tickets,
// look-back words); the resolve kernel restores it when it ends.  Scans on one stream execute in order, so sharing the
// workspace between scans that are enqueued back to back is safe; `generation` tells a pending scan whether its
// scratch contents (needed only for the exact re-emission) are still there.
struct Workspace {
    cudaStream_t stream = nullptr;
    uint8_t *zero = nullptr;     size_t zero_bytes = 0;
    uint8_t *scratch = nullptr;  size_t scratch_bytes = 0;
    uint32_t *ev = nullptr;      size_t ev_entries = 0;
    uint64_t generation = 0;
    bool dirty = false;          // a failed enqueue may have left the zero state non-zero
};

// Scans run on one of two internal streams ("lanes"), each with its own workspace, ordered after whatever the
// caller's stream holds when the scan is enqueued.  Consecutive scans therefore overlap: the latency-bound resolve
// kernel of one runs beside the bandwidth-bound filter of the next, and a host-input copy beside the previous scan.
struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr;
    Workspace ws;
};

// PROCESS-WIDE state of one device: streams, lanes and their workspaces are created once and shared by every host
// thread (the reference GUI starts a fresh thread per search, src/gui/monkey_frame.cpp:1167 -- per-thread state would
// leak streams and workspaces with every search).  `mu` serialises the enqueue side: the order in which scans reach a
// lane's stream is the order in which they use its workspace.  Objects live until the process exits, so the
// Workspace pointer a pending scan keeps stays valid whichever thread completes or frees it.
struct DeviceInfo {
    int device = -1;
    int sms = 0;
    cudaStream_t own_stream = nullptr;   // this library's non-blocking stream (callers without mmg_set_stream)
    cudaStream_t gather_stream = nullptr; // comm.cu: copies and NCCL calls of the result gathers (lives as long as the process:
                                          // result lists that a gather read are released in this stream's order, possibly
                                          // long after the communicator itself is gone)
    Lane lanes[2];
    unsigned next_lane = 0;
    std::mutex mu;
};

std::mutex g_devices_mutex;
std::deque<DeviceInfo> g_devices;        // deque: stable addresses

DeviceInfo &device_info() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) throw ScanError{fail(MMG_ERR_CUDA, "no usable CUDA device")};
    std::lock_guard<std::mutex> lock(g_devices_mutex);
    for (auto &d : g_devices)
        if (d.device == dev) return d;
    g_devices.emplace_back();
    DeviceInfo &d = g_devices.back();
    try {
        CU(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaStreamCreateWithFlags(&d.own_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&d.gather_stream, cudaStreamNonBlocking));
        for (Lane &l : d.lanes) {
            CU(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
        }
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;   // keep freed scratch in the pool: steady-state scans do not hit cudaMalloc
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    } catch (...) {
        g_devices.pop_back();
        throw;
    }
    d.device = dev;
    return d;
}

// the stream scans of the calling thread are ordered after: its mmg_set_stream stream, or the library's own
cudaStream_t caller_stream(DeviceInfo &d) { return g_use_user_stream ? g_user_stream : d.own_stream; }

// ---- host-side copies into page-locked staging memory --------------------------------------------------------------
// One CPU thread moves 5-15 GB/s, a PCIe 5 x16 link 55 GB/s: copies that feed the link (file pages or pageable user
// buffers -> pinned staging) are split over a small process-wide pool of worker threads.
class CopyPool {
 public:
    static CopyPool &get() { static CopyPool *p = new CopyPool(); return *p; }      // never destroyed: no thread joins at exit
    void copy(void *dst, const void *src, size_t n) {
        const size_t piece_min = 512u << 10;
        const size_t parts = std::min<size_t>(workers_.size() + 1, (n + piece_min - 1) / piece_min);
        if (parts <= 1) { std::memcpy(dst, src, n); return; }
        const size_t step = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
        Batch batch;
        size_t mine_len = 0;
        {
            std::lock_guard<std::mutex> lock(mu_);
            for (size_t at = step; at < n; at += step) {
                tasks_.push_back({static_cast<char *>(dst) + at, static_cast<const char *>(src) + at, std::min(step, n - at), &batch});
                batch.pending++;
            }
            mine_len = std::min(step, n);
        }
        cv_.notify_all();
        std::memcpy(dst, src, mine_len);
        std::unique_lock<std::mutex> lock(mu_);
        batch.done.wait(lock, [&] { return batch.pending == 0; });
    }
 private:
    struct Batch { size_t pending = 0; std::condition_variable done; };
    struct Task { char *dst; const char *src; size_t n; Batch *batch; };
    CopyPool() {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned n = hw <= 2 ? 0u : std::min(hw - 1, 11u);
        if (const char *e = getenv("MMG_COPY_THREADS")) n = (unsigned)std::max(0, atoi(e) - 1);       // development aid
        for (unsigned i = 0; i < n; i++) workers_.emplace_back([this] { run(); }).detach();
    }
    void run() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return !tasks_.empty(); });
                t = tasks_.front();
                tasks_.pop_front();
            }
            std::memcpy(t.dst, t.src, t.n);
            std::lock_guard<std::mutex> lock(mu_);
            if (--t.batch->pending == 0) t.batch->done.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Task> tasks_;
    std::vector<std::thread> workers_;
};

// Pageable host input -> device: chunks go through a ring of pinned staging buffers (pool copy, then an asynchronous
// H2D copy each), so the link runs at its own speed instead of the driver's single-threaded pageable path.
struct StagingRing {
    static constexpr int K = 4;
    static constexpr size_t CHUNK = 8u << 20;      // (4 MiB chunks in 1 MiB pieces kept 4 of 12 copy threads busy: 8 GB/s, profiles/r2_stage_sweep.txt)
    std::mutex mu;                       // one staged copy at a time per process
    uint8_t *buf[K] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done[K] = {nullptr, nullptr, nullptr, nullptr};
    bool ok = false, tried = false;
    static StagingRing &get() { static StagingRing *r = new StagingRing(); return *r; }
    bool init() {
        if (tried) return ok;
        tried = true;
        for (int i = 0; i < K; i++) {
            if (cudaHostAlloc((void **)&buf[i], CHUNK, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        }
        ok = true;
        return true;
    }
};

void copy_host_to_device(uint8_t *dst, const void *src, uint64_t nbytes, cudaStream_t stream) {
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    StagingRing &ring = StagingRing::get();
    // Pageable input: above this size the bytes go through the pinned ring (pool copies overlap the H2D copies, 41 GB/s
    // at 256 MiB); below it the driver's own staging is faster -- the ring's worker wake-ups cost more than they save
    // (8 MiB: 0.50 ms direct against 0.9-1.4 ms through the ring, profiles/r2_bench_search_gpu.txt vs r1).
    static const uint64_t stage_min = (getenv("MMG_STAGE_MIN_MIB") ? (uint64_t)atoi(getenv("MMG_STAGE_MIN_MIB")) : 32) << 20;
    if (!pinned && nbytes >= stage_min) {
        std::lock_guard<std::mutex> lock(ring.mu);
        if (ring.init()) {
            static const bool prof = getenv("MMG_PROFILE_COPY") != nullptr;      // development aid: phase times on stderr
            auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
            double t_wait = 0, t_copy = 0, t_issue = 0, t0 = now();
            uint64_t at = 0;
            for (int i = 0; at < nbytes; i++, at += StagingRing::CHUNK) {
                const int k = i % StagingRing::K;
                const size_t n = (size_t)std::min<uint64_t>(StagingRing::CHUNK, nbytes - at);
                double t = now();
                if (i >= StagingRing::K) CU(cudaEventSynchronize(ring.done[k]));
                t_wait += now() - t; t = now();
                CopyPool::get().copy(ring.buf[k], static_cast<const uint8_t *>(src) + at, n);
                t_copy += now() - t; t = now();
                CU(cudaMemcpyAsync(dst + at, ring.buf[k], n, cudaMemcpyHostToDevice, stream));
                CU(cudaEventRecord(ring.done[k], stream));
                t_issue += now() - t;
            }
            const double t1 = now();
            CU(cudaStreamSynchronize(stream));       // the ring is free again when the lock is released
            if (prof) fprintf(stderr, "[mmg] staged H2D of %.1f MiB: ring wait %.3f ms, pool copies %.3f ms, issue %.3f ms, drain %.3f ms, total %.3f ms\n",
                              nbytes / 1048576.0, 1e3 * t_wait, 1e3 * t_copy, 1e3 * t_issue, 1e3 * (now() - t1), 1e3 * (now() - t0));
            return;
        }
    }
    CU(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, stream));
}

// stream-ordered scratch that is released when the scan ends
struct Arena {
    cudaStream_t stream;
    std::vector<void *> ptrs;
    explicit Arena(cudaStream_t s) : stream(s) {}
    ~Arena() { for (void *p : ptrs) cudaFreeAsync(p, stream); }
    template <class T> T *get(size_t n) {
        void *p = nullptr;
        CU(cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), stream));
        ptrs.push_back(p);
        return static_cast<T *>(p);
    }
};

}  // namespace

struct ScanRequest {
    const mmg_program *prog;
    const uint8_t *d_bytes;   // device
    uint64_t S;               // valid bytes
    uint64_t B;               // block size
    uint64_t nblocks;
    uint32_t npads;
    bool big_endian;
    uint64_t base_offset;
    uint32_t report_shift;
    bool chain = false;       // slice of a longer chain (mmg_chain_*): B = bytes whose windows this slice owns
};

// what a launched-but-not-yet-finished tiled scan needs to be completed (or re-run after an overflow)
struct TiledState {
    MmgGeom G{};
    MmgScratch X{};
    int lag_bytes = 0, grid = 0;
    int chain = 0;                // slice of a longer chain: CHAIN_MAPS while the entry phase is unknown, CHAIN_FULL afterwards
    bool sparse = false;          // resolved inside the filter kernel (fused sparse resolve): no per-sub-tile match bookkeeping exists
    uint64_t total_warps = 0, per_warp = 0, cap = 0;
    uint64_t generation = 0;      // workspace generation of this scan's last enqueue
    Workspace *ws = nullptr;
};

struct mmg_results {
    uint64_t count = 0;
    uint64_t *d_off = nullptr;    // one allocation: offsets, then values
    uint32_t *d_val = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t free_stream = nullptr;   // non-null: release the lists in this stream's order (a gather still reads them there)
    DeviceInfo *dev = nullptr;            // process-wide device state the scan was enqueued on
    mmg_scan_stats stats{};
    // ---- in-flight state (mmg_*_async): completed by finish_scan()
    bool pending = false;
    int error = MMG_OK;
    bool tiled = false, from_host = false;
    ScanRequest rq{};
    TiledState t;
    Arena *arena = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // start (host input only), scan start, after filter, end
    uint64_t *status_host = nullptr;                            // pinned slot receiving X.status
    uint8_t *chain_map = nullptr;                               // pinned [2][jp]: map of a chain slice (mmg_chain_*)
    Workspace *own_ws = nullptr;                                // chain slices keep their events between the two halves: a
                                                                // workspace of their own, not the lane's (several slices are in flight at once)
    int chain_stage = 0;                                        // 0 not a chain slice, 1 maps enqueued, 2 maps read, 3 resolve enqueued
    uint32_t launches = 0;
};

enum { CHAIN_MAPS = 1, CHAIN_FULL = 2 };

namespace {

// small pools shared by all threads: CUDA events and pinned status slots are expensive to create per scan
std::mutex g_pool_mutex;
std::vector<cudaEvent_t> g_event_pool;
std::vector<uint64_t *> g_slot_pool;
std::vector<uint8_t *> g_map_pool;

cudaEvent_t take_event() {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    }
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    return e;
}

uint64_t *take_slot() {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_slot_pool.empty()) {
        uint64_t *block = nullptr;
        CU(cudaHostAlloc((void **)&block, 64 * 8 * sizeof(uint64_t), cudaHostAllocMapped | cudaHostAllocPortable));   // the resolve kernel writes the status words here
        for (int i = 0; i < 64; i++) g_slot_pool.push_back(block + 8 * i);
    }
    uint64_t *s = g_slot_pool.back();
    g_slot_pool.pop_back();
    return s;
}

uint8_t *take_map() {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_map_pool.empty()) {
        uint8_t *block = nullptr;
        CU(cudaHostAlloc((void **)&block, 16 * 2 * MMG_MAXL, cudaHostAllocMapped | cudaHostAllocPortable));   // k_slicemap writes the slice's map here
        for (int i = 0; i < 16; i++) g_map_pool.push_back(block + 2 * MMG_MAXL * i);
    }
    uint8_t *m = g_map_pool.back();
    g_map_pool.pop_back();
    return m;
}

void release_inflight(mmg_results *r) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (auto &e : r->ev) { if (e) g_event_pool.push_back(e); e = nullptr; }
    if (r->status_host) { g_slot_pool.push_back(r->status_host); r->status_host = nullptr; }
    if (r->chain_map) { g_map_pool.push_back(r->chain_map); r->chain_map = nullptr; }
    if (r->own_ws) {
        Workspace *w = r->own_ws;
        if (w->zero) cudaFreeAsync(w->zero, r->stream);
        if (w->scratch) cudaFreeAsync(w->scratch, r->stream);
        if (w->ev) cudaFreeAsync(w->ev, r->stream);
        delete w;
        r->own_ws = nullptr;
    }
    delete r->arena;
    r->arena = nullptr;
}

void alloc_results(mmg_results *res, uint64_t cap) {
    const size_t off_bytes = (cap * sizeof(uint64_t) + 255) & ~(size_t)255;
    uint8_t *buf = nullptr;
    CU(cudaMallocAsync((void **)&buf, off_bytes + cap * sizeof(uint32_t), res->stream));
    res->d_off = reinterpret_cast<uint64_t *>(buf);
    res->d_val = reinterpret_cast<uint32_t *>(buf + off_bytes);
}

void free_results(mmg_results *res) {
    if (res->d_off) cudaFreeAsync(res->d_off, res->stream);
    res->d_off = nullptr;
    res->d_val = nullptr;
}

// Device copies of the arrays of long keywords (MmgLongProgram), per program and device; made at the first scan (pattern
// compilation itself needs no GPU), released by mmg_program_free.
struct LongArrays { int device; MmgCheck *chk; int32_t *tab_key, *tab_val; };
std::mutex g_long_mutex;
std::multimap<const mmg_program *, LongArrays> g_long_arrays;

MmgLongProgram long_program_on_device(const mmg_program *prog) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_long_mutex);
    LongArrays a{dev, nullptr, nullptr, nullptr};
    bool found = false;
    auto range = g_long_arrays.equal_range(prog);
    for (auto it = range.first; it != range.second; ++it)
        if (it->second.device == dev) { a = it->second; found = true; }
    if (!found) {
        const size_t nc = prog->long_chk.size(), nt = prog->long_tab_key.size();
        CU(cudaMalloc((void **)&a.chk, std::max<size_t>(nc, 1) * sizeof(MmgCheck)));
        CU(cudaMalloc((void **)&a.tab_key, std::max<size_t>(nt, 1) * sizeof(int32_t)));
        CU(cudaMalloc((void **)&a.tab_val, std::max<size_t>(nt, 1) * sizeof(int32_t)));
        if (nc) CU(cudaMemcpy(a.chk, prog->long_chk.data(), nc * sizeof(MmgCheck), cudaMemcpyHostToDevice));
        if (nt) {
            CU(cudaMemcpy(a.tab_key, prog->long_tab_key.data(), nt * sizeof(int32_t), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(a.tab_val, prog->long_tab_val.data(), nt * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
        g_long_arrays.insert({prog, a});
    }
    const MmgProgram &d = prog->dev;
    MmgLongProgram P{};
    P.W = d.W; P.L = d.L; P.modular = d.modular; P.ncheck = (int32_t)prog->long_chk.size(); P.ntab = (int32_t)prog->long_tab_key.size();
    P.tab_default = d.tab_default; P.match_jump = d.match_jump; P.first_lit = d.first_lit; P.opp_idx = d.opp_idx;
    P.chk = a.chk; P.tab_key = a.tab_key; P.tab_val = a.tab_val;
    return P;
}

void release_long_program(const mmg_program *prog) {
    std::lock_guard<std::mutex> lock(g_long_mutex);
    auto range = g_long_arrays.equal_range(prog);
    for (auto it = range.first; it != range.second; ++it) {
        int cur = 0;
        const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != it->second.device && cudaSetDevice(it->second.device) == cudaSuccess;
        cudaFree(it->second.chk); cudaFree(it->second.tab_key); cudaFree(it->second.tab_val);
        if (switched) cudaSetDevice(cur);
    }
    g_long_arrays.erase(range.first, range.second);
    cudaGetLastError();
}

void run_generic(const ScanRequest &rq, cudaStream_t stream, Arena &arena, mmg_results *res, uint32_t &launches) {
    const MmgProgram &P = rq.prog->dev;
    const bool is_long = rq.prog->is_long;
    MmgLongProgram LP{};
    if (is_long) LP = long_program_on_device(rq.prog);
    auto walk = [&](const MmgGeom &g, uint32_t *cnt, const uint64_t *bs, uint64_t *oo, uint32_t *ov) {
        return is_long ? mmg_launch_generic_walk_long(LP, g, cnt, bs, oo, ov, stream) : mmg_launch_generic_walk(P, g, cnt, bs, oo, ov, stream);
    };
    MmgGeom G{};
    G.data = rq.d_bytes; G.S = rq.S; G.B = rq.B; G.base_offset = rq.base_offset;
    G.nblocks = (uint32_t)rq.nblocks; G.ov = (uint32_t)(P.L - 1) * P.W; G.npads = rq.npads;
    G.big_endian = rq.big_endian; G.report_shift = rq.report_shift;
    G.complete = g_complete.load() ? 1u : 0u;
    const uint64_t chains = rq.nblocks * rq.npads;
    if (chains > 0x7FFFFFFFull) throw ScanError{fail(MMG_ERR_ARG, "too many blocks for the generic path")};
    const uint32_t n = (uint32_t)chains;
    uint32_t *counts = arena.get<uint32_t>(n);
    uint64_t *bases = arena.get<uint64_t>(n);
    uint64_t *bsum = arena.get<uint64_t>((n + 1023) / 1024 + 1);
    uint64_t *total_d = arena.get<uint64_t>(1);
    CU(walk(G, counts, nullptr, nullptr, nullptr));
    CU(mmg_launch_scan(counts, n, bsum, bases, total_d, stream));
    launches += 4;
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, total_d, sizeof(total), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    res->count = total;
    if (total == 0) return;
    alloc_results(res, total);
    if (rq.npads == 1) {
        CU(walk(G, counts, bases, res->d_off, res->d_val));
        launches += 1;
    } else {
        uint64_t *tmp_off = arena.get<uint64_t>(total);
        uint32_t *tmp_val = arena.get<uint32_t>(total);
        CU(walk(G, counts, bases, tmp_off, tmp_val));
        CU(mmg_launch_generic_merge(G.nblocks, counts, bases, tmp_off, tmp_val, res->d_off, res->d_val, stream));
        launches += 2;
    }
}

template <class T> void grow(T *&ptr, size_t &have, size_t need, cudaStream_t stream, bool zero) {
    if (need <= have) return;
    if (ptr) CU(cudaFreeAsync(ptr, stream));
    ptr = nullptr; have = 0;
    const size_t want = need + need / 4;
    CU(cudaMallocAsync((void **)&ptr, want * sizeof(T), stream));
    have = want;
    if (zero) CU(cudaMemsetAsync(ptr, 0, want * sizeof(T), stream));
}

// enqueues one attempt of the tiled pipeline on the thread's workspace: filter, resolve -- nothing else
void enqueue_tiled(mmg_results *res) {
    const MmgProgram &P = res->rq.prog->dev;
    TiledState &t = res->t;
    Workspace &ws = *t.ws;
    cudaStream_t stream = res->stream;
    if (ws.stream != stream) {                    // the caller switched streams: order the workspace's users by hand
        if (ws.stream) CU(cudaStreamSynchronize(ws.stream));
        ws.stream = stream;
    }
    const MmgGeom &G = t.G;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t at = off; off = (off + bytes + 255) & ~(size_t)255; return at; };
    // zero state
    const size_t o_status = carve(4 * sizeof(uint64_t));
    const size_t o_ticket = carve(8 * sizeof(uint32_t));
    const size_t o_lookback = carve((size_t)G.nseg * sizeof(uint64_t));
    const size_t zero_need = off;
    off = 0;
    const size_t o_ext = carve((size_t)G.nsub * sizeof(uint2));
    const size_t o_brec = carve((size_t)G.nblocks * 32 * sizeof(uint32_t));
    const size_t o_bcount = carve((size_t)G.nblocks * sizeof(uint32_t));
    const size_t o_mcount = carve((size_t)G.nsub * sizeof(uint32_t));
    const size_t o_mbase = carve((size_t)G.nsub * sizeof(uint64_t));
    const size_t jp = (size_t)((P.Jmax + 15) / 16 * 16);
    const bool segmented = G.segs_per_block > 1 || t.chain;
    const size_t o_segmap = carve(segmented ? (size_t)G.nseg * 2 * jp : 0);
    const size_t o_segphase = carve(segmented ? (size_t)G.nseg * 2 : 0);
    const size_t o_rangemap = carve(t.chain || mmg_resolve_two_level(G) ? (size_t)128 * 2 * jp : 0);
    const size_t scratch_need = off;
    const bool zero_grew = zero_need > ws.zero_bytes;
    grow(ws.zero, ws.zero_bytes, zero_need, stream, true);
    if (ws.dirty && !zero_grew) CU(cudaMemsetAsync(ws.zero, 0, ws.zero_bytes, stream));
    ws.dirty = true;                              // until both kernels are enqueued
    grow(ws.scratch, ws.scratch_bytes, scratch_need, stream, false);
    grow(ws.ev, ws.ev_entries, (size_t)(t.per_warp * t.total_warps), stream, false);

    MmgScratch &X = t.X;
    X.status = reinterpret_cast<uint64_t *>(ws.zero + o_status);
    X.ticket = reinterpret_cast<uint32_t *>(ws.zero + o_ticket);
    X.lookback = reinterpret_cast<uint64_t *>(ws.zero + o_lookback);
    X.ext = reinterpret_cast<uint2 *>(ws.scratch + o_ext);
    X.brec = reinterpret_cast<uint32_t *>(ws.scratch + o_brec);
    X.bcount = reinterpret_cast<uint32_t *>(ws.scratch + o_bcount);
    X.ev_total = (uint64_t)t.per_warp * t.total_warps;
    X.out_off = res->d_off; X.out_val = res->d_val; X.capacity = t.cap;
    X.mcount = reinterpret_cast<uint32_t *>(ws.scratch + o_mcount);
    X.mbase = reinterpret_cast<uint64_t *>(ws.scratch + o_mbase);
    X.segmap = ws.scratch + o_segmap;
    X.segphase = ws.scratch + o_segphase;
    X.rangemap = ws.scratch + o_rangemap;
    X.slicemap_host = res->chain_map;
    X.ev = ws.ev;
    X.ev_per_warp = (uint32_t)t.per_warp;
    X.host_status = res->status_host;             // pinned + mapped: same address on the device (UVA)

    // A pattern whose previous scan left only a few events per engine block is resolved inside the filter kernel (one
    // launch); should this scan turn out denser the kernel says so and finish_tiled() runs the resolve kernel over the
    // same events.
    {
        const mmg_program *prog = res->rq.prog;
        const uint64_t hint_bytes = prog->last_bytes.load(), hint_events = prog->last_events.load();
        static const bool no_sparse = getenv("MMG_NO_SPARSE_RESOLVE") != nullptr;
        // Not beside a result gather: the fused kernel is launched cooperatively (its grid barrier needs every CTA
        // resident), so it cannot share the SMs with an NCCL kernel the way the plain persistent grid does -- scans and
        // gathers of a multi-GPU pipeline would take turns instead of overlapping.
        t.sparse = !t.chain && !t.G.complete && !no_sparse && g_path_override != 4 && g_live_comms.load() == 0 && hint_bytes != 0 && mmg_sparse_resolve_supported(t.G) &&
                   (double)hint_events / (double)hint_bytes * (double)t.G.B <= 128.0;       // <= 128 events per block expected
    }
    X.fuse = t.sparse ? 1u : 0u;
    if (res->launches == 0) CU(cudaEventRecord(res->ev[1], stream));      // ev[1] -> ev[2] brackets the filter kernel alone
    CU(mmg_launch_filter(P, t.G, X, t.lag_bytes, t.grid, stream));
    if (res->launches == 0) CU(cudaEventRecord(res->ev[2], stream));
    if (t.chain) {
        // slice of a longer chain: the slice's map first; the resolve half follows here only when the entry phase is
        // already known (a re-run after mmg_chain_finish), otherwise mmg_chain_finish enqueues it.  Until the resolve
        // kernel has run nobody restores the zero state: the workspace stays marked dirty.
        CU(mmg_launch_chain_maps(P, t.G, X, stream));
        res->launches += 4;
        if (t.chain == CHAIN_FULL) {
            CU(mmg_launch_chain_resolve(P, t.G, X, res->d_off, res->d_val, t.cap, stream));
            res->launches += 2;
            ws.dirty = false;
        }
        t.generation = ++ws.generation;
        return;
    }
    if (!t.sparse) CU(mmg_launch_resolve(P, t.G, X, res->d_off, res->d_val, t.cap, stream));
    ws.dirty = false;
    res->launches += t.sparse ? 1 : (t.G.segs_per_block > 1 ? (mmg_resolve_two_level(t.G) ? 5 : 4) : 2);
    t.generation = ++ws.generation;
}

void launch_tiled(mmg_results *res, DeviceInfo &dev, Workspace &ws) {
    const ScanRequest &rq = res->rq;
    const MmgProgram &P = rq.prog->dev;
    TiledState &t = res->t;
    t.ws = &ws;
    const int W = P.W;
    t.lag_bytes = (P.ncheck > 0 && P.nkeys >= 0) ? P.chk[0].lag * W : 0;
    if (!mmg_filter_supported(W, t.lag_bytes) || g_path_override == 2) t.lag_bytes = 0;   // evaluate every window exactly

    MmgGeom &G = t.G;
    G.data = rq.d_bytes; G.S = rq.S; G.base_offset = rq.base_offset;
    G.nblocks = (uint32_t)rq.nblocks; G.ov = (uint32_t)(P.L - 1) * W; G.npads = rq.npads;
    G.big_endian = rq.big_endian; G.report_shift = rq.report_shift;
    // input stream marked evict_first in the L2 (MMG_L2_HINT=0 switches it off): +2 % at 512 MiB, more on larger inputs
    G.complete = g_complete.load() ? 1u : 0u;
    static const uint32_t l2_hint = getenv("MMG_L2_HINT") ? (uint32_t)atoi(getenv("MMG_L2_HINT")) : 1u;
    G.l2_hint = l2_hint;
    // (a chain slice that is continued by another one owns exactly B bytes of windows; the bytes behind are overlap)
    const bool continued = rq.chain && rq.S > rq.B;
    if (rq.nblocks == 1 && !continued) G.B = ((std::max<uint64_t>(std::max(rq.S, rq.B), 1) + MMG_SUBTILE - 1) / MMG_SUBTILE) * MMG_SUBTILE;
    else G.B = rq.B;
    if (rq.chain) { t.chain = CHAIN_MAPS; G.chain = 1; }
    const uint64_t spb = G.B / MMG_SUBTILE;
    const uint64_t last_off = (rq.nblocks - 1) * G.B;
    const uint64_t nsub64 = (rq.nblocks - 1) * spb + (rq.S - last_off + MMG_SUBTILE - 1) / MMG_SUBTILE;
    if (spb > 0xFFFFFFFFull || nsub64 > 0xFFFFFFF0ull) throw ScanError{fail(MMG_ERR_ARG, "input too large for one scan")};
    G.spb = (uint32_t)spb;
    G.nsub = (uint32_t)nsub64;
    // resolve geometry: segments of at most 128 sub-tiles
    G.segs_per_block = (uint32_t)((spb + 127) / 128);
    static const bool noseg = getenv("MMG_NOSEG") != nullptr;      // debug: one resolve CTA per engine block
    if (noseg) G.segs_per_block = 1;
    {
        const uint64_t last_subs = nsub64 - (rq.nblocks - 1) * spb;            // sub-tiles of the last (possibly short) block
        const uint64_t nseg64 = noseg ? rq.nblocks : (rq.nblocks - 1) * G.segs_per_block + (last_subs + 127) / 128;
        if (nseg64 > 0x7FFFFFFFull) throw ScanError{fail(MMG_ERR_ARG, "input too large for one scan")};
        G.nseg = (uint32_t)nseg64;
    }

    int occ = 1;
    CU(mmg_filter_occupancy(W, t.lag_bytes, rq.big_endian, P.nkeys, &occ));
    t.grid = dev.sms * std::max(occ, 1);
    t.total_warps = (uint64_t)t.grid * 8;
    // sub-tiles per chunk: small enough that dynamic scheduling balances the tail (at least 4 chunks per warp;
    // MMG_CHUNKS_PER_WARP overrides), large enough that the u16 queue entries of k_filter can address the chunk (<= 8 sub-tiles + halo)
    static const uint64_t per_warp_min = getenv("MMG_CHUNKS_PER_WARP") ? (uint64_t)atoi(getenv("MMG_CHUNKS_PER_WARP")) : 4;
    uint32_t cs = 8;
    while (cs > 1 && (spb % cs != 0 || nsub64 / cs < per_warp_min * t.total_warps)) cs >>= 1;
    G.chunk_subs = cs;
    G.nchunks = (uint32_t)((nsub64 + cs - 1) / cs);
    // Small inputs (fewer than four chunks per resident warp even at one sub-tile per chunk): dynamic scheduling cannot
    // hide the tail any more -- 4096 chunks on 3552 warps take two full rounds with the second one 15 % busy.  Size the
    // grid so that every warp gets the same number of chunks: ceil(chunks / rounds) warps.
    if ((uint64_t)G.nchunks < per_warp_min * t.total_warps) {
        const uint64_t rounds = ((uint64_t)G.nchunks + t.total_warps - 1) / t.total_warps;
        const uint64_t warps = ((uint64_t)G.nchunks + rounds - 1) / std::max<uint64_t>(rounds, 1);
        t.grid = (int)std::max<uint64_t>(1, (warps + 7) / 8);
        t.total_warps = (uint64_t)t.grid * 8;
        G.static_chunks = 1;
    }

    // optimistic result capacity: what this pattern produced last time plus slack
    t.cap = std::max<uint64_t>(4096, rq.prog->last_count + rq.prog->last_count / 4 + 1024);
    alloc_results(res, t.cap);

    // event capacity: private, equally sized regions per filter warp; grown and re-run on overflow
    t.per_warp = std::max<uint64_t>(256, rq.S / 8 / t.total_warps);
    if (rq.prog->last_events_per_warp) t.per_warp = std::max<uint64_t>(256, rq.prog->last_events_per_warp * 2);
    if (t.lag_bytes == 0) t.per_warp = std::max<uint64_t>(t.per_warp, (rq.S / t.total_warps + MMG_SUBTILE) * 2);
    if (t.per_warp * t.total_warps > 0xFFFFFFF0ull) throw ScanError{fail(MMG_ERR_NOMEM, "event buffer would exceed 2^32 entries")};
    enqueue_tiled(res);
}

// completes a tiled scan: waits, re-runs with exact sizes when an optimistic buffer was too small
void finish_tiled(mmg_results *res) {
    const MmgProgram &P = res->rq.prog->dev;
    TiledState &t = res->t;
    cudaStream_t stream = res->stream;
    CU(cudaEventSynchronize(res->ev[3]));
    volatile uint64_t *status = res->status_host;
    for (int attempt = 0; status[0] > t.per_warp; attempt++) {
        if (attempt >= 3) throw ScanError{fail(MMG_ERR_NOMEM, "event buffer overflow persists")};
        t.per_warp = status[0] + status[0] / 4 + 64;
        if (t.per_warp * t.total_warps > 0xFFFFFFF0ull) throw ScanError{fail(MMG_ERR_NOMEM, "event buffer would exceed 2^32 entries")};
        {
            std::lock_guard<std::mutex> lock(res->dev->mu);
            enqueue_tiled(res);
            CU(cudaEventRecord(res->ev[3], stream));
        }
        CU(cudaStreamSynchronize(stream));
    }
    const uint64_t ev_max = status[0], ev_total = status[1];
    if (t.sparse && status[4] != 0) {
        // a block held more events than the sparse kernel stages: the general resolve kernel over the same event lists
        // (the zero state was restored, the lists are intact unless another scan used the workspace since)
        std::lock_guard<std::mutex> lock(res->dev->mu);
        t.sparse = false;
        res->stats.resolve_kind = 2;
        if (t.generation == t.ws->generation) {
            CU(mmg_launch_resolve(P, t.G, t.X, res->d_off, res->d_val, t.cap, stream));
            res->launches += t.G.segs_per_block > 1 ? (mmg_resolve_two_level(t.G) ? 4 : 3) : 1;
        } else {
            res->rq.prog->last_events = ev_total;       // so that the re-run picks the general kernel
            res->rq.prog->last_bytes = res->rq.S;
            enqueue_tiled(res);
        }
        CU(cudaEventRecord(res->ev[3], stream));
        CU(cudaStreamSynchronize(stream));
    }
    if (t.sparse) res->stats.resolve_kind = 1;
    res->rq.prog->last_events_per_warp = ev_max;
    res->rq.prog->last_events = ev_total;
    res->rq.prog->last_bytes = res->rq.S;
    res->rq.prog->last_count = status[2];
    res->stats.events = ev_total;
    res->count = status[2];
    if (res->count > t.cap) {
        // the optimistic buffer was too small: allocate exactly and emit again from the stored bases -- or, when
        // another scan has used the workspace since, run the whole scan again
        {
            std::lock_guard<std::mutex> lock(res->dev->mu);
            free_results(res);
            t.cap = res->count;
            alloc_results(res, t.cap);
            if (t.generation == t.ws->generation && !t.sparse) {
                CU(mmg_launch_emit(P, t.G, t.X, res->d_off, res->d_val, stream));
                res->launches += 1;
            } else {
                enqueue_tiled(res);
            }
            CU(cudaEventRecord(res->ev[3], stream));
        }
        CU(cudaStreamSynchronize(stream));
    }
}

// Blocks until the scan behind `r` is complete; returns its status code.
int finish_scan(mmg_results *r) {
    if (!r->pending) return r->error;
    r->pending = false;
    try {
        if (r->tiled) finish_tiled(r);
        else CU(cudaEventSynchronize(r->ev[3]));
        float ms = 0;
        r->stats.ms_h2d = 0.f;
        if (r->from_host) { CU(cudaEventElapsedTime(&ms, r->ev[0], r->ev[1])); r->stats.ms_h2d = ms; }
        if (cudaEventElapsedTime(&ms, r->ev[1], r->ev[2]) == cudaSuccess) r->stats.ms_filter = ms; else cudaGetLastError();
        CU(cudaEventElapsedTime(&ms, r->ev[1], r->ev[3])); r->stats.ms_total = ms;
        r->stats.launches = r->launches;
    } catch (const ScanError &e) {
        r->error = e.code;
        cudaGetLastError();
    }
    release_inflight(r);
    return r->error;
}

// Enqueues a scan on the calling thread's stream; *out is a pending results object.
int launch_scan(const mmg_program *prog, const void *bytes, uint64_t nbytes, int mem, uint64_t B, uint64_t nblocks,
                uint32_t npads, bool big_endian, uint64_t base_offset, uint32_t report_shift, mmg_results **out,
                bool chain = false) {
    *out = nullptr;
    mmg_results *res = new mmg_results();
    try {
        static const int ndev = [] { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; } return n; }();
        if (ndev == 0) throw ScanError{fail(MMG_ERR_CUDA, "no CUDA device: this library has no CPU fallback")};
        DeviceInfo &dev = device_info();
        res->dev = &dev;
        std::lock_guard<std::mutex> lock(dev.mu);             // enqueue order == workspace order (see DeviceInfo)
        // Consecutive scans alternate between two stream lanes so that the (tiny) resolve kernel of a sparse scan runs
        // beside the next filter kernel.  A scan with millions of events has a resolve kernel that wants the whole GPU:
        // behind a persistent filter grid it would only start when that grid drains, so such scans stay on one lane
        // (cfg5, 2 GiB: 0.93 ms per scan serial against 1.28 ms alternating).
        const bool heavy = prog->last_events.load() >= (1u << 20);
        Lane &lane = dev.lanes[heavy ? 0u : (dev.next_lane++ & 1u)];
        cudaStream_t stream = lane.stream;
        res->stream = stream;
        res->stats.bytes_scanned = nbytes;
        if (nbytes == 0 || nblocks == 0) { *out = res; return MMG_OK; }
        CU(cudaEventRecord(lane.fork, caller_stream(dev)));           // everything the caller enqueued so far (the input!) comes first
        CU(cudaStreamWaitEvent(stream, lane.fork, 0));
        for (auto &e : res->ev) e = take_event();
        res->status_host = take_slot();
        res->arena = new Arena(stream);
        res->from_host = mem == MMG_MEM_HOST;
        const uint8_t *d_bytes = static_cast<const uint8_t *>(bytes);
        if (mem == MMG_MEM_HOST) {
            CU(cudaEventRecord(res->ev[0], stream));
            uint8_t *buf = res->arena->get<uint8_t>(nbytes + 16);
            copy_host_to_device(buf, bytes, nbytes, stream);
            d_bytes = buf;
        } else if ((reinterpret_cast<uintptr_t>(bytes) & 15u) != 0) {
            throw ScanError{fail(MMG_ERR_ARG, "device pointer must be 16-byte aligned")};
        }
        if (mem == MMG_MEM_HOST) CU(cudaEventRecord(res->ev[1], stream));      // end of the copy (re-recorded before the filter)
        res->rq = ScanRequest{prog, d_bytes, nbytes, B, nblocks, npads, big_endian, base_offset, report_shift, chain};
        const bool regular = nblocks == 1 || (B % MMG_SUBTILE) == 0;
        if (chain && prog->is_long) throw ScanError{fail(MMG_ERR_ARG, "chain slices are not available for keywords longer than 128 elements")};
        res->tiled = regular && (g_path_override != 1 || chain) && !prog->is_long;      // long keywords: per-chain kernels only
        res->stats.fast_path = res->tiled;
        if (chain) res->chain_map = take_map();
        if (res->tiled) {
            if (chain) res->own_ws = new Workspace();
            launch_tiled(res, dev, chain ? *res->own_ws : lane.ws);
        } else {
            CU(cudaEventRecord(res->ev[1], stream));
            run_generic(res->rq, stream, *res->arena, res, res->launches);
            CU(cudaEventRecord(res->ev[2], stream));
        }
        CU(cudaEventRecord(res->ev[3], stream));
        if (chain) res->chain_stage = 1;         // completed by mmg_chain_map / mmg_chain_finish, not by finish_scan
        else res->pending = true;
        *out = res;
        return MMG_OK;
    } catch (const ScanError &e) {
        cudaGetLastError();
        release_inflight(res);
        res->chain_stage = 0;
        mmg_results_free(res);
        return e.code;
    }
}

// chain slice, first half: waits for the slice's map; grows the event buffer and re-runs on overflow
void finish_chain_maps(mmg_results *res) {
    TiledState &t = res->t;
    cudaStream_t stream = res->stream;
    CU(cudaEventSynchronize(res->ev[3]));
    volatile uint64_t *status = res->status_host;
    for (int attempt = 0; status[0] > t.per_warp; attempt++) {
        if (attempt >= 3) throw ScanError{fail(MMG_ERR_NOMEM, "event buffer overflow persists")};
        t.per_warp = status[0] + status[0] / 4 + 64;
        if (t.per_warp * t.total_warps > 0xFFFFFFF0ull) throw ScanError{fail(MMG_ERR_NOMEM, "event buffer would exceed 2^32 entries")};
        {
            std::lock_guard<std::mutex> lock(res->dev->mu);
            enqueue_tiled(res);
            CU(cudaEventRecord(res->ev[3], stream));
        }
        CU(cudaStreamSynchronize(stream));
    }
}

int run_scan(const mmg_program *prog, const void *bytes, uint64_t nbytes, int mem, uint64_t B, uint64_t nblocks,
             uint32_t npads, bool big_endian, uint64_t base_offset, uint32_t report_shift, bool async, mmg_results **out) {
    int rc = launch_scan(prog, bytes, nbytes, mem, B, nblocks, npads, big_endian, base_offset, report_shift, out);
    if (rc != MMG_OK || async) return rc;
    rc = finish_scan(*out);
    if (rc != MMG_OK) { mmg_results_free(*out); *out = nullptr; }
    return rc;
}

}  // namespace

extern "C" {

const char *mmg_last_error(void) { return g_err.c_str(); }

int mmg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mmg_set_stream(void *cuda_stream, int use_it) {
    g_user_stream = static_cast<cudaStream_t>(cuda_stream);
    g_use_user_stream = use_it != 0;
    return MMG_OK;
}

// testing knob: 0 auto, 1 force the per-chain generic kernels, 2 force exact evaluation of every window,
// 4 never fuse the resolve into the filter kernel
int mmg_set_path_override(int mode) {
    int old = g_path_override;
    g_path_override = mode;
    return old;
}

int mmg_set_complete_matches(int on) { return g_complete.exchange(on ? 1 : 0); }

int mmg_program_create_keyword(const uint32_t *keyword, int keyword_len, uint32_t wildcard, const uint32_t *char_seq,
                               int char_seq_len, int elem_bits, mmg_program **out) {
    if (!out) return fail(MMG_ERR_ARG, "out is null");
    std::string err;
    int rc = mmg_compile_pattern(keyword, keyword_len, wildcard, char_seq, char_seq_len, false, elem_bits, out, err);
    if (rc != MMG_OK) g_err = err;
    return rc;
}

int mmg_program_create_values(const int16_t *values, int n, int elem_bits, mmg_program **out) {
    if (!out) return fail(MMG_ERR_ARG, "out is null");
    *out = nullptr;
    if (n <= 0 || !values) return fail(MMG_ERR_EMPTY, "empty value list");
    // static_cast<CharType>(short): /root/reference/src/core/monkey_moore.cpp:33-35
    std::vector<uint32_t> kw(n);
    for (int i = 0; i < n; i++) kw[i] = static_cast<uint32_t>(static_cast<int32_t>(values[i]));
    std::string err;
    int rc = mmg_compile_pattern(kw.data(), n, 0, nullptr, 0, true, elem_bits, out, err);
    if (rc != MMG_OK) g_err = err;
    return rc;
}

void mmg_program_free(mmg_program *p) {
    if (p && p->is_long) release_long_program(p);
    delete p;
}
int mmg_program_keyword_len(const mmg_program *p) { return p->dev.L; }
int mmg_program_max_jump(const mmg_program *p) { return p ? p->dev.Jmax : 0; }
int mmg_program_mode(const mmg_program *p) { return p->mode; }

int mmg_program_table_size(const mmg_program *p) {
    if (p->mode == 2) return 0;
    if (p->char_seq.empty()) return 2;
    return static_cast<int>(p->seq_index.size());
}

void mmg_program_table(const mmg_program *p, uint32_t v0, uint32_t v1, uint32_t *keys, uint32_t *values) {
    const uint32_t mask = p->elem_bits == 8 ? 0xFFu : 0xFFFFu;
    if (p->mode == 2) return;
    if (p->char_seq.empty()) {
        // /root/reference/src/core/monkey_moore.cpp:380-385 (simple), :476-512 (wildcard / mixed case)
        const uint32_t ref = p->mode == 1 ? p->normalized[p->dev.first_lit] : p->keyword[0];
        const uint32_t dist = v0 - ref;
        uint32_t upper = 'A' + dist, lower = 'a' + dist;
        if (p->mode == 1 && p->has_case_change && p->dev.opp_idx >= 0) {
            const uint32_t odist = v1 - p->keyword[p->dev.opp_idx];
            if (p->mostly_lowercase) upper = 'A' + odist; else lower = 'a' + odist;
        }
        keys[0] = 'A'; values[0] = upper & mask;
        keys[1] = 'a'; values[1] = lower & mask;
        return;
    }
    // :387-391, :515-520 -- every character of the sequence, shifted by the first literal's distance
    const uint32_t refc = p->mode == 1 ? p->keyword[p->dev.first_lit] : p->keyword[0];
    const uint32_t dist = v0 - static_cast<uint32_t>(p->value_of(refc));
    int n = 0;
    for (const auto &kv : p->seq_index) {   // std::map: ascending key order, like the reference's result map
        keys[n] = kv.first;
        values[n] = (static_cast<uint32_t>(kv.second) + dist) & mask;
        n++;
    }
}

uint64_t mmg_num_blocks(uint64_t file_size, uint32_t block_size) {
    if (block_size == 0) return 0;
    return (file_size + block_size - 1) / block_size;
}

int mmg_search(const mmg_program *p, const void *data, uint64_t data_len, int mem, mmg_results **out) {
    if (!p || !out || (!data && data_len)) return fail(MMG_ERR_ARG, "null argument");
    const uint32_t W = p->dev.W;
    const uint64_t nbytes = data_len * W;
    return run_scan(p, data, nbytes, mem, nbytes, nbytes ? 1 : 0, 1, false, 0, W == 2 ? 1 : 0, false, out);
}

int mmg_chain_begin(const mmg_program *p, const void *data, uint64_t owned_len, uint64_t avail_len, int mem,
                    uint64_t first_element, mmg_results **out) {
    if (!p || !out || (!data && avail_len)) return fail(MMG_ERR_ARG, "null argument");
    if (avail_len < owned_len) return fail(MMG_ERR_ARG, "avail_len smaller than owned_len");
    const uint32_t W = p->dev.W;
    const uint64_t owned = owned_len * W, avail = std::min(avail_len, owned_len + (uint64_t)(p->dev.L - 1)) * W;
    if (avail > owned && (owned == 0 || owned % MMG_SUBTILE != 0))
        return fail(MMG_ERR_ARG, "a slice that is continued must own a positive multiple of 4096 bytes");
    if (owned == 0) return fail(MMG_ERR_ARG, "empty slice");
    if (avail_len > owned_len && avail_len < owned_len + (uint64_t)(p->dev.L - 1))
        return fail(MMG_ERR_ARG, "a slice that is continued needs keyword_len - 1 elements of the next slice behind it");
    return launch_scan(p, data, avail, mem, owned, 1, 1, false, first_element * W, W == 2 ? 1 : 0, out, true);
}

int mmg_chain_map(mmg_results *r, uint8_t *map, int capacity, int *n) {
    if (!r || !map || r->chain_stage == 0) return fail(MMG_ERR_ARG, "not a chain slice");
    if (r->error != MMG_OK) return r->error;
    const int J = r->rq.prog->dev.Jmax;
    if (capacity < J) return fail(MMG_ERR_ARG, "map buffer too small");
    if (r->chain_stage == 1) {
        try {
            finish_chain_maps(r);
        } catch (const ScanError &e) {
            r->error = e.code;
            cudaGetLastError();
            return e.code;
        }
        r->chain_stage = 2;
    }
    if (r->chain_stage != 2) return fail(MMG_ERR_ARG, "the slice's map is gone once mmg_chain_finish has run");
    for (int e = 0; e < J; e++) map[e] = r->chain_map[e];       // class 0: search() scans one alignment
    if (n) *n = J;
    return MMG_OK;
}

uint32_t mmg_chain_entry(const uint8_t *maps, int stride, int nslices) {
    uint32_t ph = 0;                                             // the chain starts at element 0 of slice 0
    for (int k = 0; k < nslices; k++) ph = maps[(size_t)k * stride + ph];
    return ph;
}

int mmg_chain_finish(mmg_results *r, uint32_t entry_phase) {
    if (!r || r->chain_stage == 0) return fail(MMG_ERR_ARG, "not a chain slice");
    if (r->error != MMG_OK) return r->error;
    if (r->chain_stage == 3) return fail(MMG_ERR_ARG, "mmg_chain_finish called twice");
    if (entry_phase >= (uint32_t)r->rq.prog->dev.Jmax) return fail(MMG_ERR_ARG, "entry phase outside the pattern's jump range");
    try {
        if (r->chain_stage == 1) finish_chain_maps(r);
        const MmgProgram &P = r->rq.prog->dev;
        TiledState &t = r->t;
        std::lock_guard<std::mutex> lock(r->dev->mu);
        t.G.entry[0] = entry_phase;
        r->stats.chain_entry = entry_phase;
        t.chain = CHAIN_FULL;                                    // any later re-run does both halves
        if (t.generation == t.ws->generation) {
            CU(mmg_launch_chain_resolve(P, t.G, t.X, r->d_off, r->d_val, t.cap, r->stream));
            r->launches += 2;
            t.ws->dirty = false;
        } else {
            enqueue_tiled(r);                                    // another scan used the workspace in between
        }
        CU(cudaEventRecord(r->ev[3], r->stream));
    } catch (const ScanError &e) {
        r->error = e.code;
        cudaGetLastError();
        return e.code;
    }
    r->chain_stage = 3;
    r->pending = true;
    return MMG_OK;
}

static int engine_scan(const mmg_program *p, const void *bytes, uint64_t nbytes, int mem, uint64_t file_size,
                       uint32_t block_size, uint64_t first_block, uint64_t num_blocks, int big_endian, bool async,
                       mmg_results **out) {
    if (!p || !out || (!bytes && nbytes)) return fail(MMG_ERR_ARG, "null argument");
    if (block_size == 0) return fail(MMG_ERR_ARG, "block_size must be positive");
    const uint64_t all_blocks = mmg_num_blocks(file_size, block_size);
    if (first_block > all_blocks) return fail(MMG_ERR_ARG, "first_block beyond the file");
    if (num_blocks == 0 || first_block + num_blocks > all_blocks) num_blocks = all_blocks - first_block;
    const uint32_t W = p->dev.W;
    const uint64_t ov = (uint64_t)(p->dev.L - 1) * W;
    const uint64_t base = first_block * (uint64_t)block_size;
    uint64_t need = std::min<uint64_t>(file_size - std::min(base, file_size), num_blocks * (uint64_t)block_size + ov);
    if (nbytes < need) return fail(MMG_ERR_ARG, "slice shorter than the blocks it must cover");
    return run_scan(p, bytes, need, mem, block_size, num_blocks, W, big_endian != 0 && W == 2, base, 0, async, out);
}

int mmg_engine_scan(const mmg_program *p, const void *bytes, uint64_t nbytes, int mem, uint64_t file_size,
                    uint32_t block_size, uint64_t first_block, uint64_t num_blocks, int big_endian,
                    mmg_results **out) {
    return engine_scan(p, bytes, nbytes, mem, file_size, block_size, first_block, num_blocks, big_endian, false, out);
}

int mmg_engine_scan_async(const mmg_program *p, const void *bytes, uint64_t nbytes, int mem, uint64_t file_size,
                          uint32_t block_size, uint64_t first_block, uint64_t num_blocks, int big_endian,
                          mmg_results **out) {
    return engine_scan(p, bytes, nbytes, mem, file_size, block_size, first_block, num_blocks, big_endian, true, out);
}

int mmg_results_wait(mmg_results *r) {
    if (!r) return fail(MMG_ERR_ARG, "null results");
    return finish_scan(r);
}

// every accessor completes a pending scan first
static mmg_results *done(const mmg_results *r) {
    mmg_results *m = const_cast<mmg_results *>(r);
    if (m && m->pending) finish_scan(m);
    return m;
}

uint64_t mmg_results_count(const mmg_results *r) { return r ? done(r)->count : 0; }

int mmg_results_copy(const mmg_results *r, uint64_t first, uint64_t n, uint64_t *offsets, uint32_t *values) {
    if (r) done(r);
    if (r && r->error != MMG_OK) return r->error;
    if (!r || first + n > r->count) return fail(MMG_ERR_ARG, "range outside the result list");
    if (n == 0) return MMG_OK;
    if (offsets) {
        if (cudaMemcpy(offsets, r->d_off + first, n * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(MMG_ERR_CUDA, "copying offsets failed");
    }
    if (values) {
        std::vector<uint32_t> packed(n);
        if (cudaMemcpy(packed.data(), r->d_val + first, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(MMG_ERR_CUDA, "copying values failed");
        for (uint64_t i = 0; i < n; i++) { values[2 * i] = packed[i] & 0xFFFFu; values[2 * i + 1] = packed[i] >> 16; }
    }
    return MMG_OK;
}

// Distinct inferred tables: see unique.cu.
int mmg_results_unique(const mmg_program *p, const mmg_results *r, uint64_t *indices, uint64_t capacity, uint64_t *n_unique) {
    if (!p || !r || !n_unique) return fail(MMG_ERR_ARG, "null argument");
    done(r);
    if (r->error != MMG_OK) return r->error;
    *n_unique = 0;
    const uint64_t M = r->count;
    if (M == 0) return MMG_OK;
    // which of the two emitted element values the table depends on (mirrors mmg_program_table)
    uint32_t keymask = 0xFFFFu;
    if (p->mode == 2) keymask = 0;                                                   // value scan: no table at all
    else if (p->char_seq.empty() && p->mode == 1 && p->has_case_change && p->dev.opp_idx >= 0) keymask = 0xFFFFFFFFu;
    std::vector<uint64_t> firsts;
    if (keymask == 0) {
        firsts.push_back(0);
    } else {
        cudaStream_t stream = r->stream;
        if (keymask != 0xFFFFu && p->elem_bits != 8 && M >= 0xFFFFFFFFull)
            return fail(MMG_ERR_NOMEM, "match list too long for the hashed unique-table pass");
        try {
            Arena tmp(stream);                      // temporaries are released on every path, error paths included
            if (keymask == 0xFFFFu || p->elem_bits == 8) {
                uint64_t *d_first = tmp.get<uint64_t>(65536);
                CU(cudaMemsetAsync(d_first, 0xFF, 65536 * sizeof(uint64_t), stream));
                CU(mmg_launch_unique_direct(r->d_val, M, keymask, keymask != 0xFFFFu, d_first, stream));
                std::vector<uint64_t> table(65536);
                CU(cudaMemcpyAsync(table.data(), d_first, 65536 * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
                CU(cudaStreamSynchronize(stream));
                for (uint64_t v : table)
                    if (v != ~0ull) firsts.push_back(v);
            } else {
                uint64_t cap = 1024;
                while (cap < 2 * M && cap < (1ull << 28)) cap <<= 1;
                const uint64_t out_cap = std::min<uint64_t>(M, cap);
                uint64_t *d_slots = tmp.get<uint64_t>(cap), *d_out = tmp.get<uint64_t>(out_cap), *d_cnt = tmp.get<uint64_t>(2);
                CU(cudaMemsetAsync(d_slots, 0xFF, cap * sizeof(uint64_t), stream));
                CU(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(uint64_t), stream));
                CU(mmg_launch_unique_hash(r->d_val, M, d_slots, (uint32_t)(cap - 1), reinterpret_cast<unsigned int *>(d_cnt + 1), stream));
                CU(mmg_launch_unique_collect(d_slots, cap, true, d_out, d_cnt, out_cap, stream));
                uint64_t h_cnt[2] = {0, 0};
                CU(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, stream));
                CU(cudaStreamSynchronize(stream));
                if ((h_cnt[1] & 0xFFFFFFFFull) != 0)
                    throw ScanError{fail(MMG_ERR_NOMEM, "more distinct tables than the hashed unique-table pass can hold")};
                firsts.resize(std::min<uint64_t>(h_cnt[0], out_cap));
                if (!firsts.empty())
                    CU(cudaMemcpyAsync(firsts.data(), d_out, firsts.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
                CU(cudaStreamSynchronize(stream));
            }
        } catch (const ScanError &e) {
            cudaGetLastError();
            return e.code;
        }
        std::sort(firsts.begin(), firsts.end());
    }
    *n_unique = firsts.size();
    if (indices)
        for (uint64_t i = 0; i < firsts.size() && i < capacity; i++) indices[i] = firsts[i];
    return MMG_OK;
}

const uint64_t *mmg_results_device_offsets(const mmg_results *r) { return r ? done(r)->d_off : nullptr; }
const uint32_t *mmg_results_device_values(const mmg_results *r) { return r ? done(r)->d_val : nullptr; }

void mmg_results_free(mmg_results *r) {
    if (!r) return;
    if (r->pending) finish_scan(r);
    if (r->chain_stage == 1 || r->chain_stage == 2) {       // abandoned chain slice: its kernels still use the buffers
        if (r->ev[3]) cudaEventSynchronize(r->ev[3]);
        release_inflight(r);
    }
    // one allocation: offsets, then values.  A gather that still reads the lists on its own stream has redirected
    // the release to that stream (mmg_internal_results_free_on): the allocator then cannot recycle them early.
    if (r->d_off) cudaFreeAsync(r->d_off, r->free_stream ? r->free_stream : r->stream);
    delete r;
}

// internal doorways for comm.cu (not part of the public header)
struct mmg_results_view { uint64_t count; const uint64_t *d_off; const uint32_t *d_val; };
int mmg_internal_results_view(const mmg_results *r, mmg_results_view *out) {
    if (r) done(r);
    out->count = r ? r->count : 0;
    out->d_off = r ? r->d_off : nullptr;
    out->d_val = r ? r->d_val : nullptr;
    return MMG_OK;
}
void *mmg_internal_gather_stream(void) {
    try { return device_info().gather_stream; } catch (const ScanError &) { return nullptr; }
}
void mmg_internal_results_free_on(const mmg_results *r, void *stream) {
    if (r) const_cast<mmg_results *>(r)->free_stream = static_cast<cudaStream_t>(stream);
}
void mmg_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }
void mmg_internal_comm_alive(int delta) { g_live_comms += delta; }

void *mmg_host_alloc(uint64_t nbytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        fail(MMG_ERR_NOMEM, "cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}

void mmg_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int mmg_host_copy(void *dst, const void *src, uint64_t nbytes) {
    if ((!dst || !src) && nbytes) return fail(MMG_ERR_ARG, "null argument");
    CopyPool::get().copy(dst, src, (size_t)nbytes);
    return MMG_OK;
}

int mmg_synth_fill(void *device_ptr, uint64_t nbytes, uint64_t seed, uint64_t first_byte, uint32_t byte_mask) {
    if ((nbytes & 7u) || (first_byte & 7u) || (reinterpret_cast<uintptr_t>(device_ptr) & 7u))
        return fail(MMG_ERR_ARG, "mmg_synth_fill works on 8-byte aligned ranges");
    try {
        cudaStream_t stream = caller_stream(device_info());
        CU(mmg_launch_synth(static_cast<uint64_t *>(device_ptr), nbytes / 8, seed, first_byte / 8, byte_mask, stream));
        CU(cudaStreamSynchronize(stream));
    } catch (const ScanError &e) {
        return e.code;
    }
    return MMG_OK;
}

int mmg_results_stats(const mmg_results *r, mmg_scan_stats *out) {
    if (!r || !out) return fail(MMG_ERR_ARG, "null argument");
    done(r);
    *out = r->stats;
    return MMG_OK;
}

}  // extern "C"
