// Host-side plumbing shared by the translation units of libkm_b200.so: error reporting, grow-only arenas,
// recycled pinned blocks and the table handle.  (The library was one 2,100-line file; it is now cut along its
// subsystems: table_api.cu, io_api.cu, plan_api.cu, text_api.cu and one .cu per kernel family.)
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <tuple>
#include <chrono>

#include <cuda.h>        // driver-API TYPES only: the entry points are fetched at run time (vmm_load)
#include <unistd.h>

#include <cuda_runtime.h>
#include "../../include/km_b200.h"
#include "table.h"

using namespace km;

extern thread_local char g_err[512];
int fail(int code, const char* fmt, ...);
struct km_table;
int km_ensure_linked(km_table* t);
// Wait for `s`.  With several processes per box (one per GPU, six pool threads each) spinning waits oversubscribe the
// host cores in principle; with KM_BLOCKING_SYNC=1 the wait sleeps on a blocking-sync event instead (off by default: measured
// slower at 8 ranks per box).
cudaError_t km_wait_stream(cudaStream_t s, cudaEvent_t* blocking_ev);      // writes the neighbour masks if the table's content changed since (table_api.cu)

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(KM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// A grow-only arena: one cudaMalloc / cudaMallocHost reused across calls (allocation calls
// cost milliseconds, the whole panel runs in about one).
struct Arena {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    bool host = false;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (base) { host ? cudaFreeHost(base) : cudaFree(base); base = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + (1 << 20);
        cudaError_t e = host ? cudaMallocHost((void**)&base, want) : cudaMalloc((void**)&base, want);
        if (e != cudaSuccess) return fail(KM_E_CUDA, "arena alloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void reset() { used = 0; }
    template <class T> T* take(size_t n) {
        used = (used + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
    void release() { if (base) { host ? cudaFreeHost(base) : cudaFree(base); base = nullptr; cap = 0; } }
};

// Pinned host blocks recycled between results: a result's arrays are the direct target of the
// device-to-host copies (no pageable staging, no zero-filled vectors), and cudaMallocHost -- which
// costs milliseconds -- is paid once per size class instead of once per call.
struct PinPool {
    std::mutex m;
    std::vector<std::pair<char*, size_t>> idle;
    int acquire(size_t bytes, char** out, size_t* cap) {
        std::lock_guard<std::mutex> g(m);
        int best = -1;
        for (size_t i = 0; i < idle.size(); ++i)
            if (idle[i].second >= bytes && (best < 0 || idle[i].second < idle[(size_t)best].second)) best = (int)i;
        if (best >= 0) { *out = idle[(size_t)best].first; *cap = idle[(size_t)best].second; idle.erase(idle.begin() + best); return 0; }
        // nothing fits: allocate (with headroom, so that a slightly larger batch next time still fits);
        // the pool is trimmed when blocks come back (release)
        const size_t want = bytes + bytes / 4 + (1 << 16);
        cudaError_t e = cudaMallocHost((void**)out, want);
        if (e != cudaSuccess) return fail(KM_E_CUDA, "pinned alloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
        *cap = want;
        return 0;
    }
    void release(char* p, size_t cap) {
        if (!p) return;
        std::lock_guard<std::mutex> g(m);
        idle.emplace_back(p, cap);
        if (idle.size() > 48) {            // keep the largest blocks
            size_t small = 0;
            for (size_t i = 1; i < idle.size(); ++i) if (idle[i].second < idle[small].second) small = i;
            cudaFreeHost(idle[small].first);
            idle.erase(idle.begin() + (long)small);
        }
    }
    ~PinPool() { for (auto& b : idle) cudaFreeHost(b.first); }
};

template <class T> struct Span {
    T* p = nullptr;
    size_t n = 0;
    T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) const { return p[i]; }
};
struct PinBlock {
    std::shared_ptr<PinPool> pool;
    char* base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(const std::shared_ptr<PinPool>& from, size_t bytes) {
        drop();
        pool = from;
        used = 0;
        return pool->acquire(bytes, &base, &cap);
    }
    template <class T> Span<T> take(size_t n) {
        used = (used + 63) & ~(size_t)63;
        Span<T> s; s.p = reinterpret_cast<T*>(base + used); s.n = n;
        used += n * sizeof(T);
        return s;
    }
    void drop() { if (base && pool) pool->release(base, cap); base = nullptr; cap = 0; }
    ~PinBlock() { drop(); }
};

struct km_table {
    int device = 0, k = 31, canonical = 1;
    uint64_t n_buckets = 0, n_keys = 0;
    Bucket* buckets = nullptr;
    unsigned long long* d_counter = nullptr;   // [0] new keys, then a u32 "full" flag at +8
    cudaStream_t stream = nullptr, side = nullptr;   // side: the second shared-memory graph pass runs beside the first
    cudaEvent_t ev[8] = {}, fork = nullptr, join = nullptr;
    cudaStream_t side2 = nullptr, side3 = nullptr; cudaEvent_t join2 = nullptr, join3 = nullptr;     // the graph passes run four abreast
    Arena dev, pin;            // lookups / inserts
    Arena dev_find, pin_find;  // km_find_batch workspace, reused across calls
    // km_find_text runs a batch as several sub-batches in flight at once: each has its own workspace,
    // stream and events, kept across calls
    // (the host vectors of a lane's last plan are kept too: their capacity saves the next plan its allocations)
    struct PlanVecs { std::vector<int64_t> seq_off, node_off, hash_off, pack_off; std::vector<int32_t> chunk_target, chunk_start, extra; };
    struct Lane {
        Arena dev, pin; cudaStream_t stream = nullptr, side = nullptr, side2 = nullptr, side3 = nullptr;
        cudaEvent_t ev[8] = {}, fork = nullptr, join = nullptr, join2 = nullptr, join3 = nullptr; PlanVecs vecs;
        // the enqueue of a sub-batch as ONE driver call: when two consecutive calls give this lane a sub-batch of the very
        // same layout (a fixed panel of targets against sample after sample is km's workflow), the sequence copy +
        // ~16 kernels is captured into a CUDA graph and replayed from then on (text_api.cu)
        cudaGraphExec_t gexec = nullptr;
        std::string gkey, last_key;
        bool vecs_pristine = false; int vecs_k = 0, vecs_extra0 = 0, vecs_maxcap = 1;     // what `vecs` describes (plan_return_vecs)
        cudaEvent_t wait_ev = nullptr;        // blocking-sync event: a waiting pool thread sleeps instead of spinning (km_wait_stream)
    };
    std::vector<std::unique_ptr<Lane>> lanes;
    std::shared_ptr<PinPool> pool = std::make_shared<PinPool>();   // result buffers (outlive the table if a result does)
    int sm_count = 148;
    // cohort mode: this table is shard `my_shard` of `n_shards`; peer[r] = rank r's buckets mapped through CUDA IPC
    int lines = 0;             // 1: family-line layout (table.h), n_buckets counts 128-byte lines
    size_t unit() const { return lines ? sizeof(Line) : sizeof(Bucket); }
    int n_shards = 1, my_shard = 0;
    const Bucket* peer[KM_MAX_SHARDS] = {};
    bool attached = false;
    int route = 0;             // km_table_set_routing: inserts go to the owner shard (peer atomics over NVLink)
    bool linked = false;       // the neighbour masks (table.h) are current; any change of content clears this
    // a shard is allocated through the virtual-memory API so that peers can map it with its own 2 MiB
    // pages (a legacy cudaIpc mapping gets small pages: random probes of a 32 GB peer shard then run
    // ~70x slower, all TLB misses -- measured, profiles/README.md)
    bool vmm = false;
    CUmemGenericAllocationHandle vmm_handle = 0, peer_handle[KM_MAX_SHARDS] = {};
    size_t vmm_size = 0;
    TableView view() const {
        TableView v{};
        v.buckets = buckets; v.n_buckets = n_buckets; v.kmask = (1ull << (2 * k)) - 1ull; v.k = k; v.canonical = canonical;
        v.n_shards = n_shards; v.my_shard = my_shard; v.lines = lines; v.route = route;
        v.linked = (linked && n_shards == 1 && !lines) ? 1 : 0;
        for (int r = 0; r < KM_MAX_SHARDS; ++r) v.shard[r] = peer[r];
        v.shard[my_shard] = buckets;
        return v;
    }
};

static inline size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int grid_for(const km_table* t, uint64_t n, int block, int per_sm) {
    uint64_t want = (n + block - 1) / block;
    uint64_t cap = (uint64_t)t->sm_count * per_sm;
    if (want < 1) want = 1;
    return (int)std::min(want, cap);
}

// A byte stream of sequences on its way into the table (km_table_count_text / _reads / _file): two pinned staging
// buffers and their device twins, so that the host fills one while the GPU copies and counts the other.
struct CountStream {
    km_table* t = nullptr;
    size_t cap = 0;
    bool want_qual = false;
    char* pin_seq[2] = {nullptr, nullptr}; char* pin_q[2] = {nullptr, nullptr};
    char* dev_seq[2] = {nullptr, nullptr}; char* dev_q[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool busy[2] = {false, false};
    int slot = 0;
    uint64_t bytes_in = 0;
    int open(km_table* table, size_t cap_bytes, bool with_qual);
    char* seq() const { return pin_seq[slot]; }
    char* qual() const { return pin_q[slot]; }
    int submit(size_t n_bytes, int min_qual);      // asynchronous; afterwards seq()/qual() are the OTHER buffer, free to fill
    int close();                                   // waits for the GPU, updates the table's key count, frees everything
    ~CountStream();
};
