// C ABI of libkm_b200.so (see include/km_b200.h).  Host orchestration only: every piece of
// arithmetic on the find_mutation path runs in the kernels of kernels.cuh.  There is no CPU
// fallback -- without a CUDA device every entry point fails with KM_E_NOGPU.
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

#include "../../include/km_b200.h"
#include "kernels.cuh"
#include "format.h"

using namespace km;

static_assert(sizeof(km_row) == sizeof(Row), "km_row must mirror km::Row");
static_assert(sizeof(Bucket) == 32, "bucket must be one 32-byte sector");

static thread_local char g_err[512] = "";
// KM_TRACE: km_find_text's timeline (struct Trace below) is reachable from the plan functions through this hook
static thread_local void (*g_trace_mark)(void*, const char*, int) = nullptr;
static thread_local void* g_trace_obj = nullptr;
static thread_local int g_trace_sub = -1;
static inline void trace_here(const char* what) { if (g_trace_mark) g_trace_mark(g_trace_obj, what, g_trace_sub); }

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(KM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* km_last_error(void) { return g_err; }
extern "C" const char* km_version(void) { return "km_b200 0.1 (sm_100a)"; }
extern "C" int km_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

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
    Arena dev, pin;            // lookups / inserts
    Arena dev_find, pin_find;  // km_find_batch workspace, reused across calls
    // km_find_text runs a batch as several sub-batches in flight at once: each has its own workspace,
    // stream and events, kept across calls
    // (the host vectors of a lane's last plan are kept too: their capacity saves the next plan its allocations)
    struct PlanVecs { std::vector<int64_t> seq_off, node_off, hash_off, pack_off; std::vector<int32_t> chunk_target, chunk_start, extra; };
    struct Lane { Arena dev, pin; cudaStream_t stream = nullptr, side = nullptr; cudaEvent_t ev[8] = {}, fork = nullptr, join = nullptr; PlanVecs vecs; };
    std::vector<std::unique_ptr<Lane>> lanes;
    std::shared_ptr<PinPool> pool = std::make_shared<PinPool>();   // result buffers (outlive the table if a result does)
    int sm_count = 148;
    // cohort mode: this table is shard `my_shard` of `n_shards`; peer[r] = rank r's buckets mapped through CUDA IPC
    int lines = 0;             // 1: family-line layout (table.h), n_buckets counts 128-byte lines
    size_t unit() const { return lines ? sizeof(Line) : sizeof(Bucket); }
    int n_shards = 1, my_shard = 0;
    const Bucket* peer[KM_MAX_SHARDS] = {};
    bool attached = false;
    // a shard is allocated through the virtual-memory API so that peers can map it with its own 2 MiB
    // pages (a legacy cudaIpc mapping gets small pages: random probes of a 32 GB peer shard then run
    // ~70x slower, all TLB misses -- measured, profiles/README.md)
    bool vmm = false;
    CUmemGenericAllocationHandle vmm_handle = 0, peer_handle[KM_MAX_SHARDS] = {};
    size_t vmm_size = 0;
    TableView view() const {
        TableView v;
        v.buckets = buckets; v.n_buckets = n_buckets; v.kmask = (1ull << (2 * k)) - 1ull; v.k = k; v.canonical = canonical;
        v.n_shards = n_shards; v.my_shard = my_shard; v.lines = lines;
        for (int r = 0; r < KM_MAX_SHARDS; ++r) v.shard[r] = peer[r];
        v.shard[my_shard] = buckets;
        return v;
    }
};

static size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }

static bool default_lines() { const char* e = getenv("KM_TABLE_LINES"); return e && *e && *e != '0'; }
// units (sector buckets or family lines) for `capacity_keys` keys: buckets hold 2 records at load <= 0.5;
// lines hold 8 slots, every key takes two of them, load 0.625
static uint64_t units_for(uint64_t capacity_keys, int lines) {
    return std::max<uint64_t>(64, lines ? (capacity_keys * 2 + 4) / 5 : capacity_keys);
}
static void clear_units(km_table* t, void* mem, uint64_t n);

static int grid_for(const km_table* t, uint64_t n, int block, int per_sm) {
    uint64_t want = (n + block - 1) / block;
    uint64_t cap = (uint64_t)t->sm_count * per_sm;
    if (want < 1) want = 1;
    return (int)std::min(want, cap);
}

static void clear_units(km_table* t, void* mem, uint64_t n) {
    if (t->lines) km_table_clear_lines_kernel<<<t->sm_count * 8, 256, 0, t->stream>>>((Line*)mem, n);
    else km_table_clear_kernel<<<t->sm_count * 8, 256, 0, t->stream>>>((Bucket*)mem, n);
}

extern "C" int km_table_create_layout(int device, int k, int canonical, uint64_t capacity_keys, int lines, km_table** out);
extern "C" int km_table_create(int device, int k, int canonical, uint64_t capacity_keys, km_table** out) {
    return km_table_create_layout(device, k, canonical, capacity_keys, default_lines() ? 1 : 0, out);
}

extern "C" int km_table_create_layout(int device, int k, int canonical, uint64_t capacity_keys, int lines, km_table** out) {
    if (!out || k < 1 || k > 31) return fail(KM_E_ARG, "km_table_create: k must be in 1..31 (got %d)", k);
    int ndev = km_device_count();
    if (ndev <= 0) return fail(KM_E_NOGPU, "no CUDA device visible: km_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(KM_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    CU(cudaSetDevice(device));
    km_table* t = new km_table();
    t->device = device; t->k = k; t->canonical = canonical ? 1 : 0;
    t->lines = lines ? 1 : 0;
    if (t->lines && k < 2) { delete t; return fail(KM_E_ARG, "the family-line layout needs k >= 2"); }
    t->n_buckets = units_for(capacity_keys, t->lines);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    t->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaMalloc((void**)&t->buckets, t->n_buckets * t->unit());
    if (e != cudaSuccess) {
        const unsigned long long want = (unsigned long long)(t->n_buckets * t->unit());
        delete t;
        return fail(KM_E_CUDA, "cudaMalloc of %llu table bytes failed: %s", want, cudaGetErrorString(e));
    }
    CU(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking));
    for (auto& ev : t->ev) CU(cudaEventCreate(&ev));
    CU(cudaEventCreateWithFlags(&t->fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&t->join, cudaEventDisableTiming));
    CU(cudaMalloc((void**)&t->d_counter, 16));
    t->pin.host = true;
    t->pin_find.host = true;
    CU(cudaFuncSetAttribute(km_graph_kernel<KM_TINY_NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)class_layout(KM_TINY_NODES).stride));
    CU(cudaFuncSetAttribute(km_graph_kernel<KM_SMALL_NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)class_layout(KM_SMALL_NODES).stride));
    clear_units(t, t->buckets, t->n_buckets);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(t->stream));
    *out = t;
    return 0;
}

// ---- cohort mode: one shard per GPU, peers mapped over NVLink ---------------------------------------
// Driver entry points of the virtual-memory API, resolved at run time so that the library still loads
// on a machine without libcuda (the CPU-only test container).
struct Vmm {
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*export_fd)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*import_fd)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
    bool ok = false;
};
static Vmm g_vmm;
static int vmm_load() {
    if (g_vmm.ok) return 0;
    auto get = [](const char* name, void** fn) -> bool {
        cudaDriverEntryPointQueryResult q;
        return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    if (!get("cuMemCreate", (void**)&g_vmm.create) || !get("cuMemAddressReserve", (void**)&g_vmm.reserve) ||
        !get("cuMemMap", (void**)&g_vmm.map) || !get("cuMemSetAccess", (void**)&g_vmm.set_access) ||
        !get("cuMemExportToShareableHandle", (void**)&g_vmm.export_fd) ||
        !get("cuMemImportFromShareableHandle", (void**)&g_vmm.import_fd) ||
        !get("cuMemGetAllocationGranularity", (void**)&g_vmm.granularity) || !get("cuMemUnmap", (void**)&g_vmm.unmap) ||
        !get("cuMemRelease", (void**)&g_vmm.release) || !get("cuMemAddressFree", (void**)&g_vmm.addr_free))
        return fail(KM_E_CUDA, "CUDA virtual-memory API is not available from this driver");
    g_vmm.ok = true;
    return 0;
}
#define DRV(call)                                                                                       \
    do {                                                                                                \
        CUresult r_ = (call);                                                                           \
        if (r_ != CUDA_SUCCESS) return fail(KM_E_CUDA, "%s failed: CUresult %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); \
    } while (0)

static CUmemAllocationProp shard_prop(int device) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

// map `handle` (size bytes) into this process for device `device`
static int vmm_map(CUmemGenericAllocationHandle handle, size_t size, size_t gran, int device, void** out) {
    CUdeviceptr ptr = 0;
    DRV(g_vmm.reserve(&ptr, size, gran, 0, 0));
    DRV(g_vmm.map(ptr, size, 0, handle, 0));
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DRV(g_vmm.set_access(ptr, size, &acc, 1));
    *out = (void*)ptr;
    return 0;
}

extern "C" int km_table_create_shard(int device, int k, int canonical, uint64_t capacity_keys_per_shard, int rank, int n_shards,
                                     km_table** out) {
    if (n_shards < 1 || n_shards > KM_MAX_SHARDS || rank < 0 || rank >= n_shards)
        return fail(KM_E_ARG, "km_table_create_shard: rank %d of %d shards (at most %d)", rank, n_shards, KM_MAX_SHARDS);
    if (!out || k < 1 || k > 31) return fail(KM_E_ARG, "km_table_create_shard: k must be in 1..31 (got %d)", k);
    if (km_device_count() <= 0) return fail(KM_E_NOGPU, "no CUDA device visible: km_b200 has no CPU fallback");
    // the ordinary constructor with a token allocation, then the bucket array is replaced by a
    // shareable one of the real size
    // shards always use sector buckets: the experimental family-line layout (KM_TABLE_LINES) is single-GPU only
    if (int rc = km_table_create_layout(device, k, canonical, 64, 0, out)) return rc;
    km_table* t = *out;
    t->n_shards = n_shards; t->my_shard = rank;
    if (int rc = vmm_load()) { km_table_close(t); *out = nullptr; return rc; }
    CUmemAllocationProp prop = shard_prop(device);
    size_t gran = 0;
    DRV(g_vmm.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const uint64_t n_buckets = units_for(capacity_keys_per_shard, t->lines);
    const size_t size = align_up_sz(n_buckets * t->unit(), gran);
    CUmemGenericAllocationHandle h = 0;
    CUresult cr = g_vmm.create(&h, size, &prop, 0);
    if (cr != CUDA_SUCCESS) { km_table_close(t); *out = nullptr; return fail(KM_E_CUDA, "cuMemCreate of %zu shard bytes failed: CUresult %d", size, (int)cr); }
    void* ptr = nullptr;
    if (int rc = vmm_map(h, size, gran, device, &ptr)) { km_table_close(t); *out = nullptr; return rc; }
    cudaFree(t->buckets);
    t->buckets = (Bucket*)ptr; t->n_buckets = n_buckets;
    t->vmm = true; t->vmm_handle = h; t->vmm_size = size;
    clear_units(t, t->buckets, t->n_buckets);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(t->stream));
    return 0;
}

// a POSIX file descriptor for this shard's memory: send it to the peers (SCM_RIGHTS), they attach it
extern "C" int km_table_shard_export_fd(km_table* t, int* fd) {
    if (!t || !fd || !t->vmm) return fail(KM_E_ARG, "km_table_shard_export_fd: not a shard");
    CU(cudaSetDevice(t->device));
    int out = -1;
    DRV(g_vmm.export_fd(&out, t->vmm_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *fd = out;
    return 0;
}

extern "C" int km_table_shard_attach_fd(km_table* t, int rank, int fd) {
    if (!t || !t->vmm || rank < 0 || rank >= t->n_shards || rank == t->my_shard || fd < 0)
        return fail(KM_E_ARG, "km_table_shard_attach_fd: bad argument");
    CU(cudaSetDevice(t->device));
    CUmemGenericAllocationHandle h = 0;
    DRV(g_vmm.import_fd(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUmemAllocationProp prop = shard_prop(t->device);
    size_t gran = 0;
    DRV(g_vmm.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    void* ptr = nullptr;
    if (int rc = vmm_map(h, t->vmm_size, gran, t->device, &ptr)) return rc;     // every shard has the same size
    t->peer[rank] = (const Bucket*)ptr;
    t->peer_handle[rank] = h;
    t->attached = true;
    close(fd);
    return 0;
}

// owner shard of each k-mer (host arithmetic, no GPU): what routes a query in the all-to-all path
extern "C" int km_shard_owner(const uint64_t* kmers, uint64_t n, int k, int canonical, int n_shards, int32_t* owner) {
    if ((n && (!kmers || !owner)) || k < 1 || k > 31 || n_shards < 1) return fail(KM_E_ARG, "km_shard_owner: bad argument");
    const uint64_t mask = (1ull << (2 * k)) - 1ull;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t v = kmers[i] & mask;
        if (canonical) {
            uint64_t rc = ~v;
            rc = ((rc >> 2) & 0x3333333333333333ull) | ((rc & 0x3333333333333333ull) << 2);
            rc = ((rc >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((rc & 0x0F0F0F0F0F0F0F0Full) << 4);
            rc = __builtin_bswap64(rc) >> (64 - 2 * k);
            if (rc < v) v = rc;
        }
        uint64_t z = v + 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        owner[i] = n_shards > 1 ? (int32_t)(((unsigned __int128)z * (unsigned __int128)(uint64_t)n_shards) >> 64) : 0;
    }
    return 0;
}

extern "C" void km_table_close(km_table* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->vmm) {
        cudaDeviceSynchronize();
        for (int r = 0; r < KM_MAX_SHARDS; ++r)
            if (t->peer[r] && r != t->my_shard) {
                g_vmm.unmap((CUdeviceptr)t->peer[r], t->vmm_size); g_vmm.addr_free((CUdeviceptr)t->peer[r], t->vmm_size);
                g_vmm.release(t->peer_handle[r]);
            }
        if (t->buckets) { g_vmm.unmap((CUdeviceptr)t->buckets, t->vmm_size); g_vmm.addr_free((CUdeviceptr)t->buckets, t->vmm_size); }
        g_vmm.release(t->vmm_handle);
        t->buckets = nullptr;
    }
    if (t->buckets) cudaFree(t->buckets);
    if (t->d_counter) cudaFree(t->d_counter);
    t->dev.release();
    t->pin.release();
    t->dev_find.release();
    t->pin_find.release();
    for (auto& L : t->lanes) {
        L->dev.release(); L->pin.release();
        for (auto& e : L->ev) if (e) cudaEventDestroy(e);
        if (L->stream) cudaStreamDestroy(L->stream);
        if (L->side) cudaStreamDestroy(L->side);
        if (L->fork) cudaEventDestroy(L->fork);
        if (L->join) cudaEventDestroy(L->join);
    }
    for (auto& ev : t->ev) if (ev) cudaEventDestroy(ev);
    if (t->stream) cudaStreamDestroy(t->stream);
    if (t->side) cudaStreamDestroy(t->side);
    if (t->fork) cudaEventDestroy(t->fork);
    if (t->join) cudaEventDestroy(t->join);
    delete t;
}

extern "C" int km_table_get_info(km_table* t, km_table_info* info) {
    if (!t || !info) return fail(KM_E_ARG, "null argument");
    info->k = t->k; info->canonical = t->canonical; info->device = t->device; info->reserved = 0;
    info->n_keys = t->n_keys; info->n_buckets = t->n_buckets; info->bytes = t->n_buckets * t->unit();
    info->reserved = t->lines;
    return 0;
}

static int finish_insert(km_table* t, const char* what) {
    unsigned long long host[2] = {0, 0};
    CU(cudaMemcpyAsync(host, t->d_counter, 16, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    t->n_keys += host[0];
    if ((uint32_t)host[1]) return fail(KM_E_FULL, "%s: table full (%llu %s)", what, (unsigned long long)t->n_buckets, t->lines ? "lines" : "buckets");
    return 0;
}

extern "C" int km_table_insert(km_table* t, const uint64_t* keys, const uint32_t* counts, uint64_t n, int mode) {
    if (!t || (n && (!keys || !counts)) || mode < 0 || mode > 2) return fail(KM_E_ARG, "km_table_insert: bad argument");
    CU(cudaSetDevice(t->device));
    const uint64_t chunk = 1ull << 24;
    for (uint64_t done = 0; done < n; done += chunk) {
        const uint64_t m = std::min(chunk, n - done);
        if (int rc = t->dev.reserve(m * 12 + 512)) return rc;
        t->dev.reset();
        uint64_t* dk = t->dev.take<uint64_t>(m);
        uint32_t* dc = t->dev.take<uint32_t>(m);
        CU(cudaMemsetAsync(t->d_counter, 0, 16, t->stream));
        CU(cudaMemcpyAsync(dk, keys + done, m * 8, cudaMemcpyHostToDevice, t->stream));
        CU(cudaMemcpyAsync(dc, counts + done, m * 4, cudaMemcpyHostToDevice, t->stream));
        km_table_insert_kernel<<<grid_for(t, m, 256, 8), 256, 0, t->stream>>>(t->view(), dk, dc, m, mode, t->d_counter,
                                                                             reinterpret_cast<uint32_t*>(t->d_counter + 1));
        CU(cudaGetLastError());
        if (int rc = finish_insert(t, "km_table_insert")) return rc;
    }
    return 0;
}

extern "C" int km_table_build_synthetic(km_table* t, uint64_t seed, uint64_t n_keys) {
    if (!t) return fail(KM_E_ARG, "null table");
    CU(cudaSetDevice(t->device));
    CU(cudaMemsetAsync(t->d_counter, 0, 16, t->stream));
    km_table_synth_kernel<<<t->sm_count * 16, 256, 0, t->stream>>>(t->view(), seed, n_keys, t->d_counter,
                                                                   reinterpret_cast<uint32_t*>(t->d_counter + 1));
    CU(cudaGetLastError());
    return finish_insert(t, "km_table_build_synthetic");
}

extern "C" int km_table_count_reads(km_table* t, const char* reads, const int64_t* off, int64_t n_reads) {
    if (!t || !reads || !off || n_reads < 0) return fail(KM_E_ARG, "km_table_count_reads: bad argument");
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(t->device));
    const int64_t total = off[n_reads];
    if (int rc = t->dev.reserve((size_t)total + (size_t)(n_reads + 1) * 8 + 1024)) return rc;
    t->dev.reset();
    char* dr = t->dev.take<char>(total);
    int64_t* doff = t->dev.take<int64_t>(n_reads + 1);
    CU(cudaMemsetAsync(t->d_counter, 0, 16, t->stream));
    CU(cudaMemcpyAsync(dr, reads, total, cudaMemcpyHostToDevice, t->stream));
    CU(cudaMemcpyAsync(doff, off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, t->stream));
    km_count_reads_kernel<<<grid_for(t, total, 256, 8), 256, 0, t->stream>>>(t->view(), dr, doff, n_reads, total, t->d_counter,
                                                                            reinterpret_cast<uint32_t*>(t->d_counter + 1));
    CU(cudaGetLastError());
    return finish_insert(t, "km_table_count_reads");
}

extern "C" int km_table_drop_below(km_table* t, uint32_t min_count, uint64_t* n_left) {
    if (!t) return fail(KM_E_ARG, "null table");
    CU(cudaSetDevice(t->device));
    Bucket* fresh = nullptr;
    CU(cudaMalloc((void**)&fresh, t->n_buckets * t->unit()));
    clear_units(t, fresh, t->n_buckets);
    TableView dst = t->view();
    dst.buckets = fresh;
    dst.shard[t->my_shard] = fresh;
    CU(cudaMemsetAsync(t->d_counter, 0, 16, t->stream));
    km_table_filter_kernel<<<t->sm_count * 8, 256, 0, t->stream>>>(t->view(), dst, min_count, t->d_counter,
                                                                   reinterpret_cast<uint32_t*>(t->d_counter + 1));
    CU(cudaGetLastError());
    t->n_keys = 0;
    int rc = finish_insert(t, "km_table_drop_below");
    if (t->vmm) {
        // a shard keeps its (peer-mapped) memory: the filtered copy goes back in place
        CU(cudaMemcpyAsync(t->buckets, fresh, t->n_buckets * t->unit(), cudaMemcpyDeviceToDevice, t->stream));
        CU(cudaStreamSynchronize(t->stream));
        cudaFree(fresh);
    } else {
        cudaFree(t->buckets);
        t->buckets = fresh;
    }
    if (n_left) *n_left = t->n_keys;
    return rc;
}



// ---- counting straight from FASTA / FASTQ files (plain or .gz) ------------------------------------------
// zlib is looked up at run time (dlopen), like the driver's virtual-memory entry points: the library loads on
// a machine without it and only .gz input is refused there.
#include <dlfcn.h>
struct ZLib {
    void* (*open)(const char*, const char*) = nullptr;
    int (*read)(void*, void*, unsigned) = nullptr;
    int (*close)(void*) = nullptr;
    int (*buffer)(void*, unsigned) = nullptr;
    bool tried = false, ok = false;
};
static ZLib g_z;
static bool zlib_load() {
    if (g_z.tried) return g_z.ok;
    g_z.tried = true;
    void* h = dlopen("libz.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libz.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return false;
    g_z.open = (void* (*)(const char*, const char*))dlsym(h, "gzopen");
    g_z.read = (int (*)(void*, void*, unsigned))dlsym(h, "gzread");
    g_z.close = (int (*)(void*))dlsym(h, "gzclose");
    g_z.buffer = (int (*)(void*, unsigned))dlsym(h, "gzbuffer");
    g_z.ok = g_z.open && g_z.read && g_z.close;
    return g_z.ok;
}
struct LineReader {                 // lines of a plain or gzip file, without their line ends
    FILE* f = nullptr; void* gz = nullptr;
    std::vector<char> buf; size_t pos = 0, end = 0; bool eof = false;
    bool fill() {
        if (eof) return false;
        if (pos > 0) { memmove(buf.data(), buf.data() + pos, end - pos); end -= pos; pos = 0; }
        if (end == buf.size()) buf.resize(buf.size() * 2);
        const size_t room = buf.size() - end;
        long got = gz ? (long)g_z.read(gz, buf.data() + end, (unsigned)std::min<size_t>(room, 1u << 30)) : (long)fread(buf.data() + end, 1, room, f);
        if (got <= 0) { eof = true; return false; }
        end += (size_t)got;
        return true;
    }
    // next line into (*p, *n); false at end of file
    bool next(const char** p, size_t* n) {
        for (;;) {
            const char* nl = (const char*)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                *p = buf.data() + pos; *n = (size_t)(nl - *p);
                pos = (size_t)(nl - buf.data()) + 1;
                if (*n && (*p)[*n - 1] == '\r') --*n;
                return true;
            }
            if (!fill()) {
                if (pos < end) { *p = buf.data() + pos; *n = end - pos; pos = end; if (*n && (*p)[*n - 1] == '\r') --*n; return true; }
                return false;
            }
        }
    }
};

// `jellyfish count` input side: every sequence of the file goes through km_table_count_reads in batches of ~64 M
// bases.  FASTQ records are four lines; with min_qual_char > 0 a base whose quality character is below it
// counts as N (jellyfish count -Q).  FASTA sequences may span lines.
extern "C" int km_table_count_file(km_table* t, const char* path, int min_qual_char, uint64_t* n_reads_out, uint64_t* n_bases_out) {
    if (!t || !path) return fail(KM_E_ARG, "km_table_count_file: bad argument");
    LineReader R;
    const size_t plen = strlen(path);
    const bool gz = plen > 3 && strcmp(path + plen - 3, ".gz") == 0;
    if (gz) {
        if (!zlib_load()) return fail(KM_E_IO, "%s: libz.so.1 not found, decompress the file first", path);
        R.gz = g_z.open(path, "rb");
        if (!R.gz) return fail(KM_E_IO, "cannot open %s", path);
        if (g_z.buffer) g_z.buffer(R.gz, 1u << 20);
    } else {
        R.f = strcmp(path, "-") == 0 ? stdin : fopen(path, "rb");
        if (!R.f) return fail(KM_E_IO, "cannot open %s", path);
    }
    R.buf.resize((size_t)8 << 20);
    std::vector<char> blob;
    std::vector<int64_t> off(1, 0);
    blob.reserve((size_t)80 << 20);
    uint64_t n_reads = 0, n_bases = 0;
    int rc = 0;
    auto flush = [&]() -> int {
        if (off.size() <= 1) return 0;
        const int r = km_table_count_reads(t, blob.data(), off.data(), (int64_t)off.size() - 1);
        blob.clear(); off.assign(1, 0);
        return r;
    };
    auto end_read = [&]() -> int {
        if ((int64_t)blob.size() == off.back()) return 0;          // empty sequence
        n_reads += 1; n_bases += (uint64_t)((int64_t)blob.size() - off.back());
        off.push_back((int64_t)blob.size());
        return blob.size() >= ((size_t)64 << 20) ? flush() : 0;
    };
    const char* ln; size_t n;
    bool first = true, fastq = false;
    while (!rc && R.next(&ln, &n)) {
        if (first) {
            if (!n) continue;
            first = false;
            if (ln[0] == '@') fastq = true;
            else if (ln[0] != '>') { rc = fail(KM_E_IO, "%s: neither FASTA nor FASTQ", path); break; }
        }
        if (fastq) {
            if (!n) continue;                              // stray blank line between records
            // ln is the header; then sequence, '+', quality
            const char* sq; size_t sn;
            if (!R.next(&sq, &sn)) break;
            const size_t at = blob.size();
            blob.insert(blob.end(), sq, sq + sn);          // (sq stays valid until the next call of next())
            const char* pl; size_t pn; const char* ql; size_t qn;
            if (!R.next(&pl, &pn) || !R.next(&ql, &qn)) { rc = end_read(); break; }
            if (min_qual_char > 0 && qn == sn)
                for (size_t i = 0; i < sn; ++i) if ((unsigned char)ql[i] < (unsigned)min_qual_char) blob[at + i] = 'N';
            rc = end_read();
        } else {
            if (n && ln[0] == '>') rc = end_read();
            else blob.insert(blob.end(), ln, ln + n);
        }
    }
    if (!rc) rc = end_read();
    if (!rc) rc = flush();
    if (R.gz) g_z.close(R.gz); else if (R.f && R.f != stdin) fclose(R.f);
    if (n_reads_out) *n_reads_out = n_reads;
    if (n_bases_out) *n_bases_out = n_bases;
    return rc;
}

// ---- export: `jellyfish dump` and a binary/sorted writer ---------------------------------------------
extern "C" int km_table_export(km_table* t, uint64_t* keys, uint32_t* counts, uint64_t cap, uint64_t* n_out) {
    if (!t || !n_out || (cap && (!keys || !counts))) return fail(KM_E_ARG, "km_table_export: bad argument");
    CU(cudaSetDevice(t->device));
    const uint64_t m = std::min<uint64_t>(cap, t->n_keys);
    if (int rc = t->dev.reserve(m * 12 + 4096)) return rc;
    t->dev.reset();
    uint64_t* dk = t->dev.take<uint64_t>(m);
    uint32_t* dc = t->dev.take<uint32_t>(m);
    CU(cudaMemsetAsync(t->d_counter, 0, 16, t->stream));
    km_table_export_kernel<<<t->sm_count * 8, 256, 0, t->stream>>>(t->view(), dk, dc, m, t->d_counter);
    CU(cudaGetLastError());
    unsigned long long found = 0;
    CU(cudaMemcpyAsync(&found, t->d_counter, 8, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    *n_out = found;
    const uint64_t got = std::min<uint64_t>(found, m);
    if (got) {
        CU(cudaMemcpyAsync(keys, dk, got * 8, cudaMemcpyDeviceToHost, t->stream));
        CU(cudaMemcpyAsync(counts, dc, got * 4, cudaMemcpyDeviceToHost, t->stream));
        CU(cudaStreamSynchronize(t->stream));
    }
    return 0;
}

// Column i of the 32 x 62 binary matrix written into the header.  Any matrix works as long as the records are
// sorted by it; this one is a fixed pseudo-random one.
static uint32_t jf_matrix_column(int i) {
    uint64_t z = 0x6B6D5F62323030ull + (uint64_t)i;          // splitmix64 finaliser (host copy)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (uint32_t)(z >> 17);
}

// A Jellyfish 2.x `binary/sorted` file: "%09d" header length, JSON header (NUL-padded to 8 bytes), then
// ceil(key_len/8)-byte LE key + counter_len-byte LE count per record.  Verified on the five files bundled with
// km (tests/test_jf_writer.py): the records are sorted by pos = M * key over GF(2), masked to
// `size`, where bit b of the key selects column c-1-b of `matrix1`.  Ties (unobserved in the bundled files)
// are broken by key.  The writer emits its own matrix, so readers that binary-search by position stay consistent.
extern "C" int km_table_write_jf(km_table* t, const char* path, uint32_t counter_len) {
    if (!t || !path) return fail(KM_E_ARG, "km_table_write_jf: bad argument");
    if (counter_len == 0) counter_len = 4;
    if (counter_len > 8) return fail(KM_E_ARG, "km_table_write_jf: counter_len must be 1..8");
    std::vector<uint64_t> keys((size_t)t->n_keys);
    std::vector<uint32_t> counts((size_t)t->n_keys);
    uint64_t n = 0;
    if (int rc = km_table_export(t, keys.data(), counts.data(), t->n_keys, &n)) return rc;
    if (n != t->n_keys) return fail(KM_E_ARG, "km_table_write_jf: table holds %llu records, expected %llu", (unsigned long long)n, (unsigned long long)t->n_keys);
    const int kbits = 2 * t->k, kbytes = (kbits + 7) / 8;
    int lsize = 10;
    while (lsize < 32 && (1ull << lsize) < 2 * n) ++lsize;
    const uint64_t size = 1ull << lsize, mask = size - 1;
    // byte-sliced matrix-vector product: tab[j][v] = XOR of the columns selected by byte j of the key
    std::vector<uint32_t> col((size_t)kbits);
    for (int i = 0; i < kbits; ++i) col[(size_t)i] = jf_matrix_column(i);
    std::vector<uint32_t> tab((size_t)8 * 256, 0);
    for (int j = 0; j < 8; ++j)
        for (int v = 0; v < 256; ++v) {
            uint32_t x = 0;
            for (int b = 0; b < 8; ++b) { const int bit = 8 * j + b; if (((v >> b) & 1) && bit < kbits) x ^= col[(size_t)(kbits - 1 - bit)]; }
            tab[(size_t)j * 256 + (size_t)v] = x;
        }
    std::vector<uint64_t> order((size_t)n);      // pos << 32 | index would lose ties on the key: sort indices by (pos, key)
    std::vector<uint32_t> pos((size_t)n);
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t x = 0;
        for (int j = 0; j < 8; ++j) x ^= tab[(size_t)j * 256 + ((keys[i] >> (8 * j)) & 0xFF)];
        pos[i] = (uint32_t)(x & mask);
        order[i] = i;
    }
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return pos[a] != pos[b] ? pos[a] < pos[b] : keys[a] < keys[b]; });
    std::string js = "{\"alignment\":8,\"canonical\":";
    js += t->canonical ? "true" : "false";
    js += ",\"cmdline\":[\"km_b200\",\"count\"],\"counter_len\":" + std::to_string(counter_len) + ",\"format\":\"binary/sorted\",\"key_len\":" +
          std::to_string(kbits) + ",\"matrix1\":{\"c\":" + std::to_string(kbits) + ",\"columns\":[";
    for (int i = 0; i < kbits; ++i) { if (i) js += ','; js += std::to_string(col[(size_t)i]); }
    js += "],\"r\":32},\"max_reprobe\":126,\"reprobes\":[1";
    for (int i = 1; i <= 126; ++i) js += "," + std::to_string(i * (i + 1) / 2);
    js += "],\"size\":" + std::to_string(size) + ",\"val_len\":" + std::to_string(8 * counter_len > 12 ? 12 : 8 * counter_len) + "}";
    while ((9 + js.size()) % 8) js += '\0';
    FILE* f = fopen(path, "wb");
    if (!f) return fail(KM_E_IO, "cannot write %s", path);
    char digits[16];
    snprintf(digits, sizeof(digits), "%09zu", js.size());
    bool ok = fwrite(digits, 1, 9, f) == 9 && fwrite(js.data(), 1, js.size(), f) == js.size();
    const size_t rec = (size_t)kbytes + counter_len;
    std::vector<unsigned char> buf;
    buf.reserve(rec << 16);
    const uint64_t cmax = counter_len >= 4 ? 0xFFFFFFFFull : ((1ull << (8 * counter_len)) - 1);
    for (uint64_t i = 0; ok && i < n; ++i) {
        const uint64_t key = keys[order[i]];
        const uint64_t cnt = std::min<uint64_t>(counts[order[i]], cmax);       // a narrow counter saturates
        for (int b = 0; b < kbytes; ++b) buf.push_back((unsigned char)(key >> (8 * b)));
        for (uint32_t b = 0; b < counter_len; ++b) buf.push_back((unsigned char)(b < 8 ? cnt >> (8 * b) : 0));
        if (buf.size() >= (rec << 16)) { ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size(); buf.clear(); }
    }
    if (ok && !buf.empty()) ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    if (fclose(f) != 0) ok = false;
    if (!ok) return fail(KM_E_IO, "short write to %s", path);
    return 0;
}
// ---- .jf loader (binary/sorted) ---------------------------------------
static bool json_field(const std::string& js, const char* name, std::string* out) {
    std::string pat = std::string("\"") + name + "\"";
    size_t p = js.find(pat);
    if (p == std::string::npos) return false;
    p = js.find(':', p + pat.size());
    if (p == std::string::npos) return false;
    ++p;
    while (p < js.size() && isspace((unsigned char)js[p])) ++p;
    size_t e = p;
    if (js[p] == '"') { e = js.find('"', p + 1); if (e == std::string::npos) return false; *out = js.substr(p + 1, e - p - 1); return true; }
    while (e < js.size() && js[e] != ',' && js[e] != '}' && !isspace((unsigned char)js[e])) ++e;
    *out = js.substr(p, e - p);
    return true;
}

extern "C" int km_table_open_jf(const char* path, int device, km_table** out) {
    if (!path || !out) return fail(KM_E_ARG, "km_table_open_jf: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(KM_E_IO, "cannot open %s", path);
    char digits[10] = {0};
    if (fread(digits, 1, 9, f) != 9) { fclose(f); return fail(KM_E_IO, "%s: truncated header", path); }
    for (int i = 0; i < 9; ++i) if (!isdigit((unsigned char)digits[i])) { fclose(f); return fail(KM_E_IO, "%s: not a Jellyfish file (no header length)", path); }
    const long hlen = atol(digits);
    std::string js((size_t)hlen, '\0');
    if (fread(&js[0], 1, (size_t)hlen, f) != (size_t)hlen) { fclose(f); return fail(KM_E_IO, "%s: truncated header", path); }
    std::string fmt, canon, key_len, counter_len;
    if (!json_field(js, "format", &fmt) || !json_field(js, "canonical", &canon) || !json_field(js, "key_len", &key_len) ||
        !json_field(js, "counter_len", &counter_len)) { fclose(f); return fail(KM_E_IO, "%s: header lacks format/canonical/key_len/counter_len", path); }
    if (fmt != "binary/sorted") { fclose(f); return fail(KM_E_IO, "%s: unsupported format '%s' (only binary/sorted)", path, fmt.c_str()); }
    const int kbits = atoi(key_len.c_str()), cbytes = atoi(counter_len.c_str());
    if (kbits < 2 || kbits > 62 || (kbits & 1) || cbytes < 1 || cbytes > 8) { fclose(f); return fail(KM_E_IO, "%s: key_len %d / counter_len %d not supported", path, kbits, cbytes); }
    const int kbytes = (kbits + 7) / 8, rec = kbytes + cbytes;
    fseek(f, 0, SEEK_END);
    const long fsize = ftell(f);
    const long payload = fsize - 9 - hlen;
    if (payload < 0 || payload % rec) { fclose(f); return fail(KM_E_IO, "%s: payload of %ld bytes is not a multiple of %d", path, payload, rec); }
    const uint64_t n = (uint64_t)(payload / rec);
    std::vector<unsigned char> raw((size_t)payload);
    fseek(f, 9 + hlen, SEEK_SET);
    if (payload && fread(raw.data(), 1, (size_t)payload, f) != (size_t)payload) { fclose(f); return fail(KM_E_IO, "%s: short read", path); }
    fclose(f);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> counts(n);
    for (uint64_t i = 0; i < n; ++i) {
        const unsigned char* p = raw.data() + i * rec;
        uint64_t key = 0, cnt = 0;
        for (int b = 0; b < kbytes; ++b) key |= (uint64_t)p[b] << (8 * b);
        for (int b = 0; b < cbytes; ++b) cnt |= (uint64_t)p[kbytes + b] << (8 * b);
        keys[i] = key;
        counts[i] = cnt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cnt;
    }
    km_table* t = nullptr;
    if (int rc = km_table_create(device, kbits / 2, canon == "true", std::max<uint64_t>(n, 1024), &t)) return rc;
    if (int rc = km_table_insert(t, keys.data(), counts.data(), n, KM_INSERT_OVERWRITE)) { km_table_close(t); return rc; }
    *out = t;
    return 0;
}

// ---- lookups --------------------------------------------------------------------------------
extern "C" int km_query_batch_device(km_table* t, const uint64_t* kmers_dev, uint64_t n, uint32_t* counts_dev, void* stream) {
    if (!t || (n && (!kmers_dev || !counts_dev))) return fail(KM_E_ARG, "km_query_batch_device: bad argument");
    if (!n) return 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : t->stream;
    if (t->lines) km_query_lines_kernel<<<grid_for(t, n, 256, 8), 256, 0, s>>>(t->view(), kmers_dev, n, counts_dev);
    else km_query_kernel<<<grid_for(t, (n + KM_QUERY_ILP - 1) / KM_QUERY_ILP, 256, 8), 256, 0, s>>>(t->view(), kmers_dev, n, counts_dev);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int km_query_batch(km_table* t, const uint64_t* kmers, uint64_t n, uint32_t* counts) {
    if (!t || (n && (!kmers || !counts))) return fail(KM_E_ARG, "km_query_batch: bad argument");
    CU(cudaSetDevice(t->device));
    // chunks staged through pinned memory so copy-in, probe and copy-out of neighbouring
    // chunks overlap on the copy engines
    const uint64_t chunk = 1ull << 22;
    const uint64_t m_max = std::min(chunk, n);
    if (int rc = t->dev.reserve(2 * (m_max * 12 + 1024))) return rc;
    if (int rc = t->pin.reserve(2 * (m_max * 12 + 1024))) return rc;
    t->dev.reset(); t->pin.reset();
    uint64_t* dk[2]; uint32_t* dc[2]; uint64_t* hk[2]; uint32_t* hc[2];
    for (int b = 0; b < 2; ++b) {
        dk[b] = t->dev.take<uint64_t>(m_max); dc[b] = t->dev.take<uint32_t>(m_max);
        hk[b] = t->pin.take<uint64_t>(m_max); hc[b] = t->pin.take<uint32_t>(m_max);
    }
    cudaEvent_t done[2] = {t->ev[4], t->ev[5]};
    uint64_t pending_off[2] = {0, 0}, pending_n[2] = {0, 0};
    int slot = 0;
    for (uint64_t off = 0; off < n; off += chunk, slot ^= 1) {
        const uint64_t m = std::min(chunk, n - off);
        if (pending_n[slot]) {
            CU(cudaEventSynchronize(done[slot]));
            memcpy(counts + pending_off[slot], hc[slot], pending_n[slot] * 4);
        }
        memcpy(hk[slot], kmers + off, m * 8);
        CU(cudaMemcpyAsync(dk[slot], hk[slot], m * 8, cudaMemcpyHostToDevice, t->stream));
        if (int rc = km_query_batch_device(t, dk[slot], m, dc[slot], t->stream)) return rc;
        CU(cudaMemcpyAsync(hc[slot], dc[slot], m * 4, cudaMemcpyDeviceToHost, t->stream));
        CU(cudaEventRecord(done[slot], t->stream));
        pending_off[slot] = off; pending_n[slot] = m;
    }
    for (int b = 0; b < 2; ++b) {
        const int s2 = slot ^ b;   // older chunk first
        if (pending_n[s2]) {
            CU(cudaEventSynchronize(done[s2]));
            memcpy(counts + pending_off[s2], hc[s2], pending_n[s2] * 4);
        }
    }
    return 0;
}

static inline int base_code(char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

extern "C" int km_query_ascii(km_table* t, const char* kmers, uint64_t n, uint32_t* counts) {
    if (!t || (n && (!kmers || !counts))) return fail(KM_E_ARG, "km_query_ascii: bad argument");
    std::vector<uint64_t> packed(n);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t v = 0;
        for (int j = 0; j < t->k; ++j) {
            const int c = base_code(kmers[i * t->k + j]);
            if (c < 0) return fail(KM_E_ARG, "k-mer %llu holds a letter outside ACGT", (unsigned long long)i);
            v = (v << 2) | (uint64_t)c;
        }
        packed[i] = v;
    }
    return km_query_batch(t, packed.data(), n, counts);
}

extern "C" int km_get_child_batch(km_table* t, const uint64_t* kmers, uint64_t n, int forward, double ratio, int64_t floor_count,
                                  uint32_t* child_counts, uint8_t* child_mask) {
    if (!t || (n && (!kmers || !child_counts || !child_mask))) return fail(KM_E_ARG, "km_get_child_batch: bad argument");
    if (!n) return 0;
    CU(cudaSetDevice(t->device));
    if (int rc = t->dev.reserve(n * 25 + 2048)) return rc;
    t->dev.reset();
    uint64_t* dk = t->dev.take<uint64_t>(n);
    uint32_t* dc = t->dev.take<uint32_t>(4 * n);
    uint8_t* dm = t->dev.take<uint8_t>(n);
    CU(cudaMemcpyAsync(dk, kmers, n * 8, cudaMemcpyHostToDevice, t->stream));
    km_get_child_kernel<<<grid_for(t, n, 256, 8), 256, 0, t->stream>>>(t->view(), dk, n, forward, ratio, floor_count, dc, dm);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(child_counts, dc, 4 * n * 4, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaMemcpyAsync(child_mask, dm, n, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return 0;
}

// Host byte buffers recycled between calls: a fresh 12 MB buffer costs more in first-touch page faults than
// the text that goes into it costs to format.  Vectors keep their capacity while they sit in the cache.
struct VecCache {
    std::mutex m;
    std::vector<std::vector<char>> idle;
    size_t max_idle;
    explicit VecCache(size_t n) : max_idle(n) {}
    // the smallest idle vector that holds `want`, else the largest; `keep_size`: handed out as it came back
    // (a text buffer is used as raw storage: growing it through resize() would zero-fill it on every call)
    std::vector<char> get(size_t want, bool keep_size = false) {
        std::lock_guard<std::mutex> g(m);
        if (idle.empty()) return std::vector<char>();
        size_t best = 0;
        for (size_t i = 1; i < idle.size(); ++i) {
            const size_t a = idle[i].capacity(), b = idle[best].capacity();
            if ((a >= want && (b < want || a < b)) || (a < want && b < want && a > b)) best = i;
        }
        std::vector<char> v = std::move(idle[best]);
        idle.erase(idle.begin() + (long)best);
        if (!keep_size) v.clear();
        return v;
    }
    void put(std::vector<char>&& v) {
        if (!v.capacity() || v.capacity() > ((size_t)256 << 20)) return;
        std::lock_guard<std::mutex> g(m);
        if (idle.size() < max_idle) idle.push_back(std::move(v));
    }
};
static VecCache& piece_cache() { static VecCache c(128); return c; }
static VecCache& text_cache() { static VecCache c(8); return c; }
// the text a result holds: a cached vector used as a plain buffer
struct TextBuf {
    std::vector<char> v;
    PinBlock pin;                 // km_find_text with device-side formatting: the copies land here directly
    char* get() { return pin.base ? pin.base : v.data(); }
    void reset() { pin.drop(); if (v.capacity()) text_cache().put(std::move(v)); v = std::vector<char>(); }
    void reset(size_t bytes) {
        reset();
        v = text_cache().get(bytes, true);
        if (v.size() < bytes) v.resize(bytes + bytes / 8);     // first use of this size: the only time it is zero-filled
    }
    ~TextBuf() { reset(); }
};

// ---- find_mutation batch ------------------------------------------------------------------------
struct km_result {
    int n_targets = 0, k = 31;
    PinBlock head, body;            // per-target arrays; paths, rows, spelled sequences (+ graph arrays)
    Span<uint32_t> status;
    Span<int32_t> n_nodes, path_first, path_count, row_first, row_count, path_len;
    Span<int64_t> path_off, path_seq_off;
    Span<unsigned long long> lookups, used;
    Span<uint64_t> node_kmer;
    Span<uint32_t> node_count;
    Span<int32_t> path_pool;
    Span<km_row> rows;
    Span<char> seq_pool;            // spelled unique paths
    std::vector<int64_t> node_off, seq_off;
    std::string targets;            // concatenated target sequences (for Reference_sequence / deleted bases)
    float ms_h2d = 0, ms_walk = 0, ms_graph = 0, ms_d2h = 0, ms_total = 0;
    int n_launches = 0, n_retries = 0;
    bool has_graph = true;
    unsigned long long bytes_h2d = 0, bytes_d2h = 0;
    // the formatted text of all targets is built once and kept (km_result_format_all / km_result_text)
    mutable std::string fmt_key;
    mutable TextBuf text;
    mutable int64_t text_len = -1;
    long long dev_text_len = 0;          // km_find_text: bytes of text the device wrote for this (sub-)batch
    uint32_t dev_text_flags = 0;         // format.h flags: non-zero = the host must format this batch
    // km_find_text: the result of a pipelined run keeps its sub-batches and the joined text
    std::vector<std::unique_ptr<km_result>> parts;
    std::vector<uint32_t> all_status;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static uint32_t pow2_at_least(uint64_t x) { uint32_t p = 64; while (p < x) p <<= 1; return p; }

// A plan = one batch of targets laid out in HBM: inputs uploaded once, kernels launchable any
// number of times (bench.py times exactly that), results fetched on demand.
struct km_plan {
    km_table* t = nullptr;
    int n = 0;
    km_find_params prm{};
    std::string targets;
    std::vector<int64_t> seq_off, node_off, hash_off, pack_off;
    std::vector<int32_t> chunk_target, chunk_start;   // <= 32 consecutive reference k-mers each (ref_probe_chunk)
    std::vector<int32_t> extra;
    int64_t pool_cap = 0, seq_cap = 0, n_node = 0, n_hash = 0, n_code = 0;
    int32_t path_cap = 0, row_cap = 0, extra_max = 0;
    int grid_tiny = 1, grid_graph = 1, grid_large = 1;
    Arena own_dev, own_pin;
    Arena* dev = nullptr;
    Arena* pin = nullptr;
    WalkView W{};
    ResultView R{};
    ScratchLayout SL{};
    FindParams P{};
    char* d_seq_pool = nullptr;
    int64_t* d_path_seq_off = nullptr;
    char* state0 = nullptr;
    size_t state_bytes = 0;
    int n_launches = 0, n_retries = 0;
    bool launched = false;
    unsigned long long bytes_h2d = 0;
    size_t upload_bytes = 0;      // span of the input block on the device (plan_layout)
    const void* h_stage = nullptr;   // the staged copy of that block in pinned memory (plan_stage)
    const char* targets_ext = nullptr;   // km_find_text: the caller's sequences, valid for the whole call -- no private copy
    bool defer_upload = false;    // plan_init stops after staging: the caller enqueues (km_find_text, one thread at a time)
    // device-side text (km_find_text): query names + database name go up with the input block, FormatView F
    // describes the buffers of format.h
    bool fmt = false;
    const char* fmt_names = nullptr; const int64_t* fmt_name_off = nullptr; std::string fmt_db;
    FormatView F{};
    int64_t text_cap = 0;
    cudaStream_t stream = nullptr, side = nullptr;      // the table's own unless the plan runs on a lane
    cudaEvent_t* ev = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

static int plan_layout(km_plan* p) {
    km_table* t = p->t;
    const int k = t->k, n = p->n;
    p->node_off.assign(n + 1, 0);
    p->hash_off.assign(n + 1, 0);
    int maxcap = 1;
    for (int i = 0; i < n; ++i) {
        const int64_t len = p->seq_off[i + 1] - p->seq_off[i];
        const int L = (int)std::max<int64_t>(0, len - k + 1);
        const int cap = L + p->extra[i];
        maxcap = std::max(maxcap, cap);
        p->node_off[i + 1] = p->node_off[i] + cap;
        p->hash_off[i + 1] = p->hash_off[i] + pow2_at_least(2 * (uint64_t)cap + 256);
    }
    p->pack_off.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) p->pack_off[i + 1] = p->pack_off[i] + (p->seq_off[i + 1] - p->seq_off[i] + 15) / 16 + 2;
    const int64_t n_pack = p->pack_off[n];
    p->chunk_target.clear(); p->chunk_start.clear();
    for (int i = 0; i < n; ++i) {
        const int L = (int)std::max<int64_t>(0, p->seq_off[i + 1] - p->seq_off[i] - k + 1);
        for (int s0 = 0; s0 < L; s0 += 32) { p->chunk_target.push_back(i); p->chunk_start.push_back(s0); }
    }
    const size_t n_chunks = p->chunk_target.size();
    const int64_t n_node = p->n_node = p->node_off[n], n_hash = p->n_hash = p->hash_off[n], n_code = p->n_code = p->seq_off[n];
    p->grid_tiny = std::max(1, std::min(n, t->sm_count * KM_GRAPH_TINY_GRID));
    p->grid_graph = std::max(1, std::min(n, t->sm_count * KM_GRAPH_SMALL_GRID));
    p->grid_large = std::max(1, std::min(n, t->sm_count * 2));
    const ScratchLayout L0 = make_layout(maxcap, KM_MAX_PATHS, KM_MAX_PATHS, KM_MAX_COLS, 0);
    const int32_t path_cap = p->path_cap, row_cap = p->row_cap;
    const int64_t pool_cap = p->pool_cap, seq_cap = p->seq_cap;

    size_t need = 4096;
    auto acc = [&](size_t bytes) { need = align_up(need, 256) + bytes; };
    acc(n_code); acc(8 * (n + 1)); acc(8 * (n + 1)); acc(8 * (n + 1)); acc(4 * n_chunks); acc(4 * n_chunks);
    acc(4 * (size_t)n_pack); acc(8 * (n + 1)); acc(n);
    acc(8 * n_node); acc(4 * n_node); acc(4 * n_node); acc(16 * n_node);          // node arrays
    acc(8 * n_hash); acc(4 * n_hash); acc(4 * n_hash); acc(n_hash);                // visited sets
    acc(4 * n); acc(4 * n); acc(4 * n); acc(8 * n);                                // n_nodes n_kept status lookups
    for (int i = 0; i < 5; ++i) acc(4 * n);                                        // per-target result ints
    acc(8 * n_node); acc(4 * n_node);                                              // canonical nodes
    acc(8 * (size_t)path_cap); acc(4 * (size_t)path_cap); acc(8 * (size_t)path_cap); acc(4 * (size_t)pool_cap);
    acc(sizeof(Row) * (size_t)row_cap); acc((size_t)seq_cap); acc(64); acc(12 * (size_t)n + 64);
    acc(L0.stride * (size_t)p->grid_large);
    const size_t n_name = p->fmt ? (size_t)p->fmt_name_off[n] : 0;
    if (p->fmt) {
        // room for the text: rows carry two sequences of about the target's length each
        p->text_cap = 16 * n_code + 512ll * n + (int64_t)row_cap * (int64_t)(p->fmt_db.size() + 64) + (1 << 16);
        acc(n_name); acc(8 * (size_t)(n + 1)); acc(p->fmt_db.size() + 1);
        acc(4 * (size_t)row_cap); acc(4 * (size_t)row_cap); acc((size_t)KM_FMT_ROW_BYTES * (size_t)row_cap);
        acc(8 * (size_t)n); acc(8 * (size_t)(n + 1)); acc((size_t)p->text_cap); acc(64);
    }
    if (int rc = p->dev->reserve(need + 16384)) return rc;
    p->dev->reset();
    Arena& A = *p->dev;
    WalkView& W = p->W;
    W.n_targets = n;
    W.codes = A.take<uint8_t>(n_code);
    // codes .. chunk_start are taken in the order (and with the alignment) plan_upload uses for its pinned staging
    // block, so the whole input goes up with ONE copy
    W.seq_off = A.take<int64_t>(n + 1); W.node_off = A.take<int64_t>(n + 1); W.hash_off = A.take<int64_t>(n + 1);
    W.pack_off = A.take<int64_t>(n + 1);
    W.chunk_target = A.take<int32_t>(n_chunks); W.chunk_start = A.take<int32_t>(n_chunks); W.n_chunks = (int)n_chunks;
    const char* input_end = (const char*)(W.chunk_start + n_chunks);
    if (p->fmt) {
        char* dn = A.take<char>(n_name);
        int64_t* dno = A.take<int64_t>(n + 1);
        char* ddb = A.take<char>(p->fmt_db.size() + 1);
        p->F.names = dn; p->F.name_off = dno; p->F.db_name = ddb; p->F.db_len = (int)p->fmt_db.size();
        input_end = ddb + p->fmt_db.size() + 1;
    }
    p->upload_bytes = (size_t)(input_end - (const char*)W.codes);
    W.pack = A.take<uint32_t>((size_t)n_pack); W.pre_bad = A.take<uint8_t>(n);
    W.node_kmer = A.take<uint64_t>(n_node); W.node_count = A.take<uint32_t>(n_node);
    W.node_slot = A.take<uint32_t>(n_node); W.node_kid = A.take<uint32_t>(4 * n_node);
    W.hkey = A.take<uint64_t>(n_hash); W.hval = A.take<uint32_t>(n_hash); W.hmeta = A.take<uint32_t>(n_hash);
    W.hflag = A.take<uint8_t>(n_hash);
    // per-target state and result ints are contiguous so one memset clears them
    p->state0 = A.take<char>(0);
    W.n_nodes = A.take<int32_t>(n); W.n_kept = A.take<int32_t>(n); W.status = A.take<uint32_t>(n);
    W.lookups = A.take<unsigned long long>(n);
    ResultView& R = p->R;
    R.t_n = A.take<int32_t>(n); R.t_n_paths = A.take<int32_t>(n); R.t_path_first = A.take<int32_t>(n);
    R.t_n_rows = A.take<int32_t>(n); R.t_row_first = A.take<int32_t>(n);
    // the pool cursors and the formatter's flags / total sit in the same block: one memset, one copy back
    R.used = A.take<unsigned long long>(8);     // [0..3] pool cursors, [4..6] work counters of the three graph passes
    p->F.flags = A.take<uint32_t>(16);          // [0] flags, [2..3] total bytes of text (64 bit)
    p->state_bytes = (size_t)(A.take<char>(0) - p->state0);
    R.out_kmer = A.take<uint64_t>(n_node); R.out_count = A.take<uint32_t>(n_node);
    R.path_off = A.take<int64_t>(path_cap); R.path_len = A.take<int32_t>(path_cap);
    p->d_path_seq_off = A.take<int64_t>(path_cap);
    R.pool = A.take<int32_t>(pool_cap); R.path_cap = path_cap; R.pool_cap = pool_cap;
    R.rows = A.take<Row>(row_cap); R.row_cap = row_cap;
    p->d_seq_pool = A.take<char>(seq_cap);
    R.seq_pool = p->d_seq_pool; R.path_seq_off = p->d_path_seq_off; R.seq_cap = seq_cap;
    R.sched_order = A.take<int32_t>(3 * (size_t)n); R.sched_count = A.take<int32_t>(4);
    p->SL = L0;
    p->SL.base = A.take<char>(L0.stride * (size_t)p->grid_large);
    if (p->fmt) {
        p->F.row_len = A.take<int32_t>(row_cap); p->F.row_pos = A.take<int32_t>(row_cap);
        p->F.row_num = A.take<char>((size_t)KM_FMT_ROW_BYTES * (size_t)row_cap);
        p->F.t_len = A.take<int64_t>(n); p->F.t_off = A.take<int64_t>(n + 1);
        p->F.text = A.take<char>((size_t)p->text_cap); p->F.text_cap = p->text_cap;
    }
    p->P.ratio = p->prm.ratio; p->P.count = p->prm.count; p->P.max_stack = p->prm.steps;
    p->P.max_break = p->prm.branchs; p->P.max_node = p->prm.nodes;
    return 0;
}

// the host half of the upload: every input of the batch into ONE pinned staging block
static int plan_stage(km_plan* p, cudaStream_t s) {
    const int n = p->n;
    const size_t n_chunks = p->chunk_target.size();
    const size_t n_name = p->fmt ? (size_t)p->fmt_name_off[n] : 0;
    if (int rc = p->pin->reserve((size_t)p->n_code + 40 * (size_t)(n + 1) + 8 * n_chunks + n_name + p->fmt_db.size() + 16384)) return rc;
    p->pin->reset();
    uint8_t* h_codes = p->pin->take<uint8_t>(p->n_code);
    int64_t* h_seq_off = p->pin->take<int64_t>(n + 1);
    int64_t* h_node_off = p->pin->take<int64_t>(n + 1);
    int64_t* h_hash_off = p->pin->take<int64_t>(n + 1);
    int64_t* h_pack_off = p->pin->take<int64_t>(n + 1);
    memcpy(h_pack_off, p->pack_off.data(), 8 * (n + 1));
    int32_t* h_ct = p->pin->take<int32_t>(n_chunks);
    int32_t* h_cs = p->pin->take<int32_t>(n_chunks);
    if (n_chunks) { memcpy(h_ct, p->chunk_target.data(), 4 * n_chunks); memcpy(h_cs, p->chunk_start.data(), 4 * n_chunks); }
    memcpy(h_codes, p->targets_ext ? p->targets_ext : p->targets.data(), p->n_code);   // letters; km_encode_kernel turns them into codes on the device
    memcpy(h_seq_off, p->seq_off.data(), 8 * (n + 1));
    memcpy(h_node_off, p->node_off.data(), 8 * (n + 1));
    memcpy(h_hash_off, p->hash_off.data(), 8 * (n + 1));
    CU(cudaEventRecord(p->ev[0], s));
    const char* h_end = (const char*)(h_cs + n_chunks);
    if (p->fmt) {
        char* hn = p->pin->take<char>(n_name);
        int64_t* hno = p->pin->take<int64_t>(n + 1);
        char* hdb = p->pin->take<char>(p->fmt_db.size() + 1);
        memcpy(hn, p->fmt_names, n_name);
        memcpy(hno, p->fmt_name_off, 8 * (size_t)(n + 1));
        memcpy(hdb, p->fmt_db.c_str(), p->fmt_db.size() + 1);
        h_end = hdb + p->fmt_db.size() + 1;
    }
    if ((size_t)(h_end - (const char*)h_codes) != p->upload_bytes)
        return fail(KM_E_ARG, "internal: staging block and device input block differ in layout");
    p->h_stage = h_codes;
    p->bytes_h2d = (unsigned long long)p->n_code + 32ull * (n + 1) + 8ull * n_chunks + (p->fmt ? n_name + 8ull * (n + 1) + p->fmt_db.size() : 0ull);
    return 0;
}

// the CUDA half of the upload: one copy of the staged block, then the packing kernel
static int plan_upload_enqueue(km_plan* p, cudaStream_t s) {
    const int n = p->n;
    CU(cudaMemcpyAsync((void*)p->W.codes, p->h_stage, p->upload_bytes, cudaMemcpyHostToDevice, s));
    if (n) {
        km_encode_kernel<<<(n + 7) / 8, 256, 0, s>>>(const_cast<uint8_t*>(p->W.codes), p->W.seq_off, const_cast<uint32_t*>(p->W.pack),
                                                     p->W.pack_off, const_cast<uint8_t*>(p->W.pre_bad), n);
        CU(cudaGetLastError());
    }
    return 0;
}

static int plan_upload(km_plan* p, cudaStream_t s) {
    if (int rc = plan_stage(p, s)) return rc;
    return plan_upload_enqueue(p, s);
}

// memsets + the two kernels, asynchronously on `s`
static int plan_launch(km_plan* p, cudaStream_t s) {
    km_table* t = p->t;
    if (p->n == 0) return 0;
    const bool timed = !p->fmt;          // km_find_text enqueues as little as it can: no per-phase events
    if (timed) CU(cudaEventRecord(p->ev[1], s));
    CU(cudaMemsetAsync(p->state0, 0, p->state_bytes, s));
    if (p->W.n_chunks) {
        km_ref_probe_kernel<<<(p->W.n_chunks + KM_PROBE_WARPS - 1) / KM_PROBE_WARPS, 32 * KM_PROBE_WARPS, 0, s>>>(t->view(), p->W, p->P);
        CU(cudaGetLastError());
    }
    if (timed) CU(cudaEventRecord(p->ev[6], s));
    km_walk_small_kernel<<<(p->n + KM_WALK_WARPS - 1) / KM_WALK_WARPS, 32 * KM_WALK_WARPS, 0, s>>>(t->view(), p->W, p->P);
    CU(cudaGetLastError());
    km_walk_kernel<<<(p->n + KM_WALK_WARPS - 1) / KM_WALK_WARPS, 32 * KM_WALK_WARPS, 0, s>>>(t->view(), p->W, p->P);
    CU(cudaGetLastError());
    if (timed) CU(cudaEventRecord(p->ev[2], s));
    // shared-memory pass first, then the general pass for large or deferred targets
    km_schedule_kernel<<<1, 1024, 0, s>>>(p->W, p->R);
    CU(cudaGetLastError());
    // the two shared-memory passes side by side (their CTAs co-reside; each pass's tail fills with the other)
    CU(cudaEventRecord(p->fork, s));
    CU(cudaStreamWaitEvent(p->side, p->fork, 0));
    km_graph_kernel<KM_SMALL_NODES><<<p->grid_graph, KM_CTA, class_layout(KM_SMALL_NODES).stride, p->side>>>(t->view(), p->W, p->SL, p->R);
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->join, p->side));
    km_graph_kernel<KM_TINY_NODES><<<p->grid_tiny, KM_CTA, class_layout(KM_TINY_NODES).stride, s>>>(t->view(), p->W, p->SL, p->R);
    CU(cudaGetLastError());
    CU(cudaStreamWaitEvent(s, p->join, 0));
    km_graph_kernel<0><<<p->grid_large, KM_CTA, 0, s>>>(t->view(), p->W, p->SL, p->R);
    CU(cudaGetLastError());
    if (timed) CU(cudaEventRecord(p->ev[3], s));
    p->n_launches += 7;
    if (p->fmt) {
        km_format_measure_kernel<<<(p->n + 3) / 4, 128, 0, s>>>(p->W, p->R, p->F, t->k);
        CU(cudaGetLastError());
        km_format_scan_kernel<<<1, 1024, 0, s>>>(p->F, p->n);
        CU(cudaGetLastError());
        km_format_write_kernel<<<(p->n + 3) / 4, 128, 0, s>>>(p->W, p->R, p->F, t->k);
        CU(cudaGetLastError());
        p->n_launches += 3;
    }
    p->launched = true;
    return 0;
}

// D2H of per-target ints, then exactly the used extents.  `want_graph` also brings back the
// node arrays and index paths (needed by the MutationFinder attribute views and the parity tests;
// the TSV formatter only needs rows + spelled paths).
static int plan_download(km_plan* p, cudaStream_t s, km_result* res, bool want_graph, bool head_only = false) {
    km_table* t = p->t;
    const int n = p->n;
    const WalkView& W = p->W;
    const ResultView& R = p->R;
    res->n_targets = n; res->k = t->k;
    // the per-target state and result ints sit back to back on the device (plan_layout): ONE copy brings the
    // block into pinned memory and the result's arrays are views into it
    if (int rc = res->head.reserve(t->pool, p->state_bytes + 1024)) return rc;
    Span<char> blk = res->head.take<char>(p->state_bytes);
    auto view = [&](const void* dev_ptr) { return blk.data() + ((const char*)dev_ptr - p->state0); };
    res->status.p = (uint32_t*)view(W.status); res->status.n = (size_t)n;
    res->n_nodes.p = (int32_t*)view(R.t_n); res->n_nodes.n = (size_t)n;
    res->path_count.p = (int32_t*)view(R.t_n_paths); res->path_count.n = (size_t)n;
    res->path_first.p = (int32_t*)view(R.t_path_first); res->path_first.n = (size_t)n;
    res->row_count.p = (int32_t*)view(R.t_n_rows); res->row_count.n = (size_t)n;
    res->row_first.p = (int32_t*)view(R.t_row_first); res->row_first.n = (size_t)n;
    res->lookups.p = (unsigned long long*)view(W.lookups); res->lookups.n = (size_t)n;
    res->used.p = (unsigned long long*)view(R.used); res->used.n = 4;
    unsigned long long* used = res->used.data();
    const uint32_t* fmt_info = (const uint32_t*)view(p->F.flags);     // device text: [0] flags, [2..3] total bytes
    if (n) CU(cudaMemcpyAsync(blk.data(), p->state0, p->state_bytes, cudaMemcpyDeviceToHost, s));
    else memset(blk.data(), 0, p->state_bytes);
    if (head_only) {
        CU(cudaStreamSynchronize(s));
        long long total = 0;
        memcpy(&total, fmt_info + 2, 8);
        res->dev_text_len = total; res->dev_text_flags = fmt_info[0];
        res->has_graph = false;
        res->bytes_h2d = p->bytes_h2d;
        res->bytes_d2h = (unsigned long long)p->state_bytes;
        res->text_len = -1; res->text.reset(); res->fmt_key.clear();
        return 0;
    }
    CU(cudaStreamSynchronize(s));
    const size_t n_paths = std::min<unsigned long long>(used[0], p->path_cap), n_pool = std::min<unsigned long long>(used[1], p->pool_cap);
    const size_t n_rows = std::min<unsigned long long>(used[2], p->row_cap), n_seq = std::min<unsigned long long>(used[3], p->seq_cap);
    const size_t n_node = want_graph ? (size_t)p->n_node : 0, n_pool_c = want_graph ? n_pool : 0;
    if (int rc = res->body.reserve(t->pool, 20 * n_paths + sizeof(Row) * n_rows + n_seq + 4 * n_pool_c + 12 * n_node + 64 * 10)) return rc;
    res->path_off = res->body.take<int64_t>(n_paths); res->path_len = res->body.take<int32_t>(n_paths);
    res->path_seq_off = res->body.take<int64_t>(n_paths);
    res->rows = res->body.take<km_row>(n_rows); res->seq_pool = res->body.take<char>(n_seq);
    res->path_pool = res->body.take<int32_t>(n_pool_c);
    res->node_kmer = res->body.take<uint64_t>(n_node); res->node_count = res->body.take<uint32_t>(n_node);
    res->node_off = p->node_off;
    if (n_paths) {
        CU(cudaMemcpyAsync(res->path_off.data(), R.path_off, 8 * n_paths, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(res->path_len.data(), R.path_len, 4 * n_paths, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(res->path_seq_off.data(), p->d_path_seq_off, 8 * n_paths, cudaMemcpyDeviceToHost, s));
    }
    if (n_rows) CU(cudaMemcpyAsync(res->rows.data(), R.rows, sizeof(Row) * n_rows, cudaMemcpyDeviceToHost, s));
    if (n_seq) CU(cudaMemcpyAsync(res->seq_pool.data(), p->d_seq_pool, n_seq, cudaMemcpyDeviceToHost, s));
    res->has_graph = want_graph;
    if (want_graph) {
        if (n_pool) CU(cudaMemcpyAsync(res->path_pool.data(), R.pool, 4 * n_pool, cudaMemcpyDeviceToHost, s));
        if (n_node) {
            CU(cudaMemcpyAsync(res->node_kmer.data(), R.out_kmer, 8 * n_node, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(res->node_count.data(), R.out_count, 4 * n_node, cudaMemcpyDeviceToHost, s));
        }
    }
    res->bytes_h2d = p->bytes_h2d;
    res->bytes_d2h = (unsigned long long)p->state_bytes + 20ull * n_paths + sizeof(Row) * n_rows + n_seq +
                     (want_graph ? 4ull * n_pool + 12ull * p->n_node : 0ull);
    res->text_len = -1; res->text.reset(); res->fmt_key.clear();
    if (!p->fmt) CU(cudaEventRecord(p->ev[4], s));
    CU(cudaStreamSynchronize(s));
    float ms;
    if (n && !p->fmt) {
        CU(cudaEventElapsedTime(&ms, p->ev[0], p->ev[1])); res->ms_h2d = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[1], p->ev[2])); res->ms_walk = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[2], p->ev[3])); res->ms_graph = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[3], p->ev[4])); res->ms_d2h = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[0], p->ev[4])); res->ms_total = ms;
    }
    return 0;
}

// a plan on a lane borrows the lane's host vectors (and hands them back, km_find_text) for their capacity
static void plan_swap_vecs(km_plan* p, km_table::PlanVecs& v) {
    p->seq_off.swap(v.seq_off); p->node_off.swap(v.node_off); p->hash_off.swap(v.hash_off); p->pack_off.swap(v.pack_off);
    p->chunk_target.swap(v.chunk_target); p->chunk_start.swap(v.chunk_start); p->extra.swap(v.extra);
}

static int plan_init(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params, km_plan* p,
                     bool borrow_arena, km_table::Lane* lane = nullptr) {
    if (lane) plan_swap_vecs(p, lane->vecs);
    p->t = t; p->n = n; p->prm = *params;
    p->stream = lane ? lane->stream : t->stream;
    p->side = lane ? lane->side : t->side;
    p->ev = lane ? lane->ev : t->ev;
    p->fork = lane ? lane->fork : t->fork;
    p->join = lane ? lane->join : t->join;
    if (p->prm.steps > 60000 || p->prm.branchs > 250) return fail(KM_E_ARG, "steps must be <= 60000 and branchs <= 250");
    const int64_t total = n ? offsets[n] : 0;
    if (!p->targets_ext) p->targets.assign(seqs ? seqs : "", (size_t)total);
    p->seq_off.assign(1, 0);
    if (n) p->seq_off.assign(offsets, offsets + n + 1);
    int64_t n_ref = 0;
    for (int i = 0; i < n; ++i) n_ref += std::max<int64_t>(0, offsets[i + 1] - offsets[i] - t->k + 1);
    p->extra.assign((size_t)n, p->prm.extra_nodes > 0 ? p->prm.extra_nodes : 256);
    p->path_cap = std::max(64, 8 * n);
    p->row_cap = std::max(64, 16 * n);
    p->pool_cap = std::max<int64_t>(1 << 16, 6 * (n_ref + 64ll * n));
    p->seq_cap = p->pool_cap + (int64_t)p->path_cap * t->k;
    p->extra_max = std::max(1024, p->prm.nodes + 4 * p->prm.steps + 4096);
    p->own_pin.host = true;
    p->dev = lane ? &lane->dev : borrow_arena ? &t->dev_find : &p->own_dev;
    p->pin = lane ? &lane->pin : borrow_arena ? &t->pin_find : &p->own_pin;
    trace_here("  init: copies");
    if (int rc = plan_layout(p)) return rc;
    trace_here("  init: layout");
    const int rc_up = p->defer_upload ? plan_stage(p, p->stream) : plan_upload(p, p->stream);
    trace_here("  init: upload");
    return rc_up;
}

// fetch with the capacity-retry loop: targets whose exploration overflowed get 8x the node
// capacity, exhausted pools grow 4x, and the batch is re-run
static int plan_fetch(km_plan* p, km_result* res, bool want_graph, bool head_only = false) {
    km_table* t = p->t;
    for (int attempt = 0; attempt < 12; ++attempt) {
        if (!p->launched) if (int rc = plan_launch(p, p->stream)) return rc;
        if (int rc = plan_download(p, p->stream, res, want_graph, head_only)) return rc;
        bool again = false, pool_over = false;
        for (int i = 0; i < p->n; ++i) {
            if (res->status[i] & KM_ST_NODE_OVERFLOW) {
                if (p->extra[i] >= p->extra_max) return fail(KM_E_LIMIT, "target %d explores more than %d nodes", i, p->extra_max);
                p->extra[i] = (int32_t)std::min<int64_t>(p->extra_max, (int64_t)p->extra[i] * 8);
                again = true;
            }
            if (res->status[i] & KM_ST_PATH_OVERFLOW) pool_over = true;
        }
        if (pool_over) {
            p->pool_cap *= 4; p->path_cap *= 4; p->row_cap *= 4;
            p->seq_cap = p->pool_cap + (int64_t)p->path_cap * t->k;
            again = true;
        }
        res->n_launches = p->n_launches; res->n_retries = p->n_retries;
        if (!again) return 0;
        p->n_retries++;
        p->launched = false;
        if (int rc = plan_layout(p)) return rc;
        if (int rc = plan_upload(p, p->stream)) return rc;
    }
    return fail(KM_E_LIMIT, "km_find: capacities still exceeded after 12 attempts");
}

extern "C" int km_find_plan_create(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params,
                                   km_plan** out) {
    if (!t || !out || n < 0 || (n && (!seqs || !offsets)) || !params) return fail(KM_E_ARG, "km_find_plan_create: bad argument");
    CU(cudaSetDevice(t->device));
    km_plan* p = new km_plan();
    if (int rc = plan_init(t, seqs, offsets, n, params, p, false)) { delete p; return rc; }
    CU(cudaStreamSynchronize(t->stream));
    *out = p;
    return 0;
}

extern "C" int km_find_plan_launch(km_plan* p, void* stream) {
    if (!p) return fail(KM_E_ARG, "null plan");
    CU(cudaSetDevice(p->t->device));
    return plan_launch(p, stream ? (cudaStream_t)stream : p->stream);
}

extern "C" int km_find_plan_last_ms(km_plan* p, float* walk_ms, float* graph_ms) {
    if (!p || !p->launched) return fail(KM_E_ARG, "km_find_plan_last_ms: nothing launched");
    CU(cudaSetDevice(p->t->device));
    CU(cudaEventSynchronize(p->ev[3]));
    if (walk_ms) CU(cudaEventElapsedTime(walk_ms, p->ev[1], p->ev[2]));
    if (graph_ms) CU(cudaEventElapsedTime(graph_ms, p->ev[2], p->ev[3]));
    return 0;
}

extern "C" int km_find_plan_kernel_ms(km_plan* p, float* out3) {
    if (!p || !p->launched || !out3) return fail(KM_E_ARG, "km_find_plan_kernel_ms: nothing launched");
    CU(cudaSetDevice(p->t->device));
    CU(cudaEventSynchronize(p->ev[3]));
    CU(cudaEventElapsedTime(&out3[0], p->ev[1], p->ev[6]));    // memsets + reference probe
    CU(cudaEventElapsedTime(&out3[1], p->ev[6], p->ev[2]));    // the two walk kernels
    CU(cudaEventElapsedTime(&out3[2], p->ev[2], p->ev[3]));    // the two graph kernels
    return 0;
}

extern "C" int km_find_plan_fetch(km_plan* p, int want_graph, km_result** out) {
    if (!p || !out) return fail(KM_E_ARG, "null argument");
    CU(cudaSetDevice(p->t->device));
    CU(cudaDeviceSynchronize());         // launches may have gone to a caller's stream
    km_result* res = new km_result();
    res->targets = p->targets;
    res->seq_off = p->seq_off;
    if (int rc = plan_fetch(p, res, want_graph != 0)) { delete res; return rc; }
    *out = res;
    return 0;
}

extern "C" void km_find_plan_free(km_plan* p) {
    if (!p) return;
    cudaSetDevice(p->t->device);
    p->own_dev.release();
    p->own_pin.release();
    delete p;
}

extern "C" int km_find_batch(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params,
                             km_result** out) {
    if (!t || !out || n < 0 || (n && (!seqs || !offsets)) || !params) return fail(KM_E_ARG, "km_find_batch: bad argument");
    CU(cudaSetDevice(t->device));
    km_plan plan;
    if (int rc = plan_init(t, seqs, offsets, n, params, &plan, true)) return rc;
    km_result* res = new km_result();
    res->seq_off = plan.seq_off;
    if (int rc = plan_fetch(&plan, res, (params->flags & KM_FIND_NO_GRAPH) == 0)) { delete res; return rc; }
    res->targets.swap(plan.targets);      // after the fetch: a capacity retry uploads the letters again
    *out = res;
    return 0;
}

extern "C" int km_result_get(const km_result* r, km_result_view* v) {
    if (!r || !v) return fail(KM_E_ARG, "null argument");
    memset(v, 0, sizeof(*v));
    v->n_targets = r->n_targets; v->n_paths = (int32_t)r->path_off.size(); v->n_rows = (int32_t)r->rows.size(); v->k = r->k;
    if (!r->parts.empty()) {      // a km_find_text result: the text and the per-target status are what it holds
        v->status = r->all_status.data();
        v->ms_h2d = r->ms_h2d; v->ms_walk = r->ms_walk; v->ms_graph = r->ms_graph; v->ms_d2h = r->ms_d2h; v->ms_total = r->ms_total;
        v->n_launches = r->n_launches; v->n_retries = r->n_retries; v->has_graph = 0;
        v->bytes_h2d = r->bytes_h2d; v->bytes_d2h = r->bytes_d2h;
        return 0;
    }
    v->status = r->status.data(); v->n_nodes = r->n_nodes.data(); v->node_off = r->node_off.data();
    v->node_kmer = r->node_kmer.data(); v->node_count = r->node_count.data();
    v->path_first = r->path_first.data(); v->path_count = r->path_count.data();
    v->path_off = r->path_off.data(); v->path_len = r->path_len.data(); v->path_pool = r->path_pool.data();
    v->row_first = r->row_first.data(); v->row_count = r->row_count.data(); v->rows = r->rows.data();
    v->lookups = reinterpret_cast<const uint64_t*>(r->lookups.data());
    v->ms_h2d = r->ms_h2d; v->ms_walk = r->ms_walk; v->ms_graph = r->ms_graph; v->ms_d2h = r->ms_d2h; v->ms_total = r->ms_total;
    v->n_launches = r->n_launches; v->n_retries = r->n_retries; v->has_graph = r->has_graph ? 1 : 0;
    v->bytes_h2d = r->bytes_h2d; v->bytes_d2h = r->bytes_d2h;
    return 0;
}

extern "C" void km_result_free(km_result* r) { delete r; }

// ---- row formatting (PathQuant.Path.__str__, MutationFinder.get_paths) ------------------------
static const char* TYPE_NAME[6] = {"Reference", "Substitution", "ITD", "Indel", "Insertion", "Deletion"};

// "%.{prec}f" of a double, digit for digit what Python / glibc print (the exact binary value rounded
// half-to-even at the last printed digit), without snprintf: |v| = m * 2^e with m < 2^53, so
// m * 10^prec fits 64 bits for prec <= 3 and the rounding is decided on integers.
static char* put_uint(char* o, unsigned long long v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}
static char* put_int(char* o, long long v) {
    if (v < 0) { *o++ = '-'; return put_uint(o, 0ull - (unsigned long long)v); }
    return put_uint(o, (unsigned long long)v);
}
static char* put_fixed(char* o, double v, int prec) {
    if (std::isnan(v)) { memcpy(o, "nan", 3); return o + 3; }          // Python prints nan without a sign
    if (std::isinf(v)) { if (v < 0) *o++ = '-'; memcpy(o, "inf", 3); return o + 3; }
    const double a = fabs(v);
    if (prec > 3 || a >= 4503599627370496.0) return o + snprintf(o, 400, "%.*f", prec, v);
    if (std::signbit(v)) *o++ = '-';
    static const unsigned long long P10[4] = {1ull, 10ull, 100ull, 1000ull};
    int e;
    const double fr = frexp(a, &e);                                    // a = fr * 2^e, fr in [0.5, 1)
    unsigned long long q = 0;
    if (a != 0.0) {
        const unsigned long long m = (unsigned long long)ldexp(fr, 53);   // exact 53-bit integer
        const int e2 = e - 53;                                            // a = m * 2^e2, e2 < 0 here
        const unsigned long long scaled = m * P10[prec];
        const int sh = -e2;
        if (sh <= 0) q = scaled << (-sh);
        else if (sh >= 64) q = 0;
        else {
            q = scaled >> sh;
            const unsigned long long rem = scaled & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
            if (rem > half || (rem == half && (q & 1ull))) ++q;
        }
    }
    const unsigned long long ip = q / P10[prec], fp = q % P10[prec];
    o = put_uint(o, ip);
    if (prec > 0) {
        *o++ = '.';
        for (int d = prec - 1; d >= 0; --d) *o++ = (char)('0' + (fp / P10[d]) % 10);
    }
    return o;
}

// common.natsortkey (common.py:95-116) on two strings without building the token lists:
// re.split('([0-9]+)', key) alternates text / digit runs starting and ending with a (possibly empty)
// text chunk; text compares lower-cased, digit runs as integers, and a list that is a prefix of the
// other sorts first.
static int nat_cmp(const char* a, size_t na, const char* b, size_t nb) {
    size_t i = 0, j = 0;
    for (;;) {
        // text chunks
        for (;;) {
            const bool ea = i >= na || isdigit((unsigned char)a[i]), eb = j >= nb || isdigit((unsigned char)b[j]);
            if (ea || eb) { if (ea != eb) return ea ? -1 : 1; break; }
            const int ca = tolower((unsigned char)a[i]), cb = tolower((unsigned char)b[j]);
            if (ca != cb) return ca < cb ? -1 : 1;
            ++i; ++j;
        }
        const bool enda = i >= na, endb = j >= nb;
        if (enda || endb) return enda == endb ? 0 : (enda ? -1 : 1);
        // digit runs as integers of any length
        size_t i2 = i, j2 = j;
        while (i2 < na && isdigit((unsigned char)a[i2])) ++i2;
        while (j2 < nb && isdigit((unsigned char)b[j2])) ++j2;
        size_t ia = i, jb = j;
        while (ia + 1 < i2 && a[ia] == '0') ++ia;
        while (jb + 1 < j2 && b[jb] == '0') ++jb;
        if (i2 - ia != j2 - jb) return i2 - ia < j2 - jb ? -1 : 1;
        const int c = memcmp(a + ia, b + jb, i2 - ia);
        if (c) return c < 0 ? -1 : 1;
        i = i2; j = j2;
    }
}

struct FmtRow {
    const km_row* w;
    const char* name; uint32_t name_len;       // variant name ("" for Reference)
    const char* line; uint32_t line_len;
};

// the rows of target tg: text into `arena` (unsorted), then sorted as MutationFinder.get_paths does
// (:825-829) and appended to `out`
static void format_rows_of(const km_result* r, int32_t tg, const char* db_name, size_t db_len, const char* qn, size_t qn_len,
                           std::vector<char>& arena, std::vector<FmtRow>& rows, std::vector<char>& out) {
    const int k = r->k;
    const char* tseq = r->targets.data() + r->seq_off[tg];
    const int nrow = r->row_count[tg];
    rows.clear();
    if (nrow <= 0) return;
    // upper bound of this target's text
    size_t need = 0;
    for (int i = 0; i < nrow; ++i) {
        const km_row& w = r->rows[r->row_first[tg] + i];
        need += db_len + qn_len + 256 + (size_t)(w.del_len + w.ins_len) + (size_t)(w.var_end - w.var_begin + k) +
                (size_t)(w.ref_end - w.ref_begin + k);
    }
    if (arena.size() < need) arena.resize(need + need / 2);
    char* o = arena.data();
    for (int i = 0; i < nrow; ++i) {
        const km_row& w = r->rows[r->row_first[tg] + i];
        const char* pseq = r->seq_pool.data() + r->path_seq_off[w.path_id];
        FmtRow fr;
        fr.w = &w;
        fr.line = o;
        memcpy(o, db_name, db_len); o += db_len; *o++ = '\t';
        memcpy(o, qn, qn_len); o += qn_len; *o++ = '\t';
        const size_t tl = strlen(TYPE_NAME[w.type]);
        memcpy(o, TYPE_NAME[w.type], tl); o += tl; *o++ = '\t';
        fr.name = o;
        if (w.type != 0) {      // "{}\t{}:{}:{}" (MutationFinder.py:483-488); Reference -> "Reference\t"
            o = put_int(o, w.name_start); *o++ = ':';
            for (int j = 0; j < w.del_len; ++j) *o++ = (char)tolower((unsigned char)tseq[w.del_begin + j + k - 1]);
            *o++ = '/';
            memcpy(o, pseq + w.ins_begin + k - 1, (size_t)w.ins_len); o += w.ins_len;
            *o++ = ':'; o = put_int(o, w.name_end);
        }
        fr.name_len = (uint32_t)(o - fr.name);
        *o++ = '\t';
        o = put_fixed(o, w.rvaf, 3); *o++ = '\t';
        o = put_fixed(o, w.expr, 1); *o++ = '\t';
        o = put_int(o, (long long)w.min_cov); *o++ = '\t';
        o = put_int(o, w.start_off); *o++ = '\t';
        if (w.var_end > w.var_begin) { const size_t n = (size_t)(w.var_end - w.var_begin + k - 1); memcpy(o, pseq + w.var_begin, n); o += n; }
        *o++ = '\t';
        o = put_fixed(o, w.ref_expr, 1); *o++ = '\t';
        if (w.ref_end > w.ref_begin) { const size_t n = (size_t)(w.ref_end - w.ref_begin + k - 1); memcpy(o, tseq + w.ref_begin, n); o += n; }
        *o++ = '\t';
        if (w.kind == 0) { memcpy(o, "vs_ref", 6); o += 6; }
        else { memcpy(o, "cluster ", 8); o += 8; o = put_int(o, w.cluster_id); memcpy(o, " n=", 3); o += 3; o = put_int(o, w.cluster_n); }
        *o++ = '\n';
        fr.line_len = (uint32_t)(o - fr.line);
        rows.push_back(fr);
    }
    // key = natsortkey(*info.split(' '), query, variant_name, type, min_coverage, rev_ix=[0]) (:825-829):
    // info is "vs_ref" or "cluster <i> n=<j>"; the first word compares REVERSED (vs_ref rows first),
    // then the words (numbers as numbers), the query (equal inside a target), the variant name, the
    // type, Min_coverage; a key that is a prefix of the other sorts first (vs_ref has one word, a
    // cluster three, but those never tie on the first word).
    if (nrow > 1) {
        auto less = [&](const FmtRow& x, const FmtRow& y) {
            const km_row& a = *x.w; const km_row& b = *y.w;
            if (a.kind != b.kind) return a.kind < b.kind;                       // "vs_ref" > "cluster", reversed
            if (a.kind != 0) {
                if (a.cluster_id != b.cluster_id) return a.cluster_id < b.cluster_id;
                if (a.cluster_n != b.cluster_n) return a.cluster_n < b.cluster_n;
            }
            int c = nat_cmp(x.name, x.name_len, y.name, y.name_len);
            if (c) return c < 0;
            c = nat_cmp(TYPE_NAME[a.type], strlen(TYPE_NAME[a.type]), TYPE_NAME[b.type], strlen(TYPE_NAME[b.type]));
            if (c) return c < 0;
            // Min_coverage prints as a decimal integer; the counts are never negative
            return a.min_cov < b.min_cov;
        };
        if (nrow <= 16) {                   // stable insertion sort: no temporary buffer for the usual 2-3 rows
            for (int i = 1; i < nrow; ++i) {
                FmtRow cur = rows[(size_t)i];
                int j = i;
                while (j > 0 && less(cur, rows[(size_t)j - 1])) { rows[(size_t)j] = rows[(size_t)j - 1]; --j; }
                rows[(size_t)j] = cur;
            }
        } else {
            std::stable_sort(rows.begin(), rows.end(), less);
        }
    }
    for (const FmtRow& f : rows) out.insert(out.end(), f.line, f.line + f.line_len);
}

extern "C" int64_t km_result_format_target(const km_result* r, int32_t tg, const char* db_name, const char* query_name, char* buf,
                                           int64_t buf_len) {
    if (!r || tg < 0 || tg >= r->n_targets || !db_name || !query_name) { fail(KM_E_ARG, "km_result_format_target: bad argument"); return -1; }
    std::vector<char> arena, text;
    std::vector<FmtRow> rows;
    format_rows_of(r, tg, db_name, strlen(db_name), query_name, strlen(query_name), arena, rows, text);
    const int64_t need = (int64_t)text.size();
    if (buf && need < buf_len) { memcpy(buf, text.data(), text.size()); buf[need] = 0; }
    return need;
}

// rows of targets [lo, hi) in order, appended to `out`
static void format_range(const km_result* r, int lo, int hi, const char* db_name, const char* names, const int64_t* name_off,
                         std::vector<char>& out) {
    const size_t db_len = strlen(db_name);
    std::vector<char> arena;
    std::vector<FmtRow> rows;
    size_t guess = 0;
    for (int t = lo; t < hi; ++t) guess += (size_t)r->row_count[t] * (size_t)(2 * (r->seq_off[t + 1] - r->seq_off[t]) + 160 + db_len);
    out.reserve(out.size() + guess);
    for (int t = lo; t < hi; ++t)
        format_rows_of(r, t, db_name, db_len, names + name_off[t], (size_t)(name_off[t + 1] - name_off[t]), arena, rows, out);
}

// Builds (once) the text of all targets in target order on host threads: every thread formats a
// contiguous range of targets into its own buffer, the pieces are then copied side by side.
static int build_text(const km_result* r, const char* db_name, const char* names, const int64_t* name_off, int32_t threads) {
    const int n = r->n_targets;
    std::string key = std::string(db_name) + '\0' + (n ? std::string(names, (size_t)name_off[n]) : std::string());
    if (r->text_len >= 0 && r->fmt_key == key) return 0;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::max(1, n / 64)));
    std::vector<std::vector<char>> piece((size_t)nt);
    auto work = [&](int w) {
        const int lo = (int)((int64_t)n * w / nt), hi = (int)((int64_t)n * (w + 1) / nt);
        format_range(r, lo, hi, db_name, names, name_off, piece[(size_t)w]);
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nt; ++i) pool.emplace_back(work, i);
        for (auto& th : pool) th.join();
    }
    int64_t total = 0;
    std::vector<int64_t> at((size_t)nt);
    for (int i = 0; i < nt; ++i) { at[(size_t)i] = total; total += (int64_t)piece[(size_t)i].size(); }
    r->text.reset((size_t)total + 1);
    char* dst = r->text.get();
    auto copy = [&](int w) { if (!piece[(size_t)w].empty()) memcpy(dst + at[(size_t)w], piece[(size_t)w].data(), piece[(size_t)w].size()); };
    if (nt == 1) copy(0);
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nt; ++i) pool.emplace_back(copy, i);
        for (auto& th : pool) th.join();
    }
    dst[total] = 0;
    r->text_len = total;
    r->fmt_key.swap(key);
    return 0;
}

extern "C" int64_t km_result_format_all(const km_result* r, const char* db_name, const char* names, const int64_t* name_off,
                                        int32_t threads, char* buf, int64_t buf_len) {
    if (!r || !db_name || (r->n_targets && (!names || !name_off))) { fail(KM_E_ARG, "km_result_format_all: bad argument"); return -1; }
    build_text(r, db_name, names, name_off, threads);
    const int64_t need = r->text_len;
    if (buf && need < buf_len) memcpy(buf, r->text.get(), (size_t)need + 1);
    return need;
}

extern "C" int64_t km_result_text(const km_result* r, const char* db_name, const char* names, const int64_t* name_off,
                                  int32_t threads, const char** text) {
    if (!r || !db_name || !text || (r->n_targets && (!names || !name_off))) { fail(KM_E_ARG, "km_result_text: bad argument"); return -1; }
    build_text(r, db_name, names, name_off, threads);
    *text = r->text.get();
    return r->text_len;
}

// ---- pipelined batch -> text ---------------------------------------------------------------------
// A small persistent pool of host threads (thread creation costs more than formatting a sub-batch).
struct HostPool {
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::vector<std::thread> workers;
    bool stop = false;
    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i)
            workers.emplace_back([this] {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (stop && q.empty()) return;
                        job = std::move(q.front());
                        q.pop_front();
                    }
                    job();
                }
            });
    }
    void submit(std::function<void()> f) { { std::lock_guard<std::mutex> g(m); q.push_back(std::move(f)); } cv.notify_one(); }
    ~HostPool() { { std::lock_guard<std::mutex> g(m); stop = true; } cv.notify_all(); for (auto& w : workers) w.join(); }
};
static HostPool& host_pool() {
    static HostPool pool((int)std::max(2u, std::min(64u, std::thread::hardware_concurrency())));
    return pool;
}
struct Latch {
    std::mutex m; std::condition_variable cv; int left;
    explicit Latch(int n) : left(n) {}
    void done() { std::lock_guard<std::mutex> g(m); if (--left == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return left == 0; }); }
};

// km find_mutation for a whole batch, host buffers in, text out, as ONE call: the batch is cut into
// sub-batches that are all enqueued at once on their own streams; while the GPU works on the later
// ones the host formats the rows of the earlier ones (pool threads), so copies, kernels and text
// building overlap.  The text equals km_find_batch + km_result_format_all.
// KM_TRACE=1: host-clock timeline of km_find_text on stderr (measurement aid)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::mutex m;
    std::vector<std::tuple<const char*, int, double>> ev;
    Trace() : on(getenv("KM_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what, int sub = -1) {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::lock_guard<std::mutex> g(m);
        ev.emplace_back(what, sub, ms);
    }
    ~Trace() {
        if (!on) return;
        for (auto& e : ev) fprintf(stderr, "[km_trace] %8.3f ms  %s %d\n", std::get<2>(e), std::get<0>(e), std::get<1>(e));
    }
};

extern "C" int km_find_text(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params,
                            const char* db_name, const char* names, const int64_t* name_off, int32_t n_sub, km_result** out) {
    if (!t || !out || n < 0 || (n && (!seqs || !offsets || !names || !name_off)) || !params || !db_name)
        return fail(KM_E_ARG, "km_find_text: bad argument");
    CU(cudaSetDevice(t->device));
    Trace tr;
    if (n_sub <= 0) n_sub = n >= 4096 ? 6 : n >= 1024 ? 2 : 1;
    n_sub = std::max(1, std::min(n_sub, std::max(1, n)));
    while ((int)t->lanes.size() < n_sub) {
        std::unique_ptr<km_table::Lane> L(new km_table::Lane());
        L->pin.host = true;
        // earlier sub-batches run at higher stream priority: they finish one after the other instead of all
        // together at the end, so the host can format the first while the GPU works on the rest
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));          // lo = least urgent (0), hi = most urgent (negative)
        const int prio = std::min(prio_lo, prio_hi + (int)t->lanes.size());
        CU(cudaStreamCreateWithPriority(&L->stream, cudaStreamNonBlocking, prio));
        CU(cudaStreamCreateWithPriority(&L->side, cudaStreamNonBlocking, prio));
        for (auto& e : L->ev) CU(cudaEventCreate(&e));
        CU(cudaEventCreateWithFlags(&L->fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&L->join, cudaEventDisableTiming));
        t->lanes.push_back(std::move(L));
    }
    // sub-batches balanced by sequence length (contiguous ranges)
    std::vector<int> cut(1, 0);
    const int64_t total = n ? offsets[n] - offsets[0] : 0;
    for (int c = 1; c < n_sub; ++c) {
        const int64_t want = offsets[0] + total * c / n_sub;
        int i = (int)(std::lower_bound(offsets, offsets + n + 1, want) - offsets);
        i = std::max(cut.back(), std::min(i, n));
        cut.push_back(i);
    }
    cut.push_back(n);
    km_result* res = new km_result();
    res->n_targets = n; res->k = t->k; res->has_graph = false;
    km_find_params prm = *params;
    prm.flags |= KM_FIND_NO_GRAPH;
    if (!getenv("KM_HOST_FORMAT")) {
        // ---- the text is formatted on the device (format.h) --------------------------------------------------
        // One pool task per sub-batch: set-up + upload + launches, a small fetch (per-target status, length of
        // the text), then -- once the lengths of the sub-batches before it are known -- ONE copy of its text
        // straight to its place in the result's pinned buffer.  The host formats nothing and joins nothing.
        const std::string db(db_name);
        std::vector<int64_t> caps((size_t)n_sub);
        int64_t cap_total = 1;
        for (int c = 0; c < n_sub; ++c) {
            const int lo = cut[(size_t)c], hi = cut[(size_t)c + 1];
            const int64_t rows = std::max(64, 16 * (hi - lo));
            caps[(size_t)c] = 16 * (offsets[hi] - offsets[lo]) + 512ll * (hi - lo) + rows * (int64_t)(db.size() + 64) + (1 << 16);
            cap_total += caps[(size_t)c];
        }
        if (int rc = res->text.pin.reserve(t->pool, (size_t)cap_total)) { delete res; return rc; }
        char* final_text = res->text.pin.base;
        std::vector<std::unique_ptr<km_plan>> plans((size_t)n_sub);
        std::vector<std::vector<int64_t>> offs((size_t)n_sub), noffs((size_t)n_sub);
        std::vector<std::vector<char>> spill((size_t)n_sub);            // text that did not go straight to its place
        std::vector<char> spilled((size_t)n_sub, 0);
        res->parts.resize((size_t)n_sub);
        std::vector<int> rcs((size_t)n_sub, 0);
        std::vector<std::string> errs((size_t)n_sub);
        std::vector<long long> lens((size_t)n_sub, -1);
        std::mutex lm; std::condition_variable lcv;
        Latch latch(n_sub);
        const int device = t->device;
        std::atomic<int> next_lane(0);
        for (int c = 0; c < n_sub; ++c) {
            host_pool().submit([=, &db, &next_lane, &cut, &plans, &offs, &noffs, &spill, &spilled, &rcs, &errs, &lens, &lm, &lcv, &latch, &prm, &tr] {
                const int lo = cut[(size_t)c], hi = cut[(size_t)c + 1];
                auto publish = [&](long long len) { { std::lock_guard<std::mutex> g(lm); lens[(size_t)c] = len; } lcv.notify_all(); };
                auto fail_all = [&](int rc) { rcs[(size_t)c] = rc; errs[(size_t)c] = g_err; publish(0); latch.done(); };
                if (cudaSetDevice(device) != cudaSuccess) { fail(KM_E_CUDA, "cudaSetDevice failed"); return fail_all(KM_E_CUDA); }
                auto& o = offs[(size_t)c]; auto& no = noffs[(size_t)c];
                o.resize((size_t)(hi - lo) + 1); no.resize((size_t)(hi - lo) + 1);
                for (int i = lo; i <= hi; ++i) { o[(size_t)(i - lo)] = offsets[i] - offsets[lo]; no[(size_t)(i - lo)] = name_off[i] - name_off[lo]; }
                plans[(size_t)c].reset(new km_plan());
                km_plan* p = plans[(size_t)c].get();
                p->fmt = true; p->fmt_names = names + name_off[lo]; p->fmt_name_off = no.data(); p->fmt_db = db;
                p->targets_ext = seqs + offsets[lo];
                const int lane_ix = next_lane.fetch_add(1);
                g_trace_obj = &tr; g_trace_sub = c;
                g_trace_mark = tr.on ? +[](void* o, const char* w, int sub) { static_cast<Trace*>(o)->mark(w, sub); } : nullptr;
                tr.mark("task start", c);
                p->defer_upload = true;
                if (int rc = plan_init(t, seqs + offsets[lo], o.data(), hi - lo, &prm, p, false, t->lanes[(size_t)lane_ix].get())) return fail_all(rc);
                tr.mark("plan_init", c);
                {
                    // ~17 driver calls per sub-batch, ~6 us each whether one thread issues them or six do at once
                    // (measured both ways: taking turns under a mutex put the last sub-batch on the GPU at 0.71 ms
                    // instead of 0.54 and gained nothing for the first)
                    int rc = plan_upload_enqueue(p, p->stream);
                    if (!rc) rc = plan_launch(p, p->stream);
                    if (rc) return fail_all(rc);
                }
                p->defer_upload = false;
                tr.mark("plan_launch", c);
                std::unique_ptr<km_result> part(new km_result());
                part->seq_off = p->seq_off;
                if (int rc = plan_fetch(p, part.get(), false, true)) return fail_all(rc);
                tr.mark("head fetched", c);
                km_result* pr = part.get();
                res->parts[(size_t)c] = std::move(part);
                long long len = pr->dev_text_len;
                const bool host_format = pr->dev_text_flags != 0;
                if (host_format) {        // the device declined (capacity, or a number it does not print): rows come back, host formats
                    if (int rc = plan_download(p, p->stream, pr, false)) return fail_all(rc);
                    pr->targets.assign(p->targets_ext, (size_t)p->n_code);
                    format_range(pr, 0, pr->n_targets, db.c_str(), names + name_off[lo], no.data(), spill[(size_t)c]);
                    len = (long long)spill[(size_t)c].size();
                }
                publish(len);
                long long at = 0;
                {
                    std::unique_lock<std::mutex> lk(lm);
                    lcv.wait(lk, [&] { for (int j = 0; j < c; ++j) if (lens[(size_t)j] < 0) return false; return true; });
                    for (int j = 0; j < c; ++j) at += lens[(size_t)j];
                }
                const bool fits = at + len + 1 <= cap_total;
                if (host_format) {
                    if (fits) { memcpy(final_text + at, spill[(size_t)c].data(), (size_t)len); spill[(size_t)c].clear(); }
                    else spilled[(size_t)c] = 1;
                } else if (len) {
                    char* dst = final_text + at;
                    if (!fits) { spill[(size_t)c].resize((size_t)len); dst = spill[(size_t)c].data(); spilled[(size_t)c] = 1; }
                    if (cudaMemcpyAsync(dst, p->F.text, (size_t)len, cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
                        cudaStreamSynchronize(p->stream) != cudaSuccess) {
                        fail(KM_E_CUDA, "copy of the text failed: %s", cudaGetErrorString(cudaGetLastError()));
                        rcs[(size_t)c] = KM_E_CUDA; errs[(size_t)c] = g_err;
                    }
                    pr->bytes_d2h += (unsigned long long)len;
                }
                tr.mark("text placed", c);
                plan_swap_vecs(p, t->lanes[(size_t)lane_ix]->vecs);
                latch.done();
            });
        }
        tr.mark("all submitted");
        latch.wait();
        tr.mark("all placed");
        for (int c = 0; c < n_sub; ++c)
            if (rcs[(size_t)c]) {
                const int rc = rcs[(size_t)c];
                fail(rc, "%s", errs[(size_t)c].c_str());
                delete res;
                return rc;
            }
        long long len = 0;
        bool any_spill = false;
        for (int c = 0; c < n_sub; ++c) { len += lens[(size_t)c]; any_spill |= spilled[(size_t)c] != 0; }
        if (any_spill) {
            // (only after capacity retries grew a sub-batch beyond the estimate) assemble in a buffer of the exact size
            PinBlock exact;
            if (int rc = exact.reserve(t->pool, (size_t)len + 1)) { delete res; return rc; }
            long long at = 0;
            for (int c = 0; c < n_sub; ++c) {
                const long long l = lens[(size_t)c];
                if (spilled[(size_t)c]) memcpy(exact.base + at, spill[(size_t)c].data(), (size_t)l);
                else if (at + l + 1 <= cap_total) memcpy(exact.base + at, final_text + at, (size_t)l);
                at += l;
            }
            res->text.pin.drop();
            res->text.pin.pool = exact.pool; res->text.pin.base = exact.base; res->text.pin.cap = exact.cap;
            exact.base = nullptr;
        }
        res->text.get()[len] = 0;
        res->text_len = len;
        for (auto& part : res->parts) {
            res->all_status.insert(res->all_status.end(), part->status.data(), part->status.data() + part->status.size());
            res->ms_h2d += part->ms_h2d; res->ms_walk += part->ms_walk; res->ms_graph += part->ms_graph; res->ms_d2h += part->ms_d2h;
            res->n_launches += part->n_launches; res->n_retries += part->n_retries;
            res->bytes_h2d += part->bytes_h2d; res->bytes_d2h += part->bytes_d2h;
        }
        res->fmt_key = std::string(db_name) + '\0' + (n ? std::string(names, (size_t)name_off[n]) : std::string());
        *out = res;
        tr.mark("done");
        return 0;
    }
    // ---- KM_HOST_FORMAT: rows come back, host threads format them -------------------------------------------
    // One pool task per sub-batch: layout + upload + launch, then the fetch (which waits for that sub-batch's
    // stream only), then its rows go to the pool in slices.  Nothing on the pool waits for another pool task; the
    // caller waits for the last slice.  Enqueueing from several threads at once keeps the host off the critical
    // path: done one after the other the six set-ups alone took as long as all the kernels.
    const int n_slice = std::max(4, std::min(16, (int)host_pool().workers.size() / 2));
    std::vector<std::vector<char>> piece((size_t)n_sub * n_slice);
    std::vector<std::unique_ptr<km_plan>> plans((size_t)n_sub);
    std::vector<std::vector<int64_t>> offs((size_t)n_sub), noffs((size_t)n_sub);
    res->parts.resize((size_t)n_sub);
    std::vector<int> rcs((size_t)n_sub, 0);
    std::vector<std::string> errs((size_t)n_sub);
    Latch latch(n_sub * n_slice);
    const int device = t->device;
    std::atomic<int> next_lane(0);
    for (int c = 0; c < n_sub; ++c) {
        host_pool().submit([=, &next_lane, &cut, &piece, &plans, &offs, &noffs, &rcs, &errs, &latch, &prm, &tr] {
            const int lo = cut[(size_t)c], hi = cut[(size_t)c + 1];
            auto fail_all = [&](int rc) { rcs[(size_t)c] = rc; errs[(size_t)c] = g_err; for (int j = 0; j < n_slice; ++j) latch.done(); };
            if (cudaSetDevice(device) != cudaSuccess) { fail(KM_E_CUDA, "cudaSetDevice failed"); return fail_all(KM_E_CUDA); }
            auto& o = offs[(size_t)c]; auto& no = noffs[(size_t)c];
            o.resize((size_t)(hi - lo) + 1); no.resize((size_t)(hi - lo) + 1);
            for (int i = lo; i <= hi; ++i) { o[(size_t)(i - lo)] = offsets[i] - offsets[lo]; no[(size_t)(i - lo)] = name_off[i] - name_off[lo]; }
            plans[(size_t)c].reset(new km_plan());
            km_plan* p = plans[(size_t)c].get();
            tr.mark("task start", c);
            // lanes are handed out in the order the tasks get here: the first one to enqueue has the most urgent streams
            const int lane_ix = next_lane.fetch_add(1);
            if (int rc = plan_init(t, seqs + offsets[lo], o.data(), hi - lo, &prm, p, false, t->lanes[(size_t)lane_ix].get())) return fail_all(rc);
            tr.mark("plan_init", c);
            if (int rc = plan_launch(p, p->stream)) return fail_all(rc);
            tr.mark("plan_launch", c);
            std::unique_ptr<km_result> part(new km_result());
            part->seq_off = p->seq_off;
            if (int rc = plan_fetch(p, part.get(), false)) return fail_all(rc);
            tr.mark("plan_fetch", c);
            part->targets.swap(p->targets);
            km_result* pr = part.get();
            res->parts[(size_t)c] = std::move(part);
            const char* nm = names + name_off[lo];
            const int64_t* nop = no.data();
            const int m = pr->n_targets;
            for (int j = 0; j < n_slice; ++j) {
                std::vector<char>* dst = &piece[(size_t)c * n_slice + (size_t)j];
                const int a = (int)((int64_t)m * j / n_slice), b = (int)((int64_t)m * (j + 1) / n_slice);
                host_pool().submit([pr, db_name, nm, nop, a, b, dst, &latch, &tr, c] {
                    *dst = piece_cache().get(0);
                    format_range(pr, a, b, db_name, nm, nop, *dst);
                    tr.mark("slice", c);
                    latch.done();
                });
            }
        });
    }
    tr.mark("all submitted");
    latch.wait();
    tr.mark("formatted");
    for (int c = 0; c < n_sub; ++c)
        if (rcs[(size_t)c]) {
            const int rc = rcs[(size_t)c];
            fail(rc, "%s", errs[(size_t)c].c_str());
            for (auto& pc : piece) piece_cache().put(std::move(pc));
            delete res;
            return rc;
        }
    int64_t len = 0;
    std::vector<int64_t> at_of;
    for (auto& pc : piece) { at_of.push_back(len); len += (int64_t)pc.size(); }
    res->text.reset((size_t)len + 1);
    {
        Latch joined((int)piece.size());
        char* dst = res->text.get();
        for (size_t c = 0; c < piece.size(); ++c) {
            std::vector<char>* pc = &piece[c];
            const int64_t at = at_of[c];
            host_pool().submit([pc, dst, at, &joined] {
                if (!pc->empty()) memcpy(dst + at, pc->data(), pc->size());
                piece_cache().put(std::move(*pc));
                joined.done();
            });
        }
        joined.wait();
    }
    tr.mark("joined");
    for (auto& part : res->parts) {
        res->all_status.insert(res->all_status.end(), part->status.data(), part->status.data() + part->status.size());
        res->ms_h2d += part->ms_h2d; res->ms_walk += part->ms_walk; res->ms_graph += part->ms_graph; res->ms_d2h += part->ms_d2h;
        res->n_launches += part->n_launches; res->n_retries += part->n_retries;
        res->bytes_h2d += part->bytes_h2d; res->bytes_d2h += part->bytes_d2h;
    }
    res->text.get()[len] = 0;
    res->text_len = len;
    res->fmt_key = std::string(db_name) + '\0' + (n ? std::string(names, (size_t)name_off[n]) : std::string());
    *out = res;
    tr.mark("done");
    return 0;
}

extern "C" int km_debug_format_fixed(double v, int prec, char* buf64) {
    if (!buf64 || prec < 0 || prec > 3) return fail(KM_E_ARG, "km_debug_format_fixed: bad argument");
    char* e = put_fixed(buf64, v, prec);
    *e = 0;
    return (int)(e - buf64);
}
extern "C" int km_debug_nat_cmp(const char* a, const char* b) { return nat_cmp(a, strlen(a), b, strlen(b)); }

// ---- measurement helpers ---------------------------------------------------------------------------
extern "C" int km_debug_phase_cycles(unsigned long long* out32, int reset) {
#ifdef KM_PHASE_TIMERS
    if (out32) CU(cudaMemcpyFromSymbol(out32, km_phase_cycles, 64 * sizeof(unsigned long long)));
    if (reset) { unsigned long long z[64] = {0}; CU(cudaMemcpyToSymbol(km_phase_cycles, z, sizeof(z))); }
    return 0;
#else
    (void)out32; (void)reset;
    return fail(KM_E_ARG, "library built without KM_PHASE_TIMERS");
#endif
}

extern "C" int km_debug_target_cycles(unsigned int* out, int n) {
#ifdef KM_PHASE_TIMERS
    if (!out || n < 0 || n > KM_DEBUG_TARGETS) return fail(KM_E_ARG, "km_debug_target_cycles: bad argument");
    CU(cudaMemcpyFromSymbol(out, km_target_cycles, (size_t)n * sizeof(unsigned int)));
    return 0;
#else
    (void)out; (void)n;
    return fail(KM_E_ARG, "library built without KM_PHASE_TIMERS");
#endif
}

extern "C" int km_bench_random_gather(int device, uint64_t bytes, uint64_t n_loads, int iters, float* best_ms) {
    if (!best_ms || bytes < 64 || iters < 1) return fail(KM_E_ARG, "km_bench_random_gather: bad argument");
    if (km_device_count() <= 0) return fail(KM_E_NOGPU, "no CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    void* buf = nullptr;
    uint32_t* sink = nullptr;
    CU(cudaMalloc(&buf, bytes));
    CU(cudaMalloc((void**)&sink, 4));
    CU(cudaMemset(buf, 0x5A, bytes));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    float best = 1e30f;
    for (int it = 0; it < iters + 1; ++it) {
        CU(cudaEventRecord(a));
        km_gather_kernel<<<prop.multiProcessorCount * 8, 256>>>((const uint4*)buf, bytes / 32, n_loads, 0x1234 + it, sink);
        CU(cudaEventRecord(b));
        CU(cudaEventSynchronize(b));
        float ms; CU(cudaEventElapsedTime(&ms, a, b));
        if (it > 0 && ms < best) best = ms;     // first pass is warm-up
    }
    CU(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(buf); cudaFree(sink);
    *best_ms = best;
    return 0;
}

extern "C" int km_bench_lookup(km_table* t, uint64_t table_seed, uint64_t table_n, uint64_t n_queries, uint64_t query_seed, int iters,
                               float* best_ms, float* mean_ms, uint64_t* n_hits) {
    if (!t || !n_queries || iters < 1) return fail(KM_E_ARG, "km_bench_lookup: bad argument");
    CU(cudaSetDevice(t->device));
    uint64_t* dq = nullptr; uint32_t* dc = nullptr;
    CU(cudaMalloc((void**)&dq, n_queries * 8));
    CU(cudaMalloc((void**)&dc, n_queries * 4));
    cudaStream_t s = t->stream;
    km_make_queries_kernel<<<t->sm_count * 8, 256, 0, s>>>(dq, n_queries, table_seed, table_n, query_seed, t->k);
    CU(cudaGetLastError());
    float best = 1e30f, sum = 0;
    for (int it = 0; it < iters + 3; ++it) {       // 3 warm-up passes
        CU(cudaEventRecord(t->ev[0], s));
        if (int rc = km_query_batch_device(t, dq, n_queries, dc, s)) return rc;
        CU(cudaEventRecord(t->ev[1], s));
        CU(cudaEventSynchronize(t->ev[1]));
        float ms; CU(cudaEventElapsedTime(&ms, t->ev[0], t->ev[1]));
        if (it >= 3) { best = std::min(best, ms); sum += ms; }
    }
    CU(cudaMemsetAsync(t->d_counter, 0, 16, s));
    km_count_nonzero_kernel<<<t->sm_count * 8, 256, 0, s>>>(dc, n_queries, t->d_counter);
    unsigned long long hits = 0;
    CU(cudaMemcpyAsync(&hits, t->d_counter, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    cudaFree(dq); cudaFree(dc);
    if (best_ms) *best_ms = best;
    if (mean_ms) *mean_ms = sum / iters;
    if (n_hits) *n_hits = hits;
    return 0;
}
