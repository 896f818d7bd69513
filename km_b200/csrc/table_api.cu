// C ABI of libkm_b200.so, part 1 (see include/km_b200.h): the k-mer count table -- creation, cohort shards
// (CUDA virtual-memory API), inserts, counting from reads, filtering, export and the batched lookups.  Host
// orchestration only: the arithmetic runs in the kernels of table_kernels.cuh.  There is no CPU fallback --
// without a CUDA device every entry point fails with KM_E_NOGPU.
#include "host_common.h"
#include "table_kernels.cuh"
#include "find_launch.h"

static_assert(sizeof(Bucket) == 32, "bucket must be one 32-byte sector");

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* km_last_error(void) { return g_err; }
extern "C" const char* km_version(void) { return "km_b200 0.1 (sm_100a)"; }
extern "C" int km_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static bool default_lines() { const char* e = getenv("KM_TABLE_LINES"); return e && *e && *e != '0'; }
// units (sector buckets or family lines) for `capacity_keys` keys: buckets hold 2 records at load <= 0.5;
// lines hold 8 slots, every key takes two of them, load 0.625
static uint64_t units_for(uint64_t capacity_keys, int lines) {
    return std::max<uint64_t>(64, lines ? (capacity_keys * 2 + 4) / 5 : capacity_keys);
}
static void clear_units(km_table* t, void* mem, uint64_t n);

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
    CU(cudaStreamCreateWithFlags(&t->side2, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&t->side3, cudaStreamNonBlocking));
    for (auto& ev : t->ev) CU(cudaEventCreate(&ev));
    CU(cudaEventCreateWithFlags(&t->fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&t->join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&t->join2, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&t->join3, cudaEventDisableTiming));
    CU(cudaMalloc((void**)&t->d_counter, 16));
    t->pin.host = true;
    t->pin_find.host = true;
    CU(km_find_kernels_init());
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

// the same on the device, and the two-pass partition of a query batch by owner (explicit all-to-all exchange,
// km_b200/cohort.py query_routed_device): nothing but the per-owner counts ever visits the host
extern "C" int km_shard_owner_device(km_table* t, const uint64_t* kmers_dev, uint64_t n, int32_t* owner_dev, void* stream) {
    if (!t || (n && (!kmers_dev || !owner_dev))) return fail(KM_E_ARG, "km_shard_owner_device: bad argument");
    if (!n) return 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : t->stream;
    km_owner_kernel<<<grid_for(t, n, 256, 8), 256, 0, s>>>(t->view(), kmers_dev, n, owner_dev);
    CU(cudaGetLastError());
    return 0;
}

// counts_dev: KM_MAX_SHARDS * 3 uint64 of device scratch; on return (stream order) [0..8) holds the number of queries
// per owner, sorted_dev the k-mers grouped by owner (owner 0 first), perm_dev[i] the position sorted_dev[i] came from
extern "C" int km_route_partition(km_table* t, const uint64_t* kmers_dev, uint64_t n, uint64_t* sorted_dev, uint32_t* perm_dev,
                                  uint64_t* counts_dev, void* stream) {
    if (!t || !counts_dev || (n && (!kmers_dev || !sorted_dev || !perm_dev)) || n >= (1ull << 32))
        return fail(KM_E_ARG, "km_route_partition: bad argument");
    cudaStream_t s = stream ? (cudaStream_t)stream : t->stream;
    unsigned long long* c = reinterpret_cast<unsigned long long*>(counts_dev);
    CU(cudaMemsetAsync(c, 0, sizeof(unsigned long long) * 3 * KM_MAX_SHARDS, s));
    if (!n) return 0;
    const int grid = grid_for(t, n, 256, 8);
    km_route_hist_kernel<<<grid, 256, 0, s>>>(t->view(), kmers_dev, n, c);
    CU(cudaGetLastError());
    km_route_prefix_kernel<<<1, 32, 0, s>>>(c);
    CU(cudaGetLastError());
    km_route_scatter_kernel<<<grid, 256, 0, s>>>(t->view(), kmers_dev, n, c + KM_MAX_SHARDS, c + 2 * KM_MAX_SHARDS, sorted_dev, perm_dev);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int km_route_unpermute(km_table* t, const uint32_t* answers_dev, const uint32_t* perm_dev, uint64_t n, uint32_t* out_dev,
                                  void* stream) {
    if (!t || (n && (!answers_dev || !perm_dev || !out_dev))) return fail(KM_E_ARG, "km_route_unpermute: bad argument");
    if (!n) return 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : t->stream;
    km_route_unpermute_kernel<<<grid_for(t, n, 256, 8), 256, 0, s>>>(answers_dev, perm_dev, n, out_dev);
    CU(cudaGetLastError());
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
        if (L->side2) cudaStreamDestroy(L->side2);
        if (L->side3) cudaStreamDestroy(L->side3);
        if (L->fork) cudaEventDestroy(L->fork);
        if (L->join) cudaEventDestroy(L->join);
        if (L->join2) cudaEventDestroy(L->join2);
        if (L->join3) cudaEventDestroy(L->join3);
        if (L->gexec) cudaGraphExecDestroy(L->gexec);
        if (L->wait_ev) cudaEventDestroy(L->wait_ev);
    }
    for (auto& ev : t->ev) if (ev) cudaEventDestroy(ev);
    if (t->stream) cudaStreamDestroy(t->stream);
    if (t->side) cudaStreamDestroy(t->side);
    if (t->side2) cudaStreamDestroy(t->side2);
    if (t->side3) cudaStreamDestroy(t->side3);
    if (t->fork) cudaEventDestroy(t->fork);
    if (t->join) cudaEventDestroy(t->join);
    if (t->join2) cudaEventDestroy(t->join2);
    if (t->join3) cudaEventDestroy(t->join3);
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
    t->linked = false;                     // the content changed: the neighbour masks are stale until km_table_link
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

// ---- counting: a byte stream of sequences into the table ------------------------------------------------------
int CountStream::open(km_table* table, size_t cap_bytes, bool with_qual) {
    t = table; want_qual = with_qual;
    cap = align_up_sz(std::max<size_t>(cap_bytes, 4096), 256);
    CU(cudaSetDevice(t->device));
    for (int b = 0; b < 2; ++b) {
        CU(cudaMallocHost((void**)&pin_seq[b], cap));
        CU(cudaMalloc((void**)&dev_seq[b], cap + 256));
        if (want_qual) {
            CU(cudaMallocHost((void**)&pin_q[b], cap));
            CU(cudaMalloc((void**)&dev_q[b], cap + 256));
        }
        CU(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    }
    CU(cudaMemsetAsync(t->d_counter, 0, 16, t->stream));
    return 0;
}
int CountStream::submit(size_t n_bytes, int min_qual) {
    if (n_bytes > cap) return fail(KM_E_ARG, "internal: count chunk of %zu bytes exceeds the staging buffer", n_bytes);
    if (n_bytes) {
        const bool q = want_qual && min_qual > 0;
        CU(cudaMemcpyAsync(dev_seq[slot], pin_seq[slot], n_bytes, cudaMemcpyHostToDevice, t->stream));
        if (q) CU(cudaMemcpyAsync(dev_q[slot], pin_q[slot], n_bytes, cudaMemcpyHostToDevice, t->stream));
        const uint64_t n_tiles = (n_bytes + KM_COUNT_TILE - 1) / KM_COUNT_TILE;
        const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)t->sm_count * 8);
        km_count_text_kernel<<<grid, KM_COUNT_CTA, 0, t->stream>>>(t->view(), (const uint32_t*)dev_seq[slot], q ? (const uint32_t*)dev_q[slot] : nullptr,
                                                                  min_qual, n_bytes, t->d_counter, reinterpret_cast<uint32_t*>(t->d_counter + 1));
        CU(cudaGetLastError());
        CU(cudaEventRecord(done[slot], t->stream));
        busy[slot] = true;
        bytes_in += n_bytes;
    }
    slot ^= 1;
    if (busy[slot]) { CU(cudaEventSynchronize(done[slot])); busy[slot] = false; }
    return 0;
}
int CountStream::close() {
    if (!t) return 0;
    km_table* table = t;
    const int rc = finish_insert(table, "counting k-mers");      // synchronises the stream; with routing on, the number of new
    for (int b = 0; b < 2; ++b) {                                // keys it adds is what THIS rank created anywhere: km_table_recount
        if (pin_seq[b]) cudaFreeHost(pin_seq[b]);
        if (pin_q[b]) cudaFreeHost(pin_q[b]);
        if (dev_seq[b]) cudaFree(dev_seq[b]);
        if (dev_q[b]) cudaFree(dev_q[b]);
        if (done[b]) cudaEventDestroy(done[b]);
        pin_seq[b] = pin_q[b] = dev_seq[b] = dev_q[b] = nullptr; done[b] = nullptr;
    }
    t = nullptr;
    return rc;
}
CountStream::~CountStream() { if (t) { cudaSetDevice(t->device); cudaStreamSynchronize(t->stream); close(); } }

#define KM_COUNT_CHUNK ((size_t)32 << 20)

extern "C" int km_table_count_text(km_table* t, const char* text, const char* qual, uint64_t n_bytes, int min_qual_char) {
    if (!t || (n_bytes && !text)) return fail(KM_E_ARG, "km_table_count_text: bad argument");
    if (t->lines) return fail(KM_E_ARG, "km_table_count_text: counting needs the sector-bucket layout");
    if (!n_bytes) return 0;
    const bool q = qual && min_qual_char > 0;
    CountStream cs;
    if (int rc = cs.open(t, std::min<size_t>(KM_COUNT_CHUNK, n_bytes + 64), q)) return rc;
    // chunks overlap by k - 1 bytes: a k-mer that straddles the cut is counted by the chunk it STARTS in, which
    // therefore sees k - 1 bytes more than it owns; the kernel is told how many start positions are its own
    const size_t k1 = (size_t)t->k - 1;
    const size_t step = cs.cap - k1;
    for (uint64_t at = 0; at < n_bytes; at += step) {
        const size_t own = (size_t)std::min<uint64_t>(step, n_bytes - at);
        const size_t take = (size_t)std::min<uint64_t>(own + k1, n_bytes - at);
        memcpy(cs.seq(), text + at, take);
        if (q) memcpy(cs.qual(), qual + at, take);
        // bytes past `own` must not START a k-mer here (the next chunk owns them): cut the stream at own + k - 1 and
        // blank the tail's ability to start k-mers by ending the chunk there -- a k-mer starting at p >= own needs the byte
        // at p + k - 1 >= own + k - 1 = take, which is past the end
        if (int rc = cs.submit(take, min_qual_char)) return rc;
    }
    return cs.close();
}

// the same from reads given as one concatenated buffer + offsets (reads are staged with a separator between them)
extern "C" int km_table_count_reads(km_table* t, const char* reads, const int64_t* off, int64_t n_reads) {
    if (!t || !reads || !off || n_reads < 0) return fail(KM_E_ARG, "km_table_count_reads: bad argument");
    if (t->lines) return fail(KM_E_ARG, "km_table_count_reads: counting needs the sector-bucket layout");
    if (n_reads == 0) return 0;
    const uint64_t total = (uint64_t)(off[n_reads] - off[0]) + (uint64_t)n_reads;
    CountStream cs;
    if (int rc = cs.open(t, std::min<size_t>(KM_COUNT_CHUNK, total + 64), false)) return rc;
    size_t fill = 0;
    const size_t k1 = (size_t)t->k - 1;
    for (int64_t r = 0; r < n_reads; ++r) {
        const char* p = reads + off[r];
        size_t len = (size_t)(off[r + 1] - off[r]);
        while (len) {                                             // a read longer than the buffer goes in overlapping pieces
            const size_t room = fill + 1 < cs.cap ? cs.cap - fill - 1 : 0;
            if (len <= room) { memcpy(cs.seq() + fill, p, len); fill += len; len = 0; }
            else if (fill == 0) {                                 // fill the whole buffer, continue k - 1 bytes back
                memcpy(cs.seq(), p, room); fill = room;
                if (int rc = cs.submit(fill, 0)) return rc;
                fill = 0; p += room - k1; len -= room - k1;
            } else { if (int rc = cs.submit(fill, 0)) return rc; fill = 0; }
        }
        cs.seq()[fill++] = '\n';
    }
    if (fill) if (int rc = cs.submit(fill, 0)) return rc;
    return cs.close();
}

// cohort mode: route inserts to the owner shard (every peer attached); 0 switches back
extern "C" int km_table_set_routing(km_table* t, int on) {
    if (!t) return fail(KM_E_ARG, "null table");
    if (on) {
        if (t->lines) return fail(KM_E_ARG, "km_table_set_routing: shards use the sector-bucket layout");
        for (int r = 0; r < t->n_shards; ++r)
            if (r != t->my_shard && !t->peer[r]) return fail(KM_E_ARG, "km_table_set_routing: shard %d is not attached", r);
    }
    t->route = on ? 1 : 0;
    return 0;
}

// The neighbour masks (table.h): which successors / predecessors of every stored k-mer are in the table too.  The
// find_mutation kernels use them to skip the lookups of absent successors; they are rebuilt here, on demand, after any
// change of the table's content (one shard, sector layout only -- a shard of a cohort table answers without them).
extern "C" int km_table_link(km_table* t) {
    if (!t) return fail(KM_E_ARG, "null table");
    if (t->lines || t->n_shards > 1) { t->linked = false; return 0; }
    if (t->linked) return 0;
    CU(cudaSetDevice(t->device));
    km_table_link_kernel<<<t->sm_count * 8, 256, 0, t->stream>>>(t->view());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(t->stream));
    t->linked = true;
    return 0;
}
cudaError_t km_wait_stream(cudaStream_t s, cudaEvent_t* blocking_ev) {
    static const int mode = [] {
        const char* e = getenv("KM_BLOCKING_SYNC");
        if (e && *e) return *e != '0' ? 1 : 0;
        return 0;          // measured at 8 ranks on one box: sleeping waits are SLOWER (e2e 2.50 vs 2.13 ms, profiles/r2n_blocking_n8.txt)
    }();
    if (!mode || !blocking_ev) return cudaStreamSynchronize(s);
    if (!*blocking_ev) {
        const cudaError_t e = cudaEventCreateWithFlags(blocking_ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(*blocking_ev, s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(*blocking_ev);
}

int km_ensure_linked(km_table* t) {
    const char* e = getenv("KM_NO_LINKS");            // A/B switch: every successor is looked up, as before
    if (e && *e && *e != '0') { t->linked = false; return 0; }
    return km_table_link(t);
}

extern "C" int km_table_recount(km_table* t, uint64_t* n_keys) {
    if (!t) return fail(KM_E_ARG, "null table");
    if (t->lines) return fail(KM_E_ARG, "km_table_recount: sector-bucket layout only");
    CU(cudaSetDevice(t->device));
    CU(cudaMemsetAsync(t->d_counter, 0, 16, t->stream));
    km_table_recount_kernel<<<t->sm_count * 8, 256, 0, t->stream>>>(t->view(), t->d_counter);
    CU(cudaGetLastError());
    t->n_keys = 0;
    if (int rc = finish_insert(t, "km_table_recount")) return rc;
    if (n_keys) *n_keys = t->n_keys;
    return 0;
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

// device-resident counting benchmark: n_reads reads of read_len bases drawn from a pseudo-random genome of `genome` bases
// (coverage = n_reads * read_len / genome) are generated on the device, then counted `iters` times into the table
// (km_count_text_kernel alone, CUDA events); the table is left holding the counts of all passes
extern "C" int km_bench_count(km_table* t, uint64_t n_reads, int read_len, uint64_t genome, uint64_t seed, int iters, float* best_ms,
                              uint64_t* n_kmers) {
    if (!t || !n_reads || read_len < t->k || genome <= (uint64_t)read_len || iters < 1 || !best_ms) return fail(KM_E_ARG, "km_bench_count: bad argument");
    if (t->lines) return fail(KM_E_ARG, "km_bench_count: sector-bucket layout only");
    CU(cudaSetDevice(t->device));
    const uint64_t n_bytes = n_reads * (uint64_t)(read_len + 1);
    uint8_t* text = nullptr;
    CU(cudaMalloc((void**)&text, n_bytes + 256));
    cudaStream_t s = t->stream;
    km_make_reads_kernel<<<t->sm_count * 8, 256, 0, s>>>(text, n_reads, read_len, genome, seed);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(t->d_counter, 0, 16, s));
    const uint64_t n_tiles = (n_bytes + KM_COUNT_TILE - 1) / KM_COUNT_TILE;
    const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)t->sm_count * 8);
    float best = 1e30f;
    for (int it = 0; it < iters + 1; ++it) {                    // the first pass creates the keys, the timed ones add to them
        CU(cudaEventRecord(t->ev[0], s));
        km_count_text_kernel<<<grid, KM_COUNT_CTA, 0, s>>>(t->view(), (const uint32_t*)text, nullptr, 0, n_bytes, t->d_counter,
                                                          reinterpret_cast<uint32_t*>(t->d_counter + 1));
        CU(cudaGetLastError());
        CU(cudaEventRecord(t->ev[1], s));
        CU(cudaEventSynchronize(t->ev[1]));
        float ms; CU(cudaEventElapsedTime(&ms, t->ev[0], t->ev[1]));
        if (it == 0 && best_ms) best_ms[1] = ms;                // [1] = the pass that inserts new keys
        if (it > 0) best = std::min(best, ms);
    }
    cudaFree(text);
    best_ms[0] = best;
    if (n_kmers) *n_kmers = n_reads * (uint64_t)(read_len - t->k + 1);
    return finish_insert(t, "km_bench_count");
}

// the config-4 lookup mix (50 % background keys on a random strand / 50 % random k-mers) into a caller's device buffer
extern "C" int km_bench_make_queries(km_table* t, uint64_t* queries_dev, uint64_t n, uint64_t table_seed, uint64_t table_n,
                                     uint64_t query_seed, void* stream) {
    if (!t || (n && !queries_dev)) return fail(KM_E_ARG, "km_bench_make_queries: bad argument");
    if (!n) return 0;
    CU(cudaSetDevice(t->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : t->stream;
    km_make_queries_kernel<<<t->sm_count * 8, 256, 0, s>>>(queries_dev, n, table_seed, table_n, query_seed, t->k);
    CU(cudaGetLastError());
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

