// A batch of targets laid out in HBM (km_plan) and what comes back (km_result): shared by plan_api.cu, which
// owns the layout / launch / fetch logic, and text_api.cu, which formats and pipelines.
#pragma once
#include "host_common.h"
#include "find_config.h"
#include "format.h"

static_assert(sizeof(km_row) == sizeof(Row), "km_row must mirror km::Row");

// KM_TRACE: km_find_text's timeline (struct Trace in text_api.cu) is reachable from the plan functions through this hook
extern thread_local void (*g_trace_mark)(void*, const char*, int);
extern thread_local void* g_trace_obj;
extern thread_local int g_trace_sub;
static inline void trace_here(const char* what) { if (g_trace_mark) g_trace_mark(g_trace_obj, what, g_trace_sub); }

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
inline VecCache& piece_cache() { static VecCache c(128); return c; }
inline VecCache& text_cache() { static VecCache c(8); return c; }
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

static inline uint32_t pow2_at_least(uint64_t x) { uint32_t p = 64; while (p < x) p <<= 1; return p; }

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
    std::vector<uint8_t> gave_up;     // 1: the target explored more nodes than extra_max allows; it keeps KM_ST_NODE_OVERFLOW
    int64_t pool_cap = 0, seq_cap = 0, n_node = 0, n_hash = 0, n_code = 0;
    int32_t path_cap = 0, row_cap = 0, extra_max = 0;
    int grid_tiny = 1, grid_graph = 1, grid_large = 1, grid_bubble_tiny = 1, grid_bubble_small = 1;
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
    cudaStream_t stream = nullptr, side = nullptr, side2 = nullptr, side3 = nullptr;      // the table's own unless the plan runs on a lane
    cudaEvent_t* ev = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr, join3 = nullptr;
    cudaEvent_t* wait_ev = nullptr;      // the lane's blocking-sync event (km_wait_stream), or null: spin
    // a plan made by km_find_plan_create owns its side stream and events, so that several plans of one table can be in
    // flight at once on different streams (bench.py launches the panel as parts that overlap each other's phases)
    cudaStream_t own_side = nullptr, own_side2 = nullptr, own_side3 = nullptr;
    cudaEvent_t own_ev[8] = {}, own_fork = nullptr, own_join = nullptr, own_join2 = nullptr, own_join3 = nullptr;
    // a resident plan (km_find_plan_create) launched again and again: the launch sequence as a CUDA graph (plan_api.cu)
    cudaGraphExec_t gexec = nullptr;
    std::string gkey;
    int direct_launches = 0;
    bool capturing = false;              // plan_launch is being captured: phase events are recorded as external event nodes
    size_t clear_bytes = 0;              // state_bytes + what is cleared with the state but not fetched (the scheduler's codes)
    bool trace_events = false;           // KM_TRACE=2: per-phase events also on the km_find_text path (device timeline)
    bool layout_reusable = false;        // plan_init: the borrowed vectors already hold this batch's layout
    int maxcap = 1;
};


int plan_layout(km_plan* p);
int plan_stage(km_plan* p, cudaStream_t s);
int plan_upload_enqueue(km_plan* p, cudaStream_t s);
int plan_upload(km_plan* p, cudaStream_t s);
int plan_launch(km_plan* p, cudaStream_t s);
int plan_download(km_plan* p, cudaStream_t s, km_result* res, bool want_graph, bool head_only = false);
void plan_swap_vecs(km_plan* p, km_table::PlanVecs& v);
void plan_return_vecs(km_plan* p, km_table::Lane* lane);
int plan_init(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params, km_plan* p,
              bool borrow_arena, km_table::Lane* lane = nullptr);
int plan_fetch(km_plan* p, km_result* res, bool want_graph, bool head_only = false);
// everything the enqueue of a staged plan depends on (pointers, sizes, grids, parameters), as bytes: two plans with equal
// keys enqueue identical work, whatever the letters of their targets
std::string plan_graph_key(const km_plan* p);
