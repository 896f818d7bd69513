// Kernels of stages 2 and 3 of find_mutation (sm_100a): the scheduler that sorts targets into work lists and the
// per-target graph / paths / FP64 quantification pass (graph.h, quant.h) -- MutationFinder.graph_analysis,
// quantify_paths, quantify_clusters (km/utils/MutationFinder.py:496-811), Graph.py, PathQuant.py.
#include <cuda_runtime.h>

#include <cstdlib>

#include "graph_bubble.h"
#include "find_config.h"
#include "find_launch.h"
#include "../../include/km_b200.h"

namespace km {

// (measurement builds) wall-clock extent of every kernel of the graph phase: [id][0] = earliest CTA start, [1] = latest CTA
// end, in %globaltimer ns; ids: 0 scheduler, 1 / 2 / 3 CTA-per-target 256 / 512 / general, 4 / 5 bubbles 256 / 512;
// [8 + i] = start / end of the i-th target a CTA-per-target pass took (first 8)
#if defined(KM_PHASE_TIMERS) || defined(KM_TIMELINE)
#define KM_HAVE_TIMELINE 1
static __device__ unsigned long long km_timeline[16][2];
static __device__ unsigned int km_timeline_targets;
__device__ __forceinline__ unsigned long long km_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
struct KernelSpan {
    int id;
    __device__ __forceinline__ explicit KernelSpan(int i) : id(i) { if (threadIdx.x == 0) atomicMin(&km_timeline[id][0], km_now_ns()); }
    __device__ __forceinline__ ~KernelSpan() { if (threadIdx.x == 0) atomicMax(&km_timeline[id][1], km_now_ns()); }
};
#else
struct KernelSpan { __device__ __forceinline__ explicit KernelSpan(int) {} };
#endif

// ---- K4 + K5: graph, paths, FP64 quantification; persistent CTAs over targets -----------------
// Three passes share one body.  The two SHARED-MEMORY passes keep the whole per-target working set
// (adjacency, both shortest-path trees, candidate edges, solver matrices) on chip; a target goes to the
// smallest class its graph fits, so the many small graphs run with twice the CTAs per SM of the
// larger ones (the pass is latency-bound: resident CTAs are throughput).  The GENERAL pass uses
// per-CTA scratch in HBM and takes the rest, plus any target a shared-memory pass deferred
// (KM_ST_RETRY_LARGE).

// NODES = node capacity of a shared-memory class, 0 = the general pass
// Work lists of the graph passes: every target whose walk succeeded goes to the smallest size class its graph fits, each
// list ordered by descending node count (64 size bins; a counting sort in one CTA).  Lists 0 / 1: the 256- / 512-node
// classes of the CTA-per-target pass, 2: the general pass, 3 / 4: the same two classes for the bubble pass -- a target
// goes there when its walk never branched (KM_ST_BRANCHED, cleared here), i.e. when it is almost surely a simple bubble.
#define KM_SCHED_CACHE 12288          // targets whose (list, bin) code is kept in shared memory between the two passes
__global__ void __launch_bounds__(1024) km_schedule_kernel(WalkView W, ResultView R, int bubbles) {
    __shared__ int hist[5][64], start[5][64];
    __shared__ uint16_t code_s[KM_SCHED_CACHE];
    KernelSpan span(0);
    const int n = W.n_targets;
    for (int i = threadIdx.x; i < 5 * 64; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    // list | bin << 3, or 0xFFFF for a target without a graph, from the walk's code (walk.h sched_code_of: class, "the walk
    // branched", size bin).  The shared-memory walk leaves the code of every target it finishes; for the others (the general
    // walk, malformed targets) the scheduler reads the target's state itself.
    auto classify = [&](int t) -> uint32_t {
        uint32_t code = W.sched_code[t];
        if (code == 0) {
            const uint32_t st = W.status[t];
            const int cap = (int)(W.node_off[t + 1] - W.node_off[t]);
            const int nn = W.n_nodes[t];
            code = sched_code_of(st, nn < cap ? nn : cap, W.n_kept[t], KM_TINY_NODES, KM_SMALL_NODES) + 1u;
            if (st & KM_ST_BRANCHED) W.status[t] = st & ~KM_ST_BRANCHED;      // the walk's hint is internal: the host never sees it
        }
        code -= 1u;
        if (code == 0xFFFFu) return code;
        const int c = (int)(code & 3u);
        const int lean = bubbles && !(code & 4u) && c < 2 ? 3 : 0;
        return (uint32_t)(c + lean) | (code & ~7u);
    };
    // pass 1: the loads of a thread's targets are independent of one another (the histogram comes after)
#pragma unroll 4
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const uint32_t code = classify(t);
        if (t < KM_SCHED_CACHE) code_s[t] = (uint16_t)code;
        else W.sched_code[t] = (uint16_t)(code == 0xFFFFu ? 0xFFFFu : 0x8000u | code);      // (beyond the cache: final list kept for pass 2)
        if (code == 0xFFFFu) { R.t_n[t] = 0; R.t_n_paths[t] = 0; R.t_path_first[t] = 0; R.t_n_rows[t] = 0; R.t_row_first[t] = 0; }
        else atomicAdd(&hist[code & 7u][code >> 3], 1);
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        int at = 0;
        for (int b = 0; b < 64; ++b) { start[threadIdx.x][b] = at; at += hist[threadIdx.x][b]; }
        R.sched_count[threadIdx.x] = at;
    } else if (threadIdx.x < 8) R.sched_count[threadIdx.x] = 0;      // [5], [6]: cursors of the bubble pass
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        uint32_t code;
        if (t < KM_SCHED_CACHE) code = (uint32_t)code_s[t];
        else { code = (uint32_t)W.sched_code[t]; code = code == 0xFFFFu ? code : code & 0x7FFFu; }
        if (code != 0xFFFFu) R.sched_order[(size_t)(code & 7u) * n + atomicAdd(&start[code & 7u][code >> 3], 1)] = t;
    }
}

// The simple bubbles (graph_bubble.h): a small group of threads per target -- KM_BUBBLE_THREADS = 32: one warp, several
// targets per CTA; 64 / 128: one CTA per target -- takes the targets of its size class from the scheduler's list 3 / 4; what
// turns out not to be a simple bubble goes to the general pass.
template <int NODES>
__global__ void __launch_bounds__(KM_BUBBLE_THREADS == 32 ? 32 * KM_BUBBLE_WARPS : KM_BUBBLE_THREADS,
                                  NODES == KM_SMALL_NODES ? KM_BUBBLE_SMALL_MINB : KM_BUBBLE_TINY_MINB)
km_graph_bubble_kernel(TableView T, WalkView W, ResultView R) {
    extern __shared__ __align__(16) char km_smem[];
    const int cls = NODES == KM_TINY_NODES ? 0 : 1;
    KernelSpan span(4 + cls);
    const int32_t* order = R.sched_order + (size_t)(3 + cls) * W.n_targets;
    const int count = R.sched_count[3 + cls];
#if KM_BUBBLE_THREADS == 32
    BubbleScratch<NODES>& B = reinterpret_cast<BubbleScratch<NODES>*>(km_smem)[threadIdx.x >> 5];
    WarpCtx ctx;
    for (;;) {
        int i = 0;
        if ((threadIdx.x & 31) == 0) i = atomicAdd(&R.sched_count[5 + cls], 1);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= count) break;
        bubble_target<NODES>(ctx, T, W, R, order[i], B);
        __syncwarp();
    }
#else
    BubbleScratch<NODES>& B = *reinterpret_cast<BubbleScratch<NODES>*>(km_smem);
    __shared__ int next_item[2];
    CtaCtx ctx;
    // work items are taken ONE AHEAD, as in km_graph_kernel: while target t is processed the node arrays the walk left for
    // the next one (evicted from L2 by the table traffic in between) are prefetched into L2
    bool primed = false;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int cur;
            if (!primed) { const int i = atomicAdd(&R.sched_count[5 + cls], 1); cur = i < count ? order[i] : -1; }
            else cur = next_item[1];
            int nxt = -1;
            if (cur >= 0) { const int j = atomicAdd(&R.sched_count[5 + cls], 1); nxt = j < count ? order[j] : -1; }
            next_item[0] = cur; next_item[1] = nxt;
        }
        primed = true;
        __syncthreads();
        const int t = next_item[0];
        if (t < 0) break;
        {
            const int tn = next_item[1];
            const int lane = (int)threadIdx.x - (KM_BUBBLE_THREADS - 32);
            if (tn >= 0 && lane >= 0) {
                const int64_t nb = W.node_off[tn];
                const int cap = (int)(W.node_off[tn + 1] - nb);
                int n = W.n_nodes[tn];
                n = n < cap ? n : cap;
                const char* a0 = reinterpret_cast<const char*>(W.node_kmer + nb);
                const char* a1 = reinterpret_cast<const char*>(W.node_count + nb);
                const char* a2 = reinterpret_cast<const char*>(W.node_slot + nb);
                for (int o = lane * 128; o < 8 * n; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(a0 + o));
                for (int o = lane * 128; o < 4 * n; o += 32 * 128) {
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(a1 + o));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(a2 + o));
                }
            }
        }
        ctx.rot = t & ((KM_BUBBLE_THREADS >> 5) - 1);
        bubble_target<NODES>(ctx, T, W, R, t, B);
    }
#endif
}

// `list`: which work list of the scheduler the pass takes its targets from (its own class's, or the one the bubble pass
// filled with what it handed on)
template <int NODES>
__global__ void __launch_bounds__(KM_CTA, NODES == KM_SMALL_NODES ? KM_GRAPH_SMALL_MINB : (NODES ? KM_GRAPH_TINY_MINB : 4)) km_graph_kernel(TableView T, WalkView W, ScratchLayout SL, ResultView R, int list) {
    extern __shared__ __align__(16) char km_smem[];
    __shared__ int sh[32];
    CtaCtx ctx;
    const GraphScratch S = NODES ? carve(class_layout(NODES ? NODES : 4), km_smem, 1)
                                 : carve(SL, SL.base + (size_t)blockIdx.x * SL.stride, 0);
    // Targets come from this pass's work list (km_schedule_kernel: largest graphs first), handed out one
    // at a time from a global cursor: their cost varies several-fold (a tandem duplication has six times
    // the novel nodes of a substitution), a fixed deal leaves most CTAs idle behind the unluckiest one.
    const int cls = NODES == KM_TINY_NODES ? 0 : NODES == KM_SMALL_NODES ? 1 : 2;
    KernelSpan span(1 + cls);
    unsigned long long* next = R.used + 4 + cls;
    const int32_t* order = R.sched_order + (size_t)list * W.n_targets;
    // Work items are taken ONE AHEAD: while target t is processed, the node arrays the walk left for the next one (evicted
    // from L2 by the table traffic in between) are prefetched into L2, so its numbering phase starts on L2 hits instead
    // of ~4 dependent DRAM round trips.
    bool primed = false;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int cur;
            if (!primed) { const int i = (int)atomicAdd(next, 1ull); cur = i < R.sched_count[list] ? order[i] : -1; }
            else cur = sh[13];
            int nxt = -1;
            if (cur >= 0) { const int j = (int)atomicAdd(next, 1ull); nxt = j < R.sched_count[list] ? order[j] : -1; }
            sh[12] = cur; sh[13] = nxt;
        }
        primed = true;
        __syncthreads();
        const int t = sh[12];
        if (t < 0) break;
        {
            const int tn = sh[13];
            const int lane = (int)threadIdx.x - (KM_CTA - 32);
            if (tn >= 0 && lane >= 0) {
                const int64_t nb = W.node_off[tn];
                const int cap = (int)(W.node_off[tn + 1] - nb);
                int n = W.n_nodes[tn];
                n = n < cap ? n : cap;
                const char* a0 = reinterpret_cast<const char*>(W.node_kmer + nb);
                const char* a1 = reinterpret_cast<const char*>(W.node_count + nb);
                const char* a2 = reinterpret_cast<const char*>(W.node_slot + nb);
                for (int o = lane * 128; o < 8 * n; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(a0 + o));
                for (int o = lane * 128; o < 4 * n; o += 32 * 128) {
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(a1 + o));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(a2 + o));
                }
            }
        }
        if (NODES == 0 && threadIdx.x == 0) atomicAnd(&W.status[t], ~KM_ST_RETRY_LARGE);
        GraphDims d;
        ctx.rot = t & 3;
#ifdef KM_PHASE_TIMERS
        const long long tc0 = clock64();
#endif
#ifdef KM_HAVE_TIMELINE
        const unsigned long long tn0 = km_now_ns();
#endif
        if (!graph_target(ctx, T, W, S, R, t, &d, sh)) continue;
        emit_rows(ctx, T, W, S, R, t, d, sh[2], sh[3], sh[6], sh);
        __syncthreads();
#ifdef KM_PHASE_TIMERS
        if (threadIdx.x == 0 && t < KM_DEBUG_TARGETS) km_target_cycles[t] = (unsigned int)(clock64() - tc0);
#endif
#ifdef KM_HAVE_TIMELINE
        if (threadIdx.x == 0) { const unsigned int i = atomicAdd(&km_timeline_targets, 1u); if (i < 8) { km_timeline[8 + i][0] = tn0; km_timeline[8 + i][1] = km_now_ns(); } }
#endif
    }
}


}  // namespace km

using namespace km;

cudaError_t km_find_kernels_init() {
    cudaError_t e = cudaFuncSetAttribute(km_graph_kernel<KM_TINY_NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)class_layout(KM_TINY_NODES).stride);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(km_graph_kernel<KM_SMALL_NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)class_layout(KM_SMALL_NODES).stride);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(km_graph_bubble_kernel<KM_TINY_NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(KM_BUBBLE_SLOTS * sizeof(BubbleScratch<KM_TINY_NODES>)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(km_graph_bubble_kernel<KM_SMALL_NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(KM_BUBBLE_SLOTS * sizeof(BubbleScratch<KM_SMALL_NODES>)));
    if (e != cudaSuccess) return e;
    return e;
}
// KM_NO_BUBBLE_KERNEL=1 (environment) = A/B switch: every target through the CTA-per-target passes
bool km_bubble_pass_enabled() {
    static const bool on = [] { const char* e = getenv("KM_NO_BUBBLE_KERNEL"); return !(e && *e && *e != '0'); }();
    return on;
}
cudaError_t km_launch_schedule(const WalkView& W, const ResultView& R, cudaStream_t s) {
    km_schedule_kernel<<<1, 1024, 0, s>>>(W, R, km_bubble_pass_enabled() ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t km_launch_bubble(int cls, int grid, const TableView& T, const WalkView& W, const ResultView& R, cudaStream_t s) {
    const int threads = KM_BUBBLE_THREADS == 32 ? 32 * KM_BUBBLE_WARPS : KM_BUBBLE_THREADS;
    if (cls == 0) km_graph_bubble_kernel<KM_TINY_NODES><<<grid, threads, KM_BUBBLE_SLOTS * sizeof(BubbleScratch<KM_TINY_NODES>), s>>>(T, W, R);
    else km_graph_bubble_kernel<KM_SMALL_NODES><<<grid, threads, KM_BUBBLE_SLOTS * sizeof(BubbleScratch<KM_SMALL_NODES>), s>>>(T, W, R);
    return cudaGetLastError();
}
cudaError_t km_launch_graph(int cls, int list, int grid, const TableView& T, const WalkView& W, const ScratchLayout& SL, const ResultView& R,
                            cudaStream_t s) {
    if (cls == 0) km_graph_kernel<KM_TINY_NODES><<<grid, KM_CTA, class_layout(KM_TINY_NODES).stride, s>>>(T, W, SL, R, list);
    else if (cls == 1) km_graph_kernel<KM_SMALL_NODES><<<grid, KM_CTA, class_layout(KM_SMALL_NODES).stride, s>>>(T, W, SL, R, list);
    else km_graph_kernel<0><<<grid, KM_CTA, 0, s>>>(T, W, SL, R, list);
    return cudaGetLastError();
}

// ---- measurement helpers (the phase timers live in this translation unit) -----------------------------------
static int dbg_fail(const char*) { return KM_E_ARG; }
#define fail(code, ...) dbg_fail("")
#define CU(call) do { if ((call) != cudaSuccess) return KM_E_CUDA; } while (0)
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

// out32: 16 x (start, end) in ns; reset: the next launch starts a new record
extern "C" int km_debug_timeline(unsigned long long* out32, int reset) {
#ifdef KM_HAVE_TIMELINE
    if (out32) CU(cudaMemcpyFromSymbol(out32, km_timeline, sizeof(unsigned long long) * 32));
    if (reset) {
        unsigned long long z[32];
        for (int i = 0; i < 16; ++i) { z[2 * i] = ~0ull; z[2 * i + 1] = 0ull; }
        CU(cudaMemcpyToSymbol(km_timeline, z, sizeof(z)));
        const unsigned int zero = 0;
        CU(cudaMemcpyToSymbol(km_timeline_targets, &zero, sizeof(zero)));
    }
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

