// __global__ entry points of libkm_b200.so (sm_100a).  Each kernel is a thin wrapper over
// the stage functions in table.h / walk.h / graph.h / quant.h.
#pragma once
#include <cuda_runtime.h>

#include "quant.h"
#include "walk_small.h"
#include "synth.h"

namespace km {

#define KM_CTA 128

// ---- table maintenance ----------------------------------------------------------------
__global__ void km_table_clear_kernel(Bucket* buckets, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += stride) {
        // two 16-byte stores per bucket, consecutive lanes on consecutive halves
        uint4* p = reinterpret_cast<uint4*>(buckets) + i;
        *p = (i & 1) ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    }
}

__global__ void km_table_insert_kernel(TableView T, const uint64_t* keys, const uint32_t* counts, uint64_t n, int mode,
                                       unsigned long long* n_new, uint32_t* full) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = table_insert(T, keys[i] & T.kmask, counts[i], mode);
        if (r < 0) *full = 1;
        mine += r > 0;
    }
    if (mine) atomicAdd(n_new, mine);
}

__global__ void km_table_synth_kernel(TableView T, uint64_t seed, uint64_t n, unsigned long long* n_new, uint32_t* full) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = synth_key(seed, i, T.k);
        const int r = table_insert(T, key, synth_count(key), KM_INSERT_KEEP);
        if (r < 0) *full = 1;
        mine += r > 0;
    }
    if (mine) atomicAdd(n_new, mine);
}

// K1': canonical k-mer counting from reads (jellyfish count -C; run_leucegene.sh:22)
__global__ void km_count_reads_kernel(TableView T, const char* reads, const int64_t* off, int64_t n_reads, int64_t total,
                                      unsigned long long* n_new, uint32_t* full) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < total; pos += stride) {
        // binary search the read that owns this position
        int64_t lo = 0, hi = n_reads;
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (off[mid] <= pos) lo = mid; else hi = mid; }
        if (pos + T.k > off[lo + 1]) continue;
        uint64_t v = 0; bool ok = true;
        for (int j = 0; j < T.k; ++j) {
            uint64_t c;
            switch (reads[pos + j]) {
                case 'A': case 'a': c = 0; break; case 'C': case 'c': c = 1; break;
                case 'G': case 'g': c = 2; break; case 'T': case 't': c = 3; break;
                default: c = 0; ok = false;
            }
            v = (v << 2) | c;
        }
        if (!ok) continue;
        const int r = table_insert(T, T.canonical ? canonical(v, T.k) : v, 1u, KM_INSERT_ADD);
        if (r < 0) *full = 1;
        mine += r > 0;
    }
    if (mine) atomicAdd(n_new, mine);
}

__global__ void km_table_clear_lines_kernel(Line* lines, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < KM_LINE_SLOTS * n; i += stride)
        reinterpret_cast<uint4*>(lines)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);      // key = empty, count = 0
}

__global__ void km_table_filter_kernel(TableView src, TableView dst, uint32_t min_count, unsigned long long* n_new, uint32_t* full) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    if (src.lines) {
        // every k-mer sits in two lines; both copies are visited, the second insertion finds the key in place
        const LineSlot* slots = reinterpret_cast<const LineSlot*>(src.buckets);
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src.n_buckets * KM_LINE_SLOTS; i += stride) {
            const uint64_t stored = slots[i].key;
            if (stored == KM_EMPTY_KEY || slots[i].count < min_count) continue;
            const int r = table_insert(dst, KM_KEY_OF(stored), slots[i].count, KM_INSERT_KEEP);
            if (r < 0) *full = 1;
            mine += r > 0;
        }
        if (mine) atomicAdd(n_new, mine);
        return;
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src.n_buckets * KM_BUCKET_SLOTS; i += stride) {
        const Bucket* b = src.buckets + i / KM_BUCKET_SLOTS;
        const int s = (int)(i % KM_BUCKET_SLOTS);
        const uint64_t key = b->key[s];
        if (key == KM_EMPTY_KEY || b->count[s] < min_count) continue;
        const int r = table_insert(dst, key, b->count[s], KM_INSERT_KEEP);
        if (r < 0) *full = 1;
        mine += r > 0;
    }
    if (mine) atomicAdd(n_new, mine);
}

// Once per upload, one warp per target: letters -> codes in place (A0 C1 G2 T3, anything else 255; the
// general walk reads these) and the 2-bit packed copy the probe and shared-memory walk kernels read
// (16 bases per word, first base in the top bits, >= 2 zero words after each target).
__global__ void km_encode_kernel(uint8_t* seq, const int64_t* seq_off, uint32_t* pack, const int64_t* pack_off, uint8_t* pre_bad,
                                 int n_targets) {
    const int t = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= n_targets) return;
    const int64_t s0 = seq_off[t], w0 = pack_off[t];
    const int len = (int)(seq_off[t + 1] - s0), nw = (int)(pack_off[t + 1] - w0);
    bool bad = false;
    for (int w = lane; w < nw; w += 32) {
        uint32_t word = 0;
        for (int j = 0; j < 16; ++j) {
            const int pos = 16 * w + j;
            uint32_t c = 0;
            if (pos < len) {
                const uint8_t ch = seq[s0 + pos];
                c = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 255u;
                seq[s0 + pos] = (uint8_t)c;
                if (c > 3u) { bad = true; c = 0u; }
            }
            word = (word << 2) | c;
        }
        pack[w0 + w] = word;
    }
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0) pre_bad[t] = bad ? 1 : 0;
}

// `jellyfish dump`: every (canonical key, count) record of this shard, compacted into two arrays in no
// particular order (a family-line table holds two copies of most k-mers: only the one without the copy bit)
__global__ void km_table_export_kernel(TableView T, uint64_t* keys, uint32_t* counts, unsigned long long cap, unsigned long long* n_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_slots = T.lines ? T.n_buckets * KM_LINE_SLOTS : T.n_buckets * KM_BUCKET_SLOTS;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        uint64_t key; uint32_t cnt;
        if (T.lines) {
            const LineSlot* sl = reinterpret_cast<const LineSlot*>(T.buckets) + i;
            key = sl->key; cnt = sl->count;
            if (key == KM_EMPTY_KEY || (key & KM_COPY_BIT)) continue;
        } else {
            const Bucket* b = T.buckets + i / KM_BUCKET_SLOTS;
            key = b->key[i % KM_BUCKET_SLOTS]; cnt = b->count[i % KM_BUCKET_SLOTS];
            if (key == KM_EMPTY_KEY) continue;
        }
        const unsigned long long at = atomicAdd(n_out, 1ull);
        if (at < cap) { keys[at] = key; counts[at] = cnt; }
    }
}


// ---- K2: batched canonical probe (Jellyfish.query) ---------------------------------------
#define KM_QUERY_ILP 4
__global__ void __launch_bounds__(256) km_query_kernel(TableView T, const uint64_t* __restrict__ kmers, uint64_t n,
                                                       uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // each thread takes KM_QUERY_ILP queries `stride` apart: coalesced key reads and count
    // writes, KM_QUERY_ILP independent random sector reads in flight
    for (uint64_t base = tid; base < n; base += stride * KM_QUERY_ILP) {
        uint64_t q[KM_QUERY_ILP];
        uint32_t r[KM_QUERY_ILP];
        uint32_t live = 0;                 // a padded slot must not be probed: in a sharded table its key may live on a peer
#pragma unroll
        for (int i = 0; i < KM_QUERY_ILP; ++i) {
            const uint64_t j = base + (uint64_t)i * stride;
            q[i] = j < n ? kmers[j] : 0ull;
            live |= j < n ? (1u << i) : 0u;
        }
        if (T.lines) {
#pragma unroll
            for (int i = 0; i < KM_QUERY_ILP; ++i) r[i] = (live >> i) & 1u ? table_query(T, q[i]) : 0u;
        } else {
            table_query_masked<KM_QUERY_ILP>(T, q, live, r);
        }
#pragma unroll
        for (int i = 0; i < KM_QUERY_ILP; ++i) {
            const uint64_t j = base + (uint64_t)i * stride;
            if (j < n) out[j] = r[i];
        }
    }
}

// The same for a table of family lines: FOUR LANES PER QUERY.  An isolated k-mer sits somewhere in the
// 128-byte line of its prefix family; lane j of a quad loads sector j, so the line is one coalesced request
// (what a random 32-byte read costs in DRAM anyway), each lane checks its two slots and the quad combines by
// shuffle.  KM_QUERY_ILP queries per quad are in flight.
__global__ void __launch_bounds__(256) km_query_lines_kernel(TableView T, const uint64_t* __restrict__ kmers, uint64_t n,
                                                             uint32_t* __restrict__ out) {
    const uint64_t n_quads = ((uint64_t)gridDim.x * blockDim.x) >> 2;
    const uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int sec = threadIdx.x & 3;
    for (uint64_t base = quad; base < n; base += n_quads * KM_QUERY_ILP) {      // the quads of a warp stay together: n_quads is a multiple of 8
        uint64_t key[KM_QUERY_ILP], k0[KM_QUERY_ILP], k1[KM_QUERY_ILP], idx[KM_QUERY_ILP];
        uint32_t c0[KM_QUERY_ILP], c1[KM_QUERY_ILP];
        const Line* lb[KM_QUERY_ILP];
#pragma unroll
        for (int i = 0; i < KM_QUERY_ILP; ++i) {
            const uint64_t j = base + (uint64_t)i * n_quads;
            const uint64_t v = (j < n ? kmers[j] : 0ull) & T.kmask;
            key[i] = T.canonical ? canonical(v, T.k) : v;
            lb[i] = locate_line(T, family_of_prefix(T, v), &idx[i]);
            k0[i] = k1[i] = KM_EMPTY_KEY; c0[i] = c1[i] = 0;
            if (j < n) load_sector(lb[i][idx[i]].s + 2 * sec, k0[i], c0[i], k1[i], c1[i]);
        }
#pragma unroll
        for (int i = 0; i < KM_QUERY_ILP; ++i) {
            const uint64_t j = base + (uint64_t)i * n_quads;
            // bits 0..31 count, 32 found, 33 the line has an empty slot
            unsigned long long r = 0ull;
            if (KM_KEY_OF(k0[i]) == key[i]) r = (1ull << 32) | c0[i];
            if (KM_KEY_OF(k1[i]) == key[i]) r = (1ull << 32) | c1[i];
            if (k0[i] == KM_EMPTY_KEY || k1[i] == KM_EMPTY_KEY) r |= 1ull << 33;
            r |= __shfl_xor_sync(0xFFFFFFFFu, r, 1);
            r |= __shfl_xor_sync(0xFFFFFFFFu, r, 2);
            if (sec == 0 && j < n) {
                uint32_t cnt = (uint32_t)r;
                if (!(r >> 32)) {            // full line without the key: the next line, on this lane's own (rare)
                    const uint64_t keys1[1] = {key[i]};
                    uint32_t r1[1] = {0};
                    line_find_from<1>(T, lb[i], idx[i] + 1 == T.n_buckets ? 0 : idx[i] + 1, keys1, 1u, r1);
                    cnt = r1[0];
                }
                out[j] = cnt;
            }
        }
    }
}

// Jellyfish.get_child (Jellyfish.py:55-72), one thread per k-mer
__global__ void __launch_bounds__(256) km_get_child_kernel(TableView T, const uint64_t* __restrict__ kmers, uint64_t n,
                                                           int forward, double ratio, int64_t floor_count,
                                                           uint32_t* __restrict__ counts, uint8_t* __restrict__ mask) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t v = kmers[i] & T.kmask;
        uint64_t ck[4]; uint32_t cc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) ck[c] = forward ? succ_kmer(v, c, T.kmask) : pred_kmer(v, c, T.k);
        table_query_family<4>(T, forward ? family_of_suffix(T, v) : family_of_prefix(T, v), ck, 15u, cc);
        const uint64_t sum = (uint64_t)cc[0] + cc[1] + cc[2] + cc[3];
        double thr = (double)sum * ratio;
        if (thr < (double)floor_count) thr = (double)floor_count;
        uint8_t m = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) { counts[4 * i + c] = cc[c]; m |= ((double)cc[c] >= thr) ? (1u << c) : 0u; }
        mask[i] = m;
    }
}

// ---- K3: walk, one WARP per target ----------------------------------------------------------
// Two launches: the shared-memory walk takes every target of ordinary size (walk_small.h); the
// general walk, whose per-target state lives in HBM, takes the targets the first one deferred.
#ifndef KM_WALK_WARPS
#define KM_WALK_WARPS 4
#endif
#ifndef KM_WALK_MINB
#define KM_WALK_MINB 9       // (measured: 7 -> 0.271, 8 -> 0.262, 9 -> 0.256, 10 -> 0.296 ms) CTAs per SM the shared-memory walk's registers are budgeted for
#endif
// (measured: 4 warps per CTA and room for ~100 registers -- no spills -- beat 8 warps at 64 registers, 0.307 vs 0.333 ms)
#ifndef KM_PROBE_WARPS
#define KM_PROBE_WARPS 4
#endif
#ifndef KM_PROBE_MINB
#define KM_PROBE_MINB 5
#endif
// K3a: level 0 of every walk, one warp per 32 reference k-mers, flat over the batch
__global__ void __launch_bounds__(32 * KM_PROBE_WARPS, KM_PROBE_MINB) km_ref_probe_kernel(TableView T, WalkView W, FindParams P) {
    WarpCtx ctx;
    const int ch = (int)blockIdx.x * KM_PROBE_WARPS + (int)(threadIdx.x >> 5);
    if (ch >= W.n_chunks) return;
    ref_probe_chunk(ctx, T, W, P, W.chunk_target[ch], W.chunk_start[ch]);
}

__global__ void __launch_bounds__(32 * KM_WALK_WARPS, KM_WALK_MINB) km_walk_small_kernel(TableView T, WalkView W, FindParams P) {
    __shared__ WalkSmall M[KM_WALK_WARPS];
    WarpCtx ctx;
    const int t = (int)blockIdx.x * KM_WALK_WARPS + (int)(threadIdx.x >> 5);
    if (t >= W.n_targets) return;
    const TargetGeom g = target_geom(W, t, T.k);
    if (!walk_small_fits(g)) { if ((threadIdx.x & 31) == 0) W.status[t] = KM_ST_WALK_DEFER; return; }   // also drops the probe's limit flag
    walk_small_target(ctx, T, W, P, t, M[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(32 * KM_WALK_WARPS) km_walk_kernel(TableView T, WalkView W, FindParams P) {
    WarpCtx ctx;
    const int t = (int)blockIdx.x * KM_WALK_WARPS + (int)(threadIdx.x >> 5);
    if (t >= W.n_targets) return;
    if (!(W.status[t] & KM_ST_WALK_DEFER)) return;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) { W.status[t] = 0; W.lookups[t] = 0; W.n_kept[t] = 0; }
    __syncwarp();
    walk_target(ctx, T, W, P, t);
}

// ---- K4 + K5: graph, paths, FP64 quantification; persistent CTAs over targets -----------------
// Three passes share one body.  The two SHARED-MEMORY passes keep the whole per-target working set
// (adjacency, both shortest-path trees, candidate edges, solver matrices) on chip; a target goes to the
// smallest class its graph fits, so the many small graphs run with twice the CTAs per SM of the
// larger ones (the pass is latency-bound: resident CTAs are throughput).  The GENERAL pass uses
// per-CTA scratch in HBM and takes the rest, plus any target a shared-memory pass deferred
// (KM_ST_RETRY_LARGE).
#ifndef KM_SMALL_NODES
#define KM_SMALL_NODES 512      // largest shared-memory class (graph nodes incl. the two caps)
#endif
// resident CTAs per SM the register allocation aims at: the 512-node class is held to 5 by its shared memory
#ifndef KM_GRAPH_SMALL_MINB
#define KM_GRAPH_SMALL_MINB 5
#endif
#ifndef KM_GRAPH_TINY_MINB
#define KM_GRAPH_TINY_MINB 8
#endif
// persistent CTAs per SM launched for each class (they take targets from a shared cursor)
#ifndef KM_GRAPH_SMALL_GRID
#define KM_GRAPH_SMALL_GRID 5
#endif
#ifndef KM_GRAPH_TINY_GRID
#define KM_GRAPH_TINY_GRID 10
#endif
#ifndef KM_TINY_NODES
#define KM_TINY_NODES 256
#endif
#define KM_SMALL_CAND 64
#define KM_SMALL_PATHS 64
#define KM_SMALL_COLS 8

__host__ __device__ inline ScratchLayout class_layout(int nodes) {
    return make_layout(nodes - 2, KM_SMALL_CAND, KM_SMALL_PATHS, KM_SMALL_COLS, 1);
}

#define KM_ST_FATAL (KM_ST_BAD_BASE | KM_ST_DUP_KMER | KM_ST_NODE_OVERFLOW | KM_ST_NODE_LIMIT | KM_ST_TOO_SHORT)

// NODES = node capacity of a shared-memory class, 0 = the general pass
// Work lists of the three graph passes: every target whose walk succeeded goes to the smallest class
// its graph fits, each list ordered by descending node count (64 size bins; a counting sort in one CTA).
__global__ void __launch_bounds__(1024) km_schedule_kernel(WalkView W, ResultView R) {
    __shared__ int hist[3][64], start[3][64];
    const int n = W.n_targets;
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    auto classify = [&](int t, int* bin) -> int {
        if (W.status[t] & KM_ST_FATAL) return -1;
        const int cap = (int)(W.node_off[t + 1] - W.node_off[t]);
        const int n_all = W.n_nodes[t] < cap ? W.n_nodes[t] : cap;
        const int kept2 = W.n_kept[t] + 2;
        const int b = 63 - (kept2 >> 3);
        *bin = b < 0 ? 0 : b;                                         // bin 0 = the largest graphs
        if (n_all <= KM_TINY_NODES - 2 && kept2 <= KM_TINY_NODES) return 0;
        if (n_all <= KM_SMALL_NODES - 2 && kept2 <= KM_SMALL_NODES) return 1;
        return 2;
    };
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        int b;
        const int c = classify(t, &b);
        if (c < 0) { R.t_n[t] = 0; R.t_n_paths[t] = 0; R.t_path_first[t] = 0; R.t_n_rows[t] = 0; R.t_row_first[t] = 0; }
        else atomicAdd(&hist[c][b], 1);
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int at = 0;
        for (int b = 0; b < 64; ++b) { start[threadIdx.x][b] = at; at += hist[threadIdx.x][b]; }
        R.sched_count[threadIdx.x] = at;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        int b;
        const int c = classify(t, &b);
        if (c >= 0) R.sched_order[(size_t)c * n + atomicAdd(&start[c][b], 1)] = t;
    }
}

template <int NODES>
__global__ void __launch_bounds__(KM_CTA, NODES == KM_SMALL_NODES ? KM_GRAPH_SMALL_MINB : (NODES ? KM_GRAPH_TINY_MINB : 4)) km_graph_kernel(TableView T, WalkView W, ScratchLayout SL, ResultView R) {
    extern __shared__ __align__(16) char km_smem[];
    __shared__ int sh[32];
    CtaCtx ctx;
    const GraphScratch S = NODES ? carve(class_layout(NODES ? NODES : 4), km_smem, 1)
                                 : carve(SL, SL.base + (size_t)blockIdx.x * SL.stride, 0);
    // Targets come from this pass's work list (km_schedule_kernel: largest graphs first), handed out one
    // at a time from a global cursor: their cost varies several-fold (a tandem duplication has six times
    // the novel nodes of a substitution), a fixed deal leaves most CTAs idle behind the unluckiest one.
    const int cls = NODES == KM_TINY_NODES ? 0 : NODES == KM_SMALL_NODES ? 1 : 2;
    unsigned long long* next = R.used + 4 + cls;
    const int32_t* order = R.sched_order + (size_t)cls * W.n_targets;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const int i = (int)atomicAdd(next, 1ull);
            sh[12] = i < R.sched_count[cls] ? order[i] : -1;
        }
        __syncthreads();
        const int t = sh[12];
        if (t < 0) break;
        if (NODES == 0 && threadIdx.x == 0) atomicAnd(&W.status[t], ~KM_ST_RETRY_LARGE);
        GraphDims d;
        ctx.rot = t & 3;
#ifdef KM_PHASE_TIMERS
        const long long tc0 = clock64();
#endif
        if (!graph_target(ctx, T, W, S, R, t, &d, sh)) continue;
        emit_rows(ctx, T, W, S, R, t, d, sh[2], sh[3], sh[6], sh);
        __syncthreads();
#ifdef KM_PHASE_TIMERS
        if (threadIdx.x == 0 && t < KM_DEBUG_TARGETS) km_target_cycles[t] = (unsigned int)(clock64() - tc0);
#endif
    }
}

// ---- measurement kernels -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) km_gather_kernel(const uint4* __restrict__ buf, uint64_t n_sectors, uint64_t n_loads,
                                                        uint64_t seed, uint32_t* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n_loads; base += stride * 4) {
        uint64_t s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = mulhi64(mix64(seed + (base + i * stride) * KM_GOLDEN), n_sectors);
        uint64_t a[4], b[4]; uint32_t c[4], e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) load_bucket(reinterpret_cast<const Bucket*>(buf) + s[i], a[i], b[i], c[i], e[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc ^= (uint32_t)a[i] ^ (uint32_t)b[i] ^ c[i] ^ e[i];
    }
    if (acc == 0x12345678u) *sink = acc;   // keeps the loads alive
}

// config-4 lookup mix generated on device: even j -> a background key (random strand), odd j -> random k-mer
__global__ void km_make_queries_kernel(uint64_t* q, uint64_t n, uint64_t table_seed, uint64_t table_n, uint64_t seed, int k) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t mask = kmer_mask(k);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t r = mix64(seed + (j + 1) * KM_GOLDEN);
        uint64_t v;
        if ((r & 1ull) && table_n) {
            const uint64_t i = mulhi64(mix64(r), table_n);
            v = mix64(table_seed + (i + 1) * KM_GOLDEN) & mask;
            if (r & 2ull) v = revcomp(v, k);
        } else {
            v = mix64(r ^ 0xA5A5A5A5A5A5A5A5ull) & mask;
        }
        q[j] = v;
    }
}

__global__ void km_count_nonzero_kernel(const uint32_t* c, uint64_t n, unsigned long long* out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) mine += c[j] != 0;
    if (mine) atomicAdd(out, mine);
}

}  // namespace km
