// Stage 1 of km_find_batch: node discovery for one target by one CTA.
//
// Reproduces MutationFinder.__init__ / __extend (km/utils/MutationFinder.py:108-120,
// 137-165) in an order-independent form (see oracle/km_oracle.py:walk_closure, which is
// checked against the literal DFS and the unmodified reference):
//   phase 1  every reference k-mer is registered with its count              (:111-112)
//   phase 2  level-synchronous forward expansion from all of them; each node issues the
//            4 successor lookups of Jellyfish.get_child (km/utils/Jellyfish.py:55-72),
//            applies threshold = max(sum*ratio, count) and `count >= threshold`, and
//            inserts accepted, unseen children into the target's visited set (atomicCAS)
//   phase 3  novel nodes none of whose accepted children survive are peeled off until a
//            fixed point: what remains is exactly what the DFS commits at :159-163
// max_stack / max_break are honoured per node as (min depth, min branch count at that
// depth); a target that touches either limit is flagged (the reference's own result is
// iteration-order dependent there).
#pragma once
#include "table.h"

namespace km {

#define KM_NO_SLOT 0xFFFFFFFFu

// status bits per target (same values as include/km_b200.h)
#ifndef KM_ST_BAD_BASE
#define KM_ST_BAD_BASE 1          // non-ACGT letter in the target (parity unpinned; host raises)
#define KM_ST_DUP_KMER 2          // repeated k-mer in the target (common.py:55-59 -> ValueError)
#define KM_ST_NODE_OVERFLOW 4     // explored more nodes than this launch's capacity -> host retries larger
#define KM_ST_NODE_LIMIT 8        // kept nodes > max_node (MutationFinder.py:143-148 -> sys.exit)
#define KM_ST_TOUCHED_LIMIT 16    // some walk hit max_stack / max_break
#define KM_ST_PATH_OVERFLOW 32    // path/row pool exhausted -> host retries larger
#define KM_ST_TOO_SHORT 64        // target shorter than k
#define KM_ST_TOO_MANY_COLS 128   // a cluster has more paths than the solver's column capacity
#define KM_ST_SOLVER_WATCHDOG 256 // refine loop exceeded the watchdog (reference has no cap)
#define KM_ST_NAME_MISMATCH 512   // MutationFinder.py:431-440 length check failed
#endif

// internal, for the scheduler of the graph passes (cleared there): the walk saw more than one way off the reference, or a
// novel node with several accepted children -- the target's graph is probably not a simple bubble (graph_bubble.h)
#define KM_ST_BRANCHED 0x10000000u

struct FindParams {
    double ratio;       // -p / Jellyfish cutoff
    int64_t count;      // -c / Jellyfish n_cutoff
    int32_t max_stack;  // -s
    int32_t max_break;  // -b
    int32_t max_node;   // -n
};

// Device-side view of one batch of targets (all pointers into HBM).
struct WalkView {
    int n_targets;
    const uint8_t* codes;     // concatenated targets, one 2-bit code per byte, 255 = not ACGT
    const int64_t* seq_off;   // [n+1]
    const int64_t* node_off;  // [n+1]  node capacity of target t = node_off[t+1]-node_off[t]
    const int64_t* hash_off;  // [n+1]  visited-set slots (power of two per target)
    // the same targets 2-bit packed, 16 bases per word, first base in the top bits; every target starts on a
    // word and is followed by >= 2 zero words (written once per upload by km_encode_kernel)
    const uint32_t* pack;
    const int64_t* pack_off;  // [n+1]  word offsets
    const uint8_t* pre_bad;   // [n]    1 = the target holds a letter outside ACGT
    // node arrays, discovery order: reference k-mers first (index = position), then novel
    uint64_t* node_kmer;
    uint32_t* node_count;
    uint32_t* node_slot;      // slot of the node in the target's visited set
    uint32_t* node_kid;       // [4 per node] visited-set slot of each accepted child, KM_NO_SLOT otherwise
    // visited set (open addressing, linear probing)
    uint64_t* hkey;
    uint32_t* hval;           // node index
    uint32_t* hmeta;          // (depth << 8) | branch count, combined with atomicMin
    uint8_t* hflag;           // 1 = kept
    // per target
    int32_t* n_nodes;         // explored nodes
    int32_t* n_kept;          // kept nodes (excludes the two caps)
    uint32_t* status;
    unsigned long long* lookups;  // [n] table lookups issued (measurement only)
    // [n] what the walk already knows about the target's graph pass, + 1 (0 = not said: the scheduler works it out):
    // bits 0-1 size class (0 / 1 the shared-memory classes, 2 general), bit 2 the walk branched, bits 3.. size bin;
    // 0xFFFF = no graph.  May be null (the CPU emulation has no scheduler).
    uint16_t* sched_code;
    uint32_t* walk_cursor;        // next target of the persistent walk warps (cleared with the per-target state)
    // chunks of <= 32 consecutive reference k-mers, flat over all targets (ref_probe_chunk)
    const int32_t* chunk_target;
    const int32_t* chunk_start;
    int n_chunks;
};

#define KM_ST_FATAL (KM_ST_BAD_BASE | KM_ST_DUP_KMER | KM_ST_NODE_OVERFLOW | KM_ST_NODE_LIMIT | KM_ST_TOO_SHORT)
// The graph pass's work-list code of a target (km_schedule_kernel): `st` its final status, `n_all` explored nodes (capped),
// `kept` kept nodes without the caps.
KM_HD uint32_t sched_code_of(uint32_t st, int n_all, int kept, int tiny_nodes, int small_nodes) {
    if (st & KM_ST_FATAL) return 0xFFFFu;
    const int kept2 = kept + 2;
    int b = 63 - (kept2 >> 3);
    b = b < 0 ? 0 : b;                                            // bin 0 = the largest graphs
    int c = 2;
    if (n_all <= tiny_nodes - 2 && kept2 <= tiny_nodes) c = 0;
    else if (n_all <= small_nodes - 2 && kept2 <= small_nodes) c = 1;
    return (uint32_t)(c | ((st & KM_ST_BRANCHED) ? 4 : 0) | (b << 3));
}

KM_HD uint32_t pack_meta(int depth, int breaks) { return ((uint32_t)depth << 8) | (uint32_t)(breaks > 255 ? 255 : breaks); }

struct TargetGeom {
    int L;            // reference k-mers
    int cap;          // node capacity
    uint32_t hmask;   // hash slots - 1
    int64_t nbase;    // node array base
    int64_t hbase;    // hash array base
    int64_t sbase;    // sequence base
};

KM_HD TargetGeom target_geom(const WalkView& W, int t, int k) {
    TargetGeom g;
    g.sbase = W.seq_off[t];
    int len = (int)(W.seq_off[t + 1] - g.sbase);
    g.L = len - k + 1;
    g.nbase = W.node_off[t];
    g.cap = (int)(W.node_off[t + 1] - g.nbase);
    g.hbase = W.hash_off[t];
    g.hmask = (uint32_t)(W.hash_off[t + 1] - g.hbase) - 1u;
    return g;
}

// k-mer starting at base i of a packed sequence (k <= 31; words[w+2] must be readable)
KM_HD uint64_t packed_kmer(const uint32_t* words, int i, int k) {
    const int w = i >> 4, o = i & 15;
    const uint64_t hi = ((uint64_t)words[w] << 32) | (uint64_t)words[w + 1];
    const uint64_t x = o ? ((hi << (2 * o)) | ((uint64_t)words[w + 2] >> (32 - 2 * o))) : hi;
    return x >> (64 - 2 * k);
}
KM_HD int packed_base(const uint32_t* words, int i) { return (int)((words[i >> 4] >> (2 * (15 - (i & 15)))) & 3u); }

// Find `key` in the target's visited set or claim a slot for it.  Returns the slot;
// *is_new tells whether THIS caller won the claim.
KM_HD uint32_t visited_find_or_insert(const WalkView& W, const TargetGeom& g, uint64_t key, int* is_new) {
    uint32_t s = (uint32_t)mix64(key) & g.hmask;
    *is_new = 0;
    for (uint32_t probes = 0; probes <= g.hmask; ++probes) {
        uint64_t* p = W.hkey + g.hbase + s;
        uint64_t cur = load_cg64(p);
        if (cur == KM_EMPTY_KEY) {
            cur = atomic_cas64(p, KM_EMPTY_KEY, key);
            if (cur == KM_EMPTY_KEY) { *is_new = 1; return s; }
        }
        if (cur == key) return s;
        s = (s + 1) & g.hmask;
    }
    return KM_NO_SLOT;   // set full: only possible after a node overflow, which the host retries
}

// Read-only find (valid after the walk kernel has finished).  KM_NO_SLOT when absent.
KM_HD uint32_t visited_find(const WalkView& W, const TargetGeom& g, uint64_t key) {
    uint32_t s = (uint32_t)mix64(key) & g.hmask;
    for (;;) {
        uint64_t cur = W.hkey[g.hbase + s];
        if (cur == key) return s;
        if (cur == KM_EMPTY_KEY) return KM_NO_SLOT;
        s = (s + 1) & g.hmask;
    }
}

// Expand node q: Jellyfish.get_child(forward=True) + the child loop of __extend.
KM_HD void expand_node(const TableView& T, const WalkView& W, const TargetGeom& g, int t,
                       const FindParams& P, int q, uint32_t* st_bits, unsigned* n_lookups) {
    const uint64_t kmer = W.node_kmer[g.nbase + q];
    const uint32_t meta = load_cg32(&W.hmeta[g.hbase + W.node_slot[g.nbase + q]]);
    const int depth = (int)(meta >> 8), breaks = (int)(meta & 255u);
    uint32_t* kid = W.node_kid + 4 * (g.nbase + q);
    kid[0] = kid[1] = kid[2] = kid[3] = KM_NO_SLOT;
    if (depth > P.max_stack) {               // MutationFinder.py:140-141
        *st_bits |= KM_ST_TOUCHED_LIMIT;
        return;
    }
    // four independent sector reads in flight per thread
    uint64_t ck[4]; uint32_t cc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ck[c] = succ_kmer(kmer, c, T.kmask);
    table_query_family<4>(T, family_of_suffix(T, kmer), ck, 15u, cc);
    *n_lookups += 4;
    // Jellyfish.py:61-72 -- Python int * float, then max with the int floor, then >=
    const uint64_t sum = (uint64_t)cc[0] + cc[1] + cc[2] + cc[3];
    double thr = (double)sum * P.ratio;
    if (thr < (double)P.count) thr = (double)P.count;
    int nk = 0;
    bool pass[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { pass[c] = (double)cc[c] >= thr; nk += pass[c] ? 1 : 0; }
    int nb = breaks;
    if (nk > 1) {                            // MutationFinder.py:153-156
        nb = breaks + 1;
        if (nb > P.max_break) { *st_bits |= KM_ST_TOUCHED_LIMIT; return; }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (!pass[c]) continue;
        int is_new;
        if ((int)load_cg32(reinterpret_cast<const uint32_t*>(&W.n_nodes[t])) >= g.cap) { *st_bits |= KM_ST_NODE_OVERFLOW; continue; }
        const uint32_t s = visited_find_or_insert(W, g, ck[c], &is_new);
        if (s == KM_NO_SLOT) { *st_bits |= KM_ST_NODE_OVERFLOW; continue; }
        kid[c] = s;
        atomic_min32(&W.hmeta[g.hbase + s], pack_meta(depth + 1, nb));
        if (is_new) {
            const int idx = atomic_addi32(&W.n_nodes[t], 1);
            if (idx >= g.cap) { *st_bits |= KM_ST_NODE_OVERFLOW; continue; }
            W.node_kmer[g.nbase + idx] = ck[c];
            W.node_count[g.nbase + idx] = cc[c];
            W.node_slot[g.nbase + idx] = s;
            W.hval[g.hbase + s] = (uint32_t)idx;
            W.hflag[g.hbase + s] = 1;
        }
    }
}

template <class Ctx>
KM_HD void walk_target(const Ctx& ctx, const TableView& T, const WalkView& W, const FindParams& P, int t) {
    const int k = T.k;
    const TargetGeom g = target_geom(W, t, k);
    const int tid = ctx.tid(), nt = ctx.nt();
    uint32_t st = 0;
    unsigned nlook = 0;

    // clear the visited set (keys empty, meta = +inf, flags 0)
    const uint32_t H = g.hmask + 1u;
    for (uint32_t s = tid; s < H; s += nt) {
        W.hkey[g.hbase + s] = KM_EMPTY_KEY;
        W.hmeta[g.hbase + s] = 0xFFFFFFFFu;
        W.hflag[g.hbase + s] = 0;
    }
    if (tid == 0) { W.n_nodes[t] = g.L > 0 ? g.L : 0; W.n_kept[t] = 0; }
    if (g.L <= 0) {
        if (tid == 0) { W.status[t] = KM_ST_TOO_SHORT; W.lookups[t] = 0; }
        return;
    }
    ctx.sync();

    // phase 1: register the reference k-mers (MutationFinder.py:111-112)
    for (int i = tid; i < g.L; i += nt) {
        uint64_t v = 0;
        bool bad = false;
        for (int j = 0; j < k; ++j) {
            const uint8_t c = W.codes[g.sbase + i + j];
            bad |= c > 3;
            v = (v << 2) | (uint64_t)(c & 3);
        }
        if (bad) { st |= KM_ST_BAD_BASE; }
        int is_new;
        const uint32_t s = visited_find_or_insert(W, g, v, &is_new);
        if (!is_new) st |= KM_ST_DUP_KMER;
        else { W.hval[g.hbase + s] = (uint32_t)i; W.hflag[g.hbase + s] = 1; }
        atomic_min32(&W.hmeta[g.hbase + s], pack_meta(1, 0));
        W.node_kmer[g.nbase + i] = v;
        W.node_slot[g.nbase + i] = s;
        W.node_count[g.nbase + i] = table_query(T, v);
        nlook += 1;
    }
    // a malformed target stops here: the host raises before any walk (common.py:55-59)
    if (ctx.sync_or((st & (KM_ST_BAD_BASE | KM_ST_DUP_KMER)) != 0)) {
        if (st) atomic_or32(&W.status[t], st);
        if (nlook) atomic_add64(&W.lookups[t], nlook);
        return;
    }

    // phase 2: breadth-first expansion; the node array itself is the queue
    int lo = 0, hi = g.L;
    while (lo < hi) {
        for (int q = lo + tid; q < hi; q += nt) expand_node(T, W, g, t, P, q, &st, &nlook);
        ctx.sync();
        lo = hi;
        int n = (int)load_cg32(reinterpret_cast<const uint32_t*>(&W.n_nodes[t]));
        hi = n < g.cap ? n : g.cap;
        if (n > g.cap) st |= KM_ST_NODE_OVERFLOW;
        ctx.sync();
    }
    const int n_all = hi;

    // phase 3: peel novel nodes with no surviving accepted child (commit rule, :159-163)
    int changed = 1;
    while (changed) {
        int mine = 0;
        for (int q = g.L + tid; q < n_all; q += nt) {
            const uint32_t s = W.node_slot[g.nbase + q];
            if (!load_cg8(&W.hflag[g.hbase + s])) continue;
            const uint32_t* kid = W.node_kid + 4 * (g.nbase + q);
            bool alive = false;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (kid[c] != KM_NO_SLOT && load_cg8(&W.hflag[g.hbase + kid[c]])) alive = true;
            if (!alive) { W.hflag[g.hbase + s] = 0; mine = 1; }
        }
        changed = ctx.sync_or(mine);
    }

    // kept-node count and the node limit (MutationFinder.py:143-148)
    // a dropped node's slot entry is overwritten with KM_NO_SLOT: the graph pass reads one array
    // (node_slot) to tell kept from dropped instead of chasing the visited set
    int kept = 0;
    for (int q = g.L + tid; q < n_all; q += nt) {
        if (load_cg8(&W.hflag[g.hbase + W.node_slot[g.nbase + q]])) kept += 1;
        else W.node_slot[g.nbase + q] = KM_NO_SLOT;
    }
    if (kept) atomic_addi32(&W.n_kept[t], kept);
    st |= KM_ST_BRANCHED;                           // what needed this kernel is no simple bubble
    if (st) atomic_or32(&W.status[t], st);
    if (nlook) atomic_add64(&W.lookups[t], nlook);
    ctx.sync();
    if (tid == 0) {
        const int total = g.L + (int)load_cg32(reinterpret_cast<const uint32_t*>(&W.n_kept[t]));
        W.n_kept[t] = total;
        if (total > P.max_node) atomic_or32(&W.status[t], KM_ST_NODE_LIMIT);
    }
}

}  // namespace km
