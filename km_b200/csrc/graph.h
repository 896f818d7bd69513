// Stage 2 of km_find_batch: per-target graph analysis by one CTA.
//
// Reproduces MutationFinder.graph_analysis (km/utils/MutationFinder.py:496-572) and
// km/utils/Graph.py on the node set kept by the walk:
//   * canonical node numbering: reference k-mers 0..L-1 in sequence order, kept novel
//     k-mers by ascending packed value, then BigBang = N-2, BigCrunch = N-1
//     (MutationFinder.py:122-123; numbering rule shared with oracle/km_oracle.py)
//   * (k-1)-overlap edges found by probing the target's visited set with the 4 successor /
//     4 predecessor k-mers of every node instead of a dense N x N matrix (:515-531); weights
//     1.0f, reference-consecutive and cap edges 0.01f (:535-551)
//   * two single-source shortest-path trees in float32 with the reference's tie rules
//     (Graph.py:63-119: settle the open node of least distance, lowest index on ties;
//     re-parent only on a strictly smaller float32 sum)
//   * removal of the reference chain's edges (Graph.py:178-198), then one candidate path
//     per remaining edge (Graph.py:200-240), de-duplicated and sorted lexicographically
#pragma once
#include "walk.h"

namespace km {

// capacities of the general (HBM-scratch) pass; the shared-memory pass uses smaller ones and
// hands a target over to the general pass when it exceeds them
#define KM_MAX_PATHS 1024       // unique alternative paths per target
#define KM_MAX_COLS 64          // columns of one least-squares problem (1 + cluster size)
#ifndef KM_PCACHE
#define KM_PCACHE 0             // 16-bit node numbers of a target's paths kept in shared memory (shared-memory passes);
                                // measured: 2048 saves 8 % of the pass's cycles per target but costs co-residency of the two
                                // classes' CTAs -- 0.717 ms against 0.695 without (profiles/r2g_variants.txt) -- so it is off
#endif
#define KM_ST_RETRY_LARGE 0x40000000u   // internal: redo this target in the general pass

// a (possibly clipped) path: idx == nullptr means the reference path, whose node at
// position p is simply p
struct PathView {
    const int32_t* idx;
    int begin;
    int len;
    const uint16_t* c16;     // the same node numbers in shared memory (16 bit), or nullptr: see GraphScratch::pcache
    // a SIMPLE BUBBLE's alternative path needs no memory at all outside its chain: position q holds q up to node a, then
    // the chain's nodes (`idx` points at them, in shared memory), then b, b+1, ...   (bub_nk < 0: not such a path)
    int bub_a, bub_nk, bub_b;
};
KM_HD PathView range_view(int begin, int len) {
    PathView v; v.idx = nullptr; v.begin = begin; v.len = len; v.c16 = nullptr; v.bub_a = 0; v.bub_nk = -1; v.bub_b = 0;
    return v;
}

// One output row (PathQuant.Path, km/utils/PathQuant.py:10-49) in numeric form; the host
// spells the strings from the node k-mers.  Mirrored by km_row in include/km_b200.h.
struct Row {
    int32_t target;
    int32_t kind;          // 0 = vs_ref, 1 = cluster
    int32_t type;          // 0 Reference 1 Substitution 2 ITD 3 Indel 4 Insertion 5 Deletion
    int32_t name_start;    // diff.start + k + offset          (MutationFinder.py:485)
    int32_t name_end;      // diff.end_ref + 1 + offset        (MutationFinder.py:487)
    int32_t path_id;       // global id of the variant path
    int32_t var_begin, var_end;   // slice of that path printed as Sequence
    int32_t ref_begin, ref_end;   // slice of the reference path printed as Reference_sequence
    int32_t del_begin, del_len;   // deleted bases = last base of reference k-mers [del_begin, +del_len)
    int32_t ins_begin, ins_len;   // inserted bases = last base of path nodes at positions [ins_begin, +ins_len)
    int32_t start_off;
    int32_t cluster_id, cluster_n;
    int32_t n_iter;        // refine_coef iterations (PathQuant.py:128-136)
    int64_t min_cov;
    double rvaf, expr, ref_rvaf, ref_expr;
};

struct ResultView {
    // per target
    int32_t* t_n;            // node count incl. the two caps
    int32_t* t_n_paths;
    int32_t* t_path_first;
    int32_t* t_n_rows;
    int32_t* t_row_first;
    // canonical node arrays, same offsets as WalkView::node_off
    uint64_t* out_kmer;
    uint32_t* out_count;
    // path table
    int64_t* path_off;
    int32_t* path_len;
    int32_t* pool;
    int32_t path_cap;
    int64_t pool_cap;
    Row* rows;
    int32_t row_cap;
    // spelled unique paths (MutationFinder.get_seq): nullptr = do not spell
    char* seq_pool;
    int64_t* path_seq_off;   // [path_cap] offset of each path's string, -1 = did not fit
    int64_t seq_cap;
    // [0] paths used, [1] pool ints used, [2] rows used, [3] sequence chars used, [4..6] work cursors of
    // the three graph passes
    unsigned long long* used;
    // work lists of the three graph passes (km_schedule_kernel): targets of each class, largest graphs
    // first; sched_count[c] entries in sched_order[c * n_targets ...]
    int32_t* sched_order;
    int32_t* sched_count;
    int32_t flags;           // KM_RESULT_*
};
#define KM_RESULT_NO_REFINE_JUMP 1   // refine_coef is iterated literally from start to end (A/B switch, KM_FIND_NO_REFINE_JUMP)

// Per-target working set of the CTA: shared memory in the small pass, per-CTA HBM scratch
// (L2 resident in practice) in the general pass.  Arrays whose lifetimes do not overlap share
// storage in the compact (shared-memory) layout -- see make_layout.
struct GraphScratch {
    int32_t* newidx;   // [maxN]     discovery index -> canonical index, -1 dropped; later nxtB
    int32_t* kept;     // [maxcap]   discovery indices of kept novel nodes          (numbering only)
    uint64_t* keptk;   // [maxcap]   their packed k-mers                             (numbering only)
    uint64_t* hk;      // [hcap]     k-mer -> canonical index map of this target's nodes (adjacency only)
    int32_t* hv;       // [hcap]
    int32_t* succ;     // [4*maxN]
    int32_t* pred;     // [4*maxN]
    uint8_t* deg;      // [maxN]  out-degree | in-degree << 4, cap edges included
    float* dist;       // [maxN]  distance from the source cap (+inf = unreachable)
    float* dist2;      // [maxN]  distance to the sink cap
    int32_t* before;   // [maxN]
    int32_t* after;    // [maxN]
    int32_t* hopF;     // [maxN]  edges between the source cap and the node along `before`
    int32_t* hopB;     // [maxN]  edges between the node and the sink cap along `after`
    int32_t* cand;     // [maxN]  open set of the forward pass; later the run list of the materialise step
    int32_t* cand2;    // [maxN]  open set of the backward pass
    uint32_t* bitsF;   // [maxN/32 + 2]  bit u: u -> u+1 is u's only out-edge and u+1's only in-edge (a reference step)
    uint32_t* bitsB;   // [maxN/32 + 2]  bit u: the same for u -> u-1 in the transposed graph
    uint8_t* eflag;    // [maxN]  bit c: edge to succ[c] still in edge_set; bit 4: cap edge
    int32_t* occ;      // [maxN]  nxtF during the tree phase, then occurrence counts of the solver
    int32_t* ce_a;     // [max_cand] edges that each yield one distinct alternative path
    int32_t* ce_b;
    int32_t* ce_len;
    int32_t* upath;    // [max_paths] candidate index of each unique path
    int32_t* pdiff;    // [4*max_paths] start, end_ref, end_var, end_ref_overlap
    int32_t* grp;      // [5*max_paths] cluster bookkeeping
    double* G;         // [max_cols*max_cols]
    double* V;         // [2*max_cols*max_cols]
    double* vec;       // [8*max_cols]
    unsigned long long* acc;  // [max_cols*max_cols + max_cols] exact integer accumulators of G and h
    PathView* cols;    // [max_cols] columns of the current least-squares problem
    int32_t* members;  // [max_cols]
    // The paths' node numbers, 16 bit, in shared memory (shared-memory passes only; pcache_cap = 0 otherwise).  Every scan
    // of a path -- ordering, diffs, the three passes of each row's solver, naming -- used to read R.pool, i.e. L2, one
    // dependent round trip per step of its loop; from here a step costs a shared-memory load.  Paths are cached in
    // order until the cache is full; ce_b[p] = offset of path p (rank order) in it, or -1.
    uint16_t* pcache;
    int pcache_cap;
    int maxN;
    int hcap;          // capacity of hk/hv (a power of two >= 2*maxN)
    int max_cand;      // capacity of the ce_* arrays
    int max_paths;
    int max_cols;
    int retry;         // 1: exceeding a capacity defers the target to the general pass
};

// The one place that knows the scratch layout (host sizing, device carving and the CPU emulation
// all use it).  compact = 1 (the shared-memory pass) overlays arrays with disjoint lifetimes:
//   kept  -> cand, keptk -> succ          numbering ends before either is written
//   hk,hv -> dist .. cand2                the adjacency map dies before the trees are initialised
//   hopF,hopB -> pdiff .. acc             hop counts die when the paths are materialised
// Each overlay is used only when it fits; otherwise the array gets storage of its own.
struct ScratchLayout {
    char* base;
    size_t stride;          // bytes per CTA
    size_t o_newidx, o_kept, o_keptk, o_hk, o_hv, o_succ, o_pred, o_deg, o_dist, o_dist2, o_before, o_after, o_hopF, o_hopB,
        o_cand, o_cand2, o_bitsF, o_bitsB, o_eflag, o_occ;
    size_t o_ce_a, o_ce_b, o_ce_len, o_upath, o_pdiff, o_grp, o_G, o_V, o_vec, o_acc, o_cols, o_members, o_pcache;
    int maxN, hcap, max_cand, max_paths, max_cols, pcache_cap;
};

KM_HOSTDEV ScratchLayout make_layout(int maxcap, int max_cand, int max_paths, int max_cols, int compact) {
    ScratchLayout L = {};
    const size_t maxN = (size_t)maxcap + 2, nce = (size_t)max_cand, np = (size_t)max_paths, nc = (size_t)max_cols;
    size_t hcap = 64;
    while (hcap < 2 * maxN) hcap <<= 1;
    size_t o = 0;
    auto put = [&](size_t bytes) { size_t at = o; o = (o + bytes + 15) / 16 * 16; return at; };
    L.o_newidx = put(4 * maxN);
    L.o_succ = put(16 * maxN); L.o_pred = put(16 * maxN); L.o_deg = put(maxN);
    L.o_dist = put(4 * maxN); L.o_dist2 = put(4 * maxN); L.o_before = put(4 * maxN); L.o_after = put(4 * maxN);
    L.o_cand = put(4 * maxN); L.o_cand2 = put(4 * maxN);
    const size_t tree_end = o;
    L.o_bitsF = put(4 * (maxN / 32 + 2)); L.o_bitsB = put(4 * (maxN / 32 + 2));
    L.o_eflag = put(maxN); L.o_occ = put(4 * maxN);
    L.o_ce_a = put(4 * nce); L.o_ce_b = put(4 * nce); L.o_ce_len = put(4 * nce);
    L.o_upath = put(4 * np);
    L.o_pdiff = put(16 * np); L.o_grp = put(20 * np);
    L.o_G = put(8 * nc * nc); L.o_V = put(16 * nc * nc); L.o_vec = put(64 * nc);
    L.o_acc = put(8 * (nc * nc + nc));
    const size_t quant_end = o;
    L.o_cols = put(sizeof(PathView) * nc); L.o_members = put(4 * nc);
    L.pcache_cap = compact ? KM_PCACHE : 0;
    L.o_pcache = put(2 * (size_t)L.pcache_cap);
    if (compact) { L.o_kept = L.o_cand; L.o_keptk = L.o_succ; }
    else { L.o_kept = put(4 * maxN); L.o_keptk = put(8 * maxN); }
    if (compact && 12 * hcap <= tree_end - L.o_dist) { L.o_hk = L.o_dist; L.o_hv = L.o_dist + 8 * hcap; }
    else { L.o_hk = put(8 * hcap); L.o_hv = put(4 * hcap); }
    if (compact && 8 * maxN + 16 <= quant_end - L.o_pdiff) { L.o_hopF = L.o_pdiff; L.o_hopB = L.o_pdiff + (4 * maxN + 15) / 16 * 16; }
    else { L.o_hopF = put(4 * maxN); L.o_hopB = put(4 * maxN); }
    L.stride = (o + 255) / 256 * 256;
    L.maxN = (int)maxN; L.hcap = (int)hcap; L.max_cand = max_cand; L.max_paths = max_paths; L.max_cols = max_cols;
    return L;
}

KM_HOSTDEV GraphScratch carve(const ScratchLayout& L, char* p, int retry) {
    GraphScratch S;
    S.newidx = (int32_t*)(p + L.o_newidx); S.kept = (int32_t*)(p + L.o_kept); S.keptk = (uint64_t*)(p + L.o_keptk);
    S.hk = (uint64_t*)(p + L.o_hk); S.hv = (int32_t*)(p + L.o_hv);
    S.succ = (int32_t*)(p + L.o_succ); S.pred = (int32_t*)(p + L.o_pred); S.deg = (uint8_t*)(p + L.o_deg);
    S.dist = (float*)(p + L.o_dist); S.dist2 = (float*)(p + L.o_dist2);
    S.before = (int32_t*)(p + L.o_before); S.after = (int32_t*)(p + L.o_after);
    S.hopF = (int32_t*)(p + L.o_hopF); S.hopB = (int32_t*)(p + L.o_hopB);
    S.cand = (int32_t*)(p + L.o_cand); S.cand2 = (int32_t*)(p + L.o_cand2); S.eflag = (uint8_t*)(p + L.o_eflag);
    S.bitsF = (uint32_t*)(p + L.o_bitsF); S.bitsB = (uint32_t*)(p + L.o_bitsB);
    S.occ = (int32_t*)(p + L.o_occ);
    S.ce_a = (int32_t*)(p + L.o_ce_a); S.ce_b = (int32_t*)(p + L.o_ce_b); S.ce_len = (int32_t*)(p + L.o_ce_len);
    S.upath = (int32_t*)(p + L.o_upath); S.pdiff = (int32_t*)(p + L.o_pdiff); S.grp = (int32_t*)(p + L.o_grp);
    S.G = (double*)(p + L.o_G); S.V = (double*)(p + L.o_V); S.vec = (double*)(p + L.o_vec);
    S.acc = (unsigned long long*)(p + L.o_acc);
    S.cols = (PathView*)(p + L.o_cols); S.members = (int32_t*)(p + L.o_members);
    S.pcache = (uint16_t*)(p + L.o_pcache); S.pcache_cap = L.pcache_cap;
    S.maxN = L.maxN; S.hcap = L.hcap; S.max_cand = L.max_cand; S.max_paths = L.max_paths; S.max_cols = L.max_cols; S.retry = retry;
    return S;
}

// hand target t to the general pass (its list is read after the shared-memory passes have finished)
KM_HD void defer_to_general(const ResultView& R, const WalkView& W, int t) {
    const int pos = atomic_addi32(&R.sched_count[2], 1);
    R.sched_order[2 * (size_t)W.n_targets + pos] = t;
}

#define KM_REF_W 0.01f
#define KM_ALT_W 1.0f
#define KM_NXT_REF 0x40000000          // nxt entry: the edge carries the reference weight
#define KM_NXT_MASK 0x3FFFFFFF

struct GraphDims {
    int L, N, src, snk;
};

// weight of edge i -> j (MutationFinder.py:512, 535-551)
KM_HD bool edge_is_ref(const GraphDims& d, int i, int j) {
    return i == d.src || j == d.snk || (i < d.L - 1 && j == i + 1);
}

struct alignas(16) Slot4 { int32_t v[4]; };

// Out-neighbours of u in the forward graph: up to 4 overlap successors + the cap edges.
// Calls f(j, slot) with slot 0..3 for overlap edges and 4 for the cap edge.
template <class F>
KM_HD void for_each_succ(const GraphScratch& S, const GraphDims& d, int u, F f) {
    if (u == d.src) { f(0, 4); return; }
    if (u == d.snk) return;
    const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * u);
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (s4.v[c] >= 0) f(s4.v[c], c);
    if (u == d.L - 1) f(d.snk, 4);
}

// k-mer -> canonical node index of this target (adjacency phase only)
KM_HD uint32_t node_hash(uint64_t key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 37); }

KM_HD int node_find(const GraphScratch& S, uint32_t hmask, uint64_t key) {
    uint32_t s = node_hash(key) & hmask;
    for (;;) {
        const uint64_t cur = S.hk[s];
        if (cur == key) return S.hv[s];
        if (cur == KM_EMPTY_KEY) return -1;
        s = (s + 1) & hmask;
    }
}

// number of consecutive set bits at u, u+1, ... / at u, u-1, ... (bits past the last node are zero)
KM_HD int run_up(const uint32_t* bits, int u) {
    int r = 0;
    for (;;) {
        const int sh = (u + r) & 31;
        const uint32_t inv = ~(bits[(u + r) >> 5] >> sh);       // the shifted-in top bits read as "not set"
        const int c = ffs32(inv) - 1;                            // inv != 0 whenever sh > 0
        if (inv != 0u && c < 32 - sh) return r + c;
        r += 32 - sh;
    }
}
KM_HD int run_down(const uint32_t* bits, int u) {
    int r = 0;
    for (;;) {
        if (u - r < 0) return r;
        const int sh = 31 - ((u - r) & 31);
        const uint32_t inv = ~(bits[(u - r) >> 5] << sh);
        const int c = inv ? clz32(inv) : 32;
        if (c < 32 - sh) return r + c;
        r += 32 - sh;
    }
}
KM_HD bool bit_at(const uint32_t* bits, int u) { return (bits[u >> 5] >> (u & 31)) & 1u; }

// Graph._get_paths (Graph.py:63-119) on adjacency lists; `forward` walks w, otherwise w
// transposed.  Every distance is the float32 sum of its parent's distance and one weight, so the
// pass is sequential and runs on ONE lane.  Three things keep the dependent chain short:
//   * nxt[u] >= 0 marks a node whose only out-edge leads to a node with no other in-edge.  If that
//     successor's distance is strictly below every open node's, it is the next node the reference
//     would settle (Graph.py:113-114) and nothing else changes -- almost every node of a target
//     graph is of this kind; the next nxt entry is fetched while the distance is being compared.
//   * a run of such steps along the reference (u -> u+1 -> ..., marked in `bits`) is taken in
//     registers: one float32 add, one compare and one store per node, no dependent load.  The
//     parents and hop counts of the nodes inside a run are filled in by the whole CTA afterwards
//     (fill_runs): they are determined by the run alone.
//   * otherwise the general step: the 4 neighbour slots come in with one 16-byte load, "unseen" is
//     dist == +inf, and the open set is an explicit list that rarely holds more than two nodes.
// Settling order = ascending (distance, index) and a node is re-parented only on a strictly smaller
// float32 sum (Graph.py:103), exactly as in the reference.
// dist/prev must be pre-filled with +inf / -1.
// (KM_TREE_TIMERS: per-step ticks inside the tree pass -- they cost more than the steps they time, so they are
// separate from the per-phase marks)
#if KM_DEVICE_BUILD && defined(KM_PHASE_TIMERS) && defined(KM_TREE_TIMERS)
#define KM_DBG_DECL long long dbg_t0 = 0;
#define KM_DBG_TICK(p) dbg_t0 = clock64();
#define KM_DBG_TOCK(p, n) { atomicAdd(&km_phase_cycles[p], (unsigned long long)(clock64() - dbg_t0)); atomicAdd(&km_phase_cycles[(p) + 8], (unsigned long long)(n)); }
#else
#define KM_DBG_DECL
#define KM_DBG_TICK(p)
#define KM_DBG_TOCK(p, n)
#endif
KM_HD void shortest_tree(const GraphScratch& S, const GraphDims& d, bool forward, float* dist, int32_t* prev, int32_t* hop,
                         int32_t* cand, const int32_t* nxt, const uint32_t* bits) {
    const int32_t* nbr = forward ? S.succ : S.pred;
    const int step = forward ? 1 : -1;
    int nc = 0;
    int u = forward ? d.src : d.snk;
    int hu = 0;
    float du = 0.0f, open_min = INFINITY;
    dist[u] = 0.0f;
    hop[u] = 0;
    int x = nxt[u];
    KM_DBG_DECL
    for (;;) {
        KM_DBG_TICK(forward ? 16 : 20)
        if (x >= 0) {
            if (bit_at(bits, u)) {
                const int r = forward ? run_up(bits, u) : run_down(bits, u);
                int done = 0;
                float t = du;
                // four steps at a time: the sums only grow, so testing the fourth against the open
                // set covers the three before it
                while (done + 4 <= r) {
                    const float t1 = add_f32(KM_REF_W, t), t2 = add_f32(KM_REF_W, t1), t3 = add_f32(KM_REF_W, t2),
                                t4 = add_f32(KM_REF_W, t3);
                    if (!(t4 < open_min)) break;
                    dist[u + step] = t1; dist[u + 2 * step] = t2; dist[u + 3 * step] = t3; dist[u + 4 * step] = t4;
                    t = t4; u += 4 * step; done += 4;
                }
                while (done < r) {
                    const float tn = add_f32(KM_REF_W, t);
                    if (!(tn < open_min)) break;
                    t = tn; u += step; dist[u] = t; ++done;
                }
                if (done) {
                    du = t; hu += done;
                    hop[u] = hu; prev[u] = u - step;
                    x = nxt[u];
                    KM_DBG_TOCK(forward ? 17 : 21, done)
                    continue;
                }
            } else {
                const int j = x & KM_NXT_MASK;
                const int xn = nxt[j];                                   // fetched early: off the critical path
                const float trial = add_f32((x & KM_NXT_REF) ? KM_REF_W : KM_ALT_W, du);
                if (trial < open_min) {
                    dist[j] = trial; prev[j] = u; hop[j] = ++hu;
                    u = j; du = trial; x = xn;
                    KM_DBG_TOCK(forward ? 18 : 22, 1)
                    continue;
                }
            }
        }
        auto relax = [&](int j, float w) {
            const float trial = add_f32(w, du);              // w[i, :] + dist[i] in float32 (:93)
            const float old = dist[j];
            if (trial < old) {                               // strict (:103)
                dist[j] = trial;
                prev[j] = u;
                hop[j] = hu + 1;
                if (old == INFINITY) cand[nc++] = j;         // first time seen -> joins the open set
            }
        };
        if (forward) {
            if (u == d.src) relax(0, KM_REF_W);              // BigBang -> first k-mer (:545-547)
            else if (u != d.snk) {
                const Slot4 s4 = *reinterpret_cast<const Slot4*>(nbr + 4 * u);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = s4.v[c];
                    if (j >= 0) relax(j, (j == u + 1 && u < d.L - 1) ? KM_REF_W : KM_ALT_W);
                }
                if (u == d.L - 1) relax(d.snk, KM_REF_W);    // last k-mer -> BigCrunch (:549-551)
            }
        } else {
            if (u == d.snk) relax(d.L - 1, KM_REF_W);
            else if (u != d.src) {
                const Slot4 s4 = *reinterpret_cast<const Slot4*>(nbr + 4 * u);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = s4.v[c];
                    if (j >= 0) relax(j, (u == j + 1 && j < d.L - 1) ? KM_REF_W : KM_ALT_W);
                }
                if (u == 0) relax(d.src, KM_REF_W);
            }
        }
        if (nc == 0) { KM_DBG_TOCK(forward ? 19 : 23, 1) break; }
        // open node of least distance, lowest index on ties (Graph.py:113-114)
        int best = 0;
        float db = dist[cand[0]];
        for (int c = 1; c < nc; ++c) {
            const int a = cand[c];
            const float da = dist[a];
            if (da < db || (da == db && a < cand[best])) { best = c; db = da; }
        }
        u = cand[best];
        du = db;
        hu = hop[u];
        cand[best] = cand[--nc];
        open_min = INFINITY;
        for (int c = 0; c < nc; ++c) { const float da = dist[cand[c]]; open_min = da < open_min ? da : open_min; }
        x = nxt[u];
        KM_DBG_TOCK(forward ? 19 : 23, 1)
    }
}

// Builds the graph of target t and emits its unique alternative paths (caps stripped), written
// to the result pool in lexicographic order; sh[2] = their number, sh[3] = the first path id.
// Returns false (uniformly) when a scratch capacity was exceeded and the target was deferred.
// All threads of the CTA must call this.
template <class Ctx>
KM_HD bool graph_target(const Ctx& ctx, const TableView& T, const WalkView& W, const GraphScratch& S,
                        const ResultView& R, int t, GraphDims* dims_out, int* sh /* 16 ints of CTA-shared memory */) {
    const int k = T.k;
    const TargetGeom g = target_geom(W, t, k);
    const int tid = ctx.tid(), nt = ctx.nt();
    const int n_all = W.n_nodes[t] < g.cap ? W.n_nodes[t] : g.cap;
    const int L = g.L;

    PhaseTimer pt;
    // ---- canonical numbering -------------------------------------------------
    if (tid == 0) sh[0] = 0;
    ctx.sync();
    for (int q = L + tid; q < n_all; q += nt) {
        S.newidx[q] = -1;
        if (W.node_slot[g.nbase + q] != KM_NO_SLOT) {            // the walk marks dropped nodes (walk.h)
            const int pos = atomic_addi32(&sh[0], 1);
            S.kept[pos] = q;
            S.keptk[pos] = W.node_kmer[g.nbase + q];
        }
    }
    ctx.sync();
    const int nk = sh[0];
    // rank kept novel nodes by packed k-mer value (all distinct)
    for (int a = tid; a < nk; a += nt) {
        const uint64_t ka = S.keptk[a];
        int rank = 0;
        for (int b = 0; b < nk; ++b) rank += S.keptk[b] < ka ? 1 : 0;
        S.newidx[S.kept[a]] = L + rank;
    }
    GraphDims d;
    d.L = L; d.N = L + nk + 2; d.src = d.N - 2; d.snk = d.N - 1;
    *dims_out = d;
    const int n_real = d.N - 2;
    uint32_t hmask = 63;
    while (hmask + 1 < 2u * (uint32_t)n_real) hmask = 2 * hmask + 1;
    ctx.sync();                                       // kept / keptk are dead from here on
    for (uint32_t s = tid; s <= hmask; s += nt) S.hk[s] = KM_EMPTY_KEY;
    if (tid == 0) R.t_n[t] = d.N;
    ctx.sync();
    // canonical node arrays (what the host sees) + the k-mer -> index map used for the edges
    for (int q = tid; q < n_all; q += nt) {
        const int i = q < L ? q : S.newidx[q];
        if (i < 0) continue;
        const uint64_t km = W.node_kmer[g.nbase + q];
        R.out_kmer[g.nbase + i] = km;
        R.out_count[g.nbase + i] = W.node_count[g.nbase + q];
        uint32_t s = node_hash(km) & hmask;
        while (atomic_cas64(&S.hk[s], KM_EMPTY_KEY, km) != KM_EMPTY_KEY) s = (s + 1) & hmask;   // keys are distinct
        S.hv[s] = i;
    }
    ctx.sync();

    pt.mark(0);
    // ---- adjacency (MutationFinder.py:515-531): edge i -> j iff kmer[i][1:] == kmer[j][:-1], i != j ------
    // Each node probes the map with its 4 successor k-mers; predecessor lists are filled by the
    // successors' owners (slot order is irrelevant to every later step).
    int32_t* indeg = S.hopB;                          // free until the tree phase
    for (int i = tid; i < n_real; i += nt) {
        Slot4 none; none.v[0] = none.v[1] = none.v[2] = none.v[3] = -1;
        *reinterpret_cast<Slot4*>(S.pred + 4 * i) = none;
        indeg[i] = 0;
    }
    for (int w = tid; w < (d.N >> 5) + 2; w += nt) { S.bitsF[w] = 0u; S.bitsB[w] = 0u; }
    ctx.sync();
    for (int i = tid; i < n_real; i += nt) {
        const uint64_t km = R.out_kmer[g.nbase + i];
        Slot4 sv;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int js = node_find(S, hmask, succ_kmer(km, c, T.kmask));
            if (js == i) js = -1;                       // `if i != j` (:530)
            sv.v[c] = js;
            if (js >= 0) S.pred[4 * js + atomic_addi32(&indeg[js], 1)] = i;
        }
        *reinterpret_cast<Slot4*>(S.succ + 4 * i) = sv;
    }
    ctx.sync();                                       // the k-mer map is dead from here on
    for (int i = tid; i < d.N; i += nt) {
        int od = 0, id = 0;
        if (i == d.src) od = 1;
        else if (i == d.snk) id = 1;
        else {
            const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * i);
            od = (s4.v[0] >= 0) + (s4.v[1] >= 0) + (s4.v[2] >= 0) + (s4.v[3] >= 0) + (i == L - 1);   // cap edges (:545-551)
            id = indeg[i] + (i == 0);
        }
        S.deg[i] = (uint8_t)(od | (id << 4));
    }
    ctx.sync();
    // ---- SIMPLE BUBBLE (most targets): the reference chain plus ONE chain of novel nodes that leaves it at node a and
    // rejoins it at node b -- downstream for a substitution, an insertion, a deletion; upstream (b <= a, the path then runs
    // through b..a twice) for a tandem duplication -- or no novel node at all.  Every non-reference
    // edge then lies on that one chain, so Graph.all_shortest (Graph.py:220-240) can only stitch two paths: the reference
    // (from the source cap's edge, which init_paths keeps, Graph.py:184-198) and 0..a + chain + b..L-1; both shortest-path
    // trees, the chain strip and the candidate edges are known without being computed, provided the trees would follow
    // the reference wherever there is a choice -- true while the by-passed stretch costs less than the chain
    // (0.01 per reference edge against 1.0 per other edge; the test below keeps a factor 2 clear of the inversion that
    // Graph.py:153-163 documents; upstream of a the reference is always the cheaper way).  Anything else -- branching, two
    // variants, extra overlaps between reference k-mers -- takes the general path below.
    {
        int* bub = sh + 24;                  // [0] branch nodes, [1] a, [2] head, [3] join nodes, [4] b, [5] tail, [6] violations
        if (tid < 8) bub[tid] = 0;
        ctx.sync();
        for (int i = tid; i < n_real; i += nt) {
            const int dg = S.deg[i], od = dg & 15, id = dg >> 4;
            const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * i);
            int bad = 0;
            if (i < L) {
                int novel_succ = -1, n_next = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = s4.v[c];
                    if (j < 0) continue;
                    if (j == i + 1) n_next += 1;
                    else if (j >= L) novel_succ = j;
                    else bad = 1;                                       // an overlap between two reference k-mers that are not neighbours
                }
                if (i < L - 1 && n_next != 1) bad = 1;
                if (od == 2 && novel_succ >= 0 && !bad) { atomic_addi32(&bub[0], 1); bub[1] = i; bub[2] = novel_succ; }
                else if (od != 1) bad = 1;
                if (id == 2) {
                    const Slot4 p4 = *reinterpret_cast<const Slot4*>(S.pred + 4 * i);
                    int novel_pred = -1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (p4.v[c] >= L) novel_pred = p4.v[c];
                    if (novel_pred >= 0) { atomic_addi32(&bub[3], 1); bub[4] = i; bub[5] = novel_pred; }
                    else bad = 1;
                } else if (id != 1) bad = 1;
            } else if (od != 1 || id != 1) bad = 1;                     // a novel node inside a simple chain
            if (bad) bub[6] = 1;
        }
        ctx.sync();
        if (tid == 0) {
            int simple = 0;
            if (!bub[6] && bub[0] == 0 && bub[3] == 0 && nk == 0) simple = 1;                 // the reference alone
            else if (!bub[6] && bub[0] == 1 && bub[3] == 1 && nk >= 1 && nk + 2 <= S.maxN && (bub[4] - bub[1]) < 50 * (nk + 1)) {
                // follow the chain: it must visit every novel node once and end in the join's predecessor
                int cur = bub[2], n = 0;
                while (cur >= L && n < nk) {
                    S.cand[n++] = cur;
                    const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * cur);
                    int nx = -1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (s4.v[c] >= 0) nx = s4.v[c];
                    if (nx == bub[4] && cur == bub[5]) { cur = -2; break; }
                    cur = nx;
                }
                if (cur == -2 && n == nk) simple = 2;
            }
            bub[7] = simple;
            if (simple) atomic_add64(&R.used[7], 1ull);                // (measurement: how many targets took this path)
        }
        ctx.sync();
    }
    const int simple = sh[24 + 7];
    const int bub_a = sh[24 + 1], bub_b = sh[24 + 4];
    if (simple) {
        if (tid == 0) {
            S.ce_len[0] = L + 2;                                              // nodes incl. both caps, like the general path
            if (simple == 2) S.ce_len[1] = (bub_a + 1) + nk + (L - bub_b) + 2;
            sh[1] = simple;                                                   // candidate count = number of paths
        }
        ctx.sync();
    } else {
    // simple edges: (u, j) with u's only out-edge and j's only in-edge (see shortest_tree)
    int32_t* nxtF = S.occ;
    int32_t* nxtB = S.newidx;
    for (int i = tid; i < d.N; i += nt) {
        S.eflag[i] = 0x1F;
        S.dist[i] = INFINITY; S.dist2[i] = INFINITY; S.before[i] = -1; S.after[i] = -1;
        int f = -1, b = -1;
        if (i == d.src) { if ((S.deg[0] >> 4) == 1) f = 0 | KM_NXT_REF; }
        else if (i == d.snk) { if ((S.deg[L - 1] & 15) == 1) b = (L - 1) | KM_NXT_REF; }
        else {
            const int dg = S.deg[i];
            if ((dg & 15) == 1) {
                int j = d.snk;                           // the cap edge unless an overlap successor exists
                const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * i);
#pragma unroll
                for (int c = 0; c < 4; ++c) if (s4.v[c] >= 0) j = s4.v[c];
                if ((S.deg[j] >> 4) == 1) {
                    f = j | (edge_is_ref(d, i, j) ? KM_NXT_REF : 0);
                    if (j == i + 1 && i < L - 1) atomic_or32(&S.bitsF[i >> 5], 1u << (i & 31));
                }
            }
            if ((dg >> 4) == 1) {
                int u = d.src;
                const Slot4 p4 = *reinterpret_cast<const Slot4*>(S.pred + 4 * i);
#pragma unroll
                for (int c = 0; c < 4; ++c) if (p4.v[c] >= 0) u = p4.v[c];
                if ((S.deg[u] & 15) == 1) {
                    b = u | (edge_is_ref(d, u, i) ? KM_NXT_REF : 0);
                    if (u == i - 1 && u < L - 1) atomic_or32(&S.bitsB[i >> 5], 1u << (i & 31));
                }
            }
        }
        nxtF[i] = f; nxtB[i] = b;
    }
    ctx.sync();

    pt.mark(1);
    // ---- two shortest-path trees (Graph.py:175-176), concurrently on two warps ------
    {
        const int lane_b = nt > 32 ? 32 : 0;                 // the backward pass's lane
        if (tid == 0) shortest_tree(S, d, true, S.dist, S.before, S.hopF, S.cand, nxtF, S.bitsF);
        if (tid == lane_b) shortest_tree(S, d, false, S.dist2, S.after, S.hopB, S.cand2, nxtB, S.bitsB);
    }
    ctx.sync();
    // parents and hop counts inside reference runs: node i entered over the step (i-1 -> i) of a
    // run has parent i-1 and lies (i - s) hops past the run's first node s
    for (int i = tid; i < L; i += nt) {
        if (i >= 1 && bit_at(S.bitsF, i - 1) && S.dist[i] < INFINITY) {
            const int r = run_down(S.bitsF, i - 1);
            S.before[i] = i - 1; S.hopF[i] = S.hopF[i - r] + r;
        }
        if (i + 1 < L && bit_at(S.bitsB, i + 1) && S.dist2[i] < INFINITY) {
            const int r = run_up(S.bitsB, i + 1);
            S.after[i] = i + 1; S.hopB[i] = S.hopB[i + r] + r;
        }
    }
    ctx.sync();
    pt.mark(2);
    // ---- strip the reference chain (Graph.py:178-198) ----------------------
    // The only out-edge of the source cap goes to node 0, so node 0 is the one start whose
    // predecessor is the source (`np.where(before == first_node)`, :184).  The walk follows
    // `after` from it and drops edge (last_cur, cur) from the second step on (`if last_cur and`:
    // None is falsy).  Nearly always that chain is 0,1,..,L-1,sink: checked by the whole CTA, and
    // then every lane strips its own edges; any other shape is walked by lane 0.
    {
        int bad = 0;
        for (int i = tid; i < L; i += nt) bad |= S.after[i] != (i == L - 1 ? d.snk : i + 1);
        const bool straight = !ctx.sync_or(bad);
        if (S.before[0] == d.src) {
            if (straight) {
                for (int last = 1 + tid; last < L; last += nt) {
                    if (last == L - 1) { S.eflag[last] &= (uint8_t)~0x10; continue; }
                    const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * last);
                    uint8_t m = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (s4.v[c] == last + 1) m |= (uint8_t)(1u << c);
                    if (m) S.eflag[last] &= (uint8_t)~m;
                }
            } else if (tid == 0) {
                int cur = 0, last = -1;
                for (int nxt = S.after[cur]; nxt != -1; nxt = S.after[cur]) {
                    cur = nxt;
                    if (last > 0) {                      // `if last_cur and ...`: None and 0 are falsy
                        if (cur == d.snk) { if (last == d.L - 1) S.eflag[last] &= (uint8_t)~0x10; }
                        else {
                            const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * last);
                            uint8_t m = 0;
                            for (int c = 0; c < 4; ++c) if (s4.v[c] == cur) m |= (uint8_t)(1u << c);
                            if (m) S.eflag[last] &= (uint8_t)~m;
                        }
                    }
                    last = cur;
                }
            }
        }
        if (tid == 0) sh[1] = 0;   // candidate count
    }
    ctx.sync();

    pt.mark(3);
    // ---- candidate edges (Graph.py:220-240) -----------------------------------
    // Edge (a, b) of the remaining edge_set yields the path src..a (forward tree) + b..snk
    // (backward tree) iff a is reachable from the source and b reaches the sink.  The reference
    // de-duplicates the resulting tuples in a set; here duplicates are recognised WITHOUT
    // building them: (a, b) repeats the path of (before[a], a) exactly when after[a] == b and that
    // earlier edge is itself still in the edge_set.  (The edges of one path that reproduce it are
    // contiguous -- a stripped reference edge can only follow them, never sit between two --
    // so the earliest one is the unique representative.)
    for (int a = tid; a < d.N; a += nt) {
        if (!(S.dist[a] < INFINITY)) continue;
        const uint8_t fl = S.eflag[a];
        for_each_succ(S, d, a, [&](int b, int slot) {
            if (!(fl & (1u << slot))) return;
            if (!(S.dist2[b] < INFINITY)) return;
            if (a != d.src && S.after[a] == b) {
                const int pa = S.before[a];
                bool earlier;
                if (pa == d.src) earlier = (S.eflag[pa] & 0x10) != 0;
                else {
                    const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * pa);
                    uint8_t m = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (s4.v[c] == a) m |= (uint8_t)(1u << c);
                    earlier = (S.eflag[pa] & m) != 0;
                }
                if (earlier) return;
            }
            const int pos = atomic_addi32(&sh[1], 1);
            if (pos < S.max_cand) {
                S.ce_a[pos] = a; S.ce_b[pos] = b;
                S.ce_len[pos] = S.hopF[a] + S.hopB[b] + 2;   // nodes incl. both caps
                S.upath[pos] = pos;
            }
        });
    }
    ctx.sync();
    }
    const int n_cand = sh[1];
    if (n_cand > S.max_cand || n_cand > S.max_paths) {
        if (tid == 0) {
            atomic_or32(&W.status[t], S.retry ? KM_ST_RETRY_LARGE : (uint32_t)KM_ST_TOO_MANY_COLS);
            if (S.retry) defer_to_general(R, W, t);
            R.t_n_paths[t] = 0; R.t_path_first[t] = 0; R.t_n_rows[t] = 0; R.t_row_first[t] = 0;
        }
        ctx.sync();
        return false;
    }

    pt.mark(4);
    // ---- allocate everything this target will write, in one place (lane 0) -------------------
    // path ids, pool ints, rows and spelled characters come from four independent atomics issued
    // back to back; rows are reserved by their upper bound (every path gives one vs_ref row and at
    // most one cluster row, MutationFinder.py:613-648, 758-811).
    if (tid == 0) {
        int nu = n_cand;
        int64_t total = 0;
        for (int c = 0; c < nu; ++c) total += S.ce_len[c] - 2;     // caps stripped (MutationFinder.py:562)
        int first = 0, first_row = 0;
        int64_t off = 0, soff = 0;
        bool overflow = false;
        if (nu > 0) {
            const unsigned long long a0 = atomic_add64(&R.used[0], (unsigned long long)nu);
            const unsigned long long a1 = atomic_add64(&R.used[1], (unsigned long long)total);
            const unsigned long long a2 = atomic_add64(&R.used[2], 2ull * (unsigned long long)nu);
            const unsigned long long a3 = R.seq_pool ? atomic_add64(&R.used[3], (unsigned long long)(total + (int64_t)nu * (k - 1))) : 0ull;
            first = (int)a0; off = (int64_t)a1; first_row = (int)a2; soff = (int64_t)a3;
            overflow = a0 + nu > (unsigned long long)R.path_cap || a1 + total > (unsigned long long)R.pool_cap ||
                       a2 + 2ull * nu > (unsigned long long)R.row_cap ||
                       (R.seq_pool && a3 + total + (int64_t)nu * (k - 1) > (unsigned long long)R.seq_cap);
        }
        if (!overflow) for (int u = 0; u < nu; ++u) {
            const int len = S.ce_len[u] - 2;
            R.path_off[first + u] = off;
            R.path_len[first + u] = len;
            off += len;
        }
        if (overflow) { atomic_or32(&W.status[t], KM_ST_PATH_OVERFLOW); nu = -1; first = 0; }
        R.t_n_paths[t] = nu < 0 ? 0 : nu;
        R.t_path_first[t] = first;
        if (nu <= 0) { R.t_n_rows[t] = 0; R.t_row_first[t] = 0; }
        sh[2] = nu;
        sh[3] = first;
        sh[6] = first_row;
        sh[10] = (int)(soff & 0x7FFFFFFF); sh[11] = (int)(soff >> 31);
    }
    ctx.sync();
    const int nu = sh[2], first = sh[3];
    if (nu < 0) return false;

    pt.mark(5);
    if (simple) {
        // the two paths written in place: the reference, and 0..a + chain (S.cand, in chain order) + b..L-1; they ARE in
        // lexicographic order (the chain's first node is numbered >= L > a + 1)
        int32_t* p0 = R.pool + R.path_off[first];
        for (int i = tid; i < L; i += nt) p0[i] = i;
        if (simple == 2) {
            int32_t* p1 = R.pool + R.path_off[first + 1];
            const int n1 = (bub_a + 1) + nk + (L - bub_b);
            for (int i = tid; i < n1; i += nt)
                p1[i] = i <= bub_a ? i : (i <= bub_a + nk ? S.cand[i - bub_a - 1] : bub_b + (i - bub_a - 1 - nk));
        }
        ctx.sync();
        if (tid == 0) {
            // ce_b codes for path_view (quant.h): -2 = the identity path 0..L-1, -3 = the bubble path (a, nk, b in ce_a[0..2],
            // its chain in S.cand, which nothing overwrites before the rows are done)
            S.ce_len[0] = L; S.ce_b[0] = -2;
            if (simple == 2) { S.ce_len[1] = (bub_a + 1) + nk + (L - bub_b); S.ce_b[1] = -3; S.ce_a[0] = bub_a; S.ce_a[1] = nk; S.ce_a[2] = bub_b; }
        }
        ctx.sync();
    } else {
    // ---- materialise: two lanes per unique path, one per chain; positions come from the hop counts.
    // A lane steps through novel nodes one dependent load at a time but crosses a reference run in
    // one go (the run bitmaps give its length); long runs are only recorded and then written by the
    // whole CTA.
    int32_t* runs = S.cand;                                  // 4 ints per entry: path, position, first node, length
    const int run_cap = S.maxN / 4;
    if (tid == 0) sh[9] = 0;
    ctx.sync();
    for (int w = tid; w < 2 * nu; w += nt) {
        const int u = w >> 1;
        int32_t* dst = R.pool + R.path_off[first + u];
        const int a = S.ce_a[u];
        auto emit = [&](int pos, int node, int len) {         // dst[pos + o] = node + o
            int e = run_cap;
            if (len >= 8) e = atomic_addi32(&sh[9], 1);
            if (e < run_cap) { runs[4 * e] = u; runs[4 * e + 1] = pos; runs[4 * e + 2] = node; runs[4 * e + 3] = len; }
            else for (int o = 0; o < len; ++o) dst[pos + o] = node + o;
        };
        if (!(w & 1)) {
            int p = S.hopF[a] - 1;                                   // source cap dropped
            int cur = a;
            while (cur != d.src) {
                if (cur >= 1 && cur < L && bit_at(S.bitsF, cur - 1)) {
                    const int r = run_down(S.bitsF, cur - 1);        // cur-r .. cur are consecutive ancestors
                    emit(p - r, cur - r, r + 1);
                    p -= r + 1;
                    cur = S.before[cur - r];
                } else { dst[p--] = cur; cur = S.before[cur]; }
            }
        } else {
            int p = S.hopF[a];
            int cur = S.ce_b[u];
            while (cur != d.snk) {
                if (cur + 1 < L && bit_at(S.bitsB, cur + 1)) {
                    const int r = run_up(S.bitsB, cur + 1);          // cur .. cur+r follow each other
                    emit(p, cur, r + 1);
                    p += r + 1;
                    cur = S.after[cur + r];
                } else { dst[p++] = cur; cur = S.after[cur]; }
            }
        }
    }
    ctx.sync();
    {
        const int n_runs = sh[9] < run_cap ? sh[9] : run_cap;
        for (int e = 0; e < n_runs; ++e) {
            int32_t* dst = R.pool + R.path_off[first + runs[4 * e]] + runs[4 * e + 1];
            const int node = runs[4 * e + 2], len = runs[4 * e + 3];
            for (int o = tid; o < len; o += nt) dst[o] = node + o;
        }
    }
    ctx.sync();

    pt.mark(6);
    // ---- the paths once more in shared memory, 16 bit (GraphScratch::pcache): ce_a[u] = offset of path u, -1 = not cached
    if (tid == 0) {
        int at = 0;
        for (int u = 0; u < nu; ++u) {
            const int len = S.ce_len[u] - 2;
            if (at + len <= S.pcache_cap) { S.ce_a[u] = at; at += len; }
            else S.ce_a[u] = -1;
        }
    }
    ctx.sync();
    for (int u = 0; u < nu; ++u) {
        const int at = S.ce_a[u];
        if (at < 0) continue;
        const int32_t* src = R.pool + R.path_off[first + u];
        const int len = S.ce_len[u] - 2;
        for (int i = tid; i < len; i += nt) S.pcache[at + i] = (uint16_t)src[i];
    }
    ctx.sync();
    // ---- lexicographic order (sorted(set of tuples)): rank by pairwise CTA-parallel compares ----
    if (nu > 1) {
        int* slot = sh + 8;
        for (int u = tid; u < nu; u += nt) S.upath[u] = 0;       // rank of each path
        ctx.sync();
        for (int x = 0; x < nu - 1; ++x)
            for (int y = x + 1; y < nu; ++y) {
                const int lx = S.ce_len[x] - 2, ly = S.ce_len[y] - 2, m = lx < ly ? lx : ly;
                const int cx = S.ce_a[x], cy = S.ce_a[y];
                const bool cached = cx >= 0 && cy >= 0;
                const int32_t* px = R.pool + (cached ? 0 : R.path_off[first + x]);
                const int32_t* py = R.pool + (cached ? 0 : R.path_off[first + y]);
                if (tid == 0) *slot = m;
                ctx.sync();
                if (cached) {
                    for (int q = tid; q < m; q += nt)
                        if (S.pcache[cx + q] != S.pcache[cy + q]) { atomic_mini32(slot, q); break; }
                } else {
                    for (int q = tid; q < m; q += nt)
                        if (px[q] != py[q]) { atomic_mini32(slot, q); break; }
                }
                ctx.sync();
                if (tid == 0) {
                    const int q = *slot;
                    const bool x_less = q < m ? (cached ? S.pcache[cx + q] < S.pcache[cy + q] : px[q] < py[q]) : lx < ly;
                    S.upath[x_less ? y : x] += 1;
                }
                ctx.sync();
            }
        if (tid == 0) {
            // permute (offset, length, cache offset) into rank order; pdiff / ce_b serve as temporaries
            for (int u = 0; u < nu; ++u) {
                const int64_t off = R.path_off[first + u];
                S.pdiff[4 * u] = (int32_t)(off & 0x7FFFFFFF); S.pdiff[4 * u + 1] = (int32_t)(off >> 31);
                S.pdiff[4 * u + 2] = S.ce_len[u] - 2; S.pdiff[4 * u + 3] = S.ce_a[u];
            }
            for (int u = 0; u < nu; ++u) {
                const int r = S.upath[u];
                R.path_off[first + r] = ((int64_t)S.pdiff[4 * u + 1] << 31) | (int64_t)S.pdiff[4 * u];
                R.path_len[first + r] = S.pdiff[4 * u + 2];
                S.ce_len[r] = S.pdiff[4 * u + 2];
                S.ce_b[r] = S.pdiff[4 * u + 3];
            }
        }
        ctx.sync();
    } else if (nu == 1) {
        if (tid == 0) { S.ce_b[0] = S.ce_a[0]; S.ce_len[0] = S.ce_len[0] - 2; }
        ctx.sync();
    }
    }
    // from here on: ce_len[p] = length and ce_b[p] = cache offset of path p in rank order
    pt.mark(7);
    // ---- spell every unique path once (MutationFinder.get_seq, :375-403): first k-mer, then the last
    // base of each following node; rows print slices of these strings
    if (R.seq_pool) {
        int64_t soff = ((int64_t)sh[11] << 31) | (int64_t)sh[10];
        for (int p = 0; p < nu; ++p) {
            const int len = S.ce_len[p];
            const int cat = S.ce_b[p];
            const int32_t* idx = R.pool + (cat >= 0 ? 0 : R.path_off[first + p]);
            if (tid == 0) R.path_seq_off[first + p] = soff;
            if (len > 0) {
                const uint64_t k0 = R.out_kmer[g.nbase + (cat >= 0 ? (int)S.pcache[cat] : idx[0])];
                for (int c = tid; c < len + k - 1; c += nt) {
                    int code;
                    if (c < k) code = (int)((k0 >> (2 * (k - 1 - c))) & 3ull);
                    else code = (int)(R.out_kmer[g.nbase + (cat >= 0 ? (int)S.pcache[cat + c - k + 1] : idx[c - k + 1])] & 3ull);
                    R.seq_pool[soff + c] = "ACGT"[code];
                }
            }
            soff += len + k - 1;
        }
    }
    ctx.sync();
    pt.mark(8);
    return sh[2] >= 0;
}

}  // namespace km
