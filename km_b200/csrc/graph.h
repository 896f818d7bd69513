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
#define KM_ST_RETRY_LARGE 0x40000000u   // internal: redo this target in the general pass

// a (possibly clipped) path: idx == nullptr means the reference path, whose node at
// position p is simply p
struct PathView {
    const int32_t* idx;
    int begin;
    int len;
};

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
    // [0] paths used, [1] pool ints used, [2] rows used, [3] sequence chars used
    unsigned long long* used;
};

// Per-CTA scratch in HBM (L2 resident in practice), sized for the largest target.
struct GraphScratch {
    int32_t* newidx;   // [maxcap]   discovery index -> canonical index, -1 dropped
    int32_t* kept;     // [maxcap]   discovery indices of kept novel nodes
    int32_t* succ;     // [4*maxN]
    int32_t* pred;     // [4*maxN]
    float* dist;       // [maxN]  distance from the source cap (+inf = unreachable)
    float* dist2;      // [maxN]  distance to the sink cap
    int32_t* before;   // [maxN]
    int32_t* after;    // [maxN]
    int32_t* cand;     // [maxN]  open set of the forward pass
    int32_t* cand2;    // [maxN]  open set of the backward pass
    uint8_t* eflag;    // [maxN]  bit c: edge to succ[c] still in edge_set; bit 4: cap edge
    int32_t* occ;      // [maxN]
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
    int maxN;
    int max_cand;      // capacity of the ce_* arrays
    int max_paths;
    int max_cols;
    int retry;         // 1: exceeding a capacity defers the target to the general pass
};

#define KM_REF_W 0.01f
#define KM_ALT_W 1.0f

struct GraphDims {
    int L, N, src, snk;
};

// weight of edge i -> j (MutationFinder.py:512, 535-551)
KM_HD float edge_weight(const GraphDims& d, int i, int j) {
    if (i == d.src || j == d.snk) return KM_REF_W;
    if (i < d.L - 1 && j == i + 1) return KM_REF_W;
    return KM_ALT_W;
}

// Out-neighbours of u in the forward graph: up to 4 overlap successors + the cap edges.
// Calls f(j, slot) with slot 0..3 for overlap edges and 4 for the cap edge.
template <class F>
KM_HD void for_each_succ(const GraphScratch& S, const GraphDims& d, int u, F f) {
    if (u == d.src) { f(0, 4); return; }
    if (u == d.snk) return;
    for (int c = 0; c < 4; ++c) {
        const int j = S.succ[4 * u + c];
        if (j >= 0) f(j, c);
    }
    if (u == d.L - 1) f(d.snk, 4);
}

template <class F>
KM_HD void for_each_pred(const GraphScratch& S, const GraphDims& d, int u, F f) {
    if (u == d.snk) { f(d.L - 1); return; }
    if (u == d.src) return;
    for (int c = 0; c < 4; ++c) {
        const int j = S.pred[4 * u + c];
        if (j >= 0) f(j);
    }
    if (u == 0) f(d.src);
}

// Graph._get_paths (Graph.py:63-119) on adjacency lists; `forward` walks w, otherwise w
// transposed.  Sequential by nature -- every distance is the float32 sum of its parent's distance
// and one weight -- so it runs on ONE lane, and the loop is kept to two dependent shared-memory
// round trips per settled node: the 4 neighbour slots come in with one 16-byte load, the current
// node and its distance live in registers, "unseen" is dist == +inf (no state array), and the
// open set is an explicit list that almost always holds a single node.
// dist/prev must be pre-filled with +inf / -1.
struct alignas(16) Slot4 { int32_t v[4]; };

KM_HD void shortest_tree(const GraphScratch& S, const GraphDims& d, bool forward, float* dist, int32_t* prev,
                         int32_t* cand) {
    const int32_t* nbr = forward ? S.succ : S.pred;
    int nc = 0;
    int u = forward ? d.src : d.snk;
    float du = 0.0f;
    dist[u] = 0.0f;
    for (;;) {
        auto relax = [&](int j, float w) {
            const float trial = add_f32(w, du);              // w[i, :] + dist[i] in float32 (:93)
            const float old = dist[j];
            if (trial < old) {                               // strict (:103)
                dist[j] = trial;
                prev[j] = u;
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
        if (nc == 0) break;
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
        cand[best] = cand[--nc];
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
    for (int q = tid; q < n_all; q += nt) {
        if (q < L) { S.newidx[q] = q; continue; }
        S.newidx[q] = -1;
        if (W.hflag[g.hbase + W.node_slot[g.nbase + q]]) {
            const int pos = atomic_addi32(&sh[0], 1);
            S.kept[pos] = q;
        }
    }
    ctx.sync();
    const int nk = sh[0];
    // rank kept novel nodes by packed k-mer value (all distinct)
    for (int a = tid; a < nk; a += nt) {
        const uint64_t ka = W.node_kmer[g.nbase + S.kept[a]];
        int rank = 0;
        for (int b = 0; b < nk; ++b) rank += W.node_kmer[g.nbase + S.kept[b]] < ka ? 1 : 0;
        S.newidx[S.kept[a]] = L + rank;
    }
    ctx.sync();
    GraphDims d;
    d.L = L; d.N = L + nk + 2; d.src = d.N - 2; d.snk = d.N - 1;
    *dims_out = d;
    for (int q = tid; q < n_all; q += nt) {
        const int i = S.newidx[q];
        if (i >= 0) {
            R.out_kmer[g.nbase + i] = W.node_kmer[g.nbase + q];
            R.out_count[g.nbase + i] = W.node_count[g.nbase + q];
        }
    }
    if (tid == 0) R.t_n[t] = d.N;
    ctx.sync();

    pt.mark(0);
    // ---- adjacency (MutationFinder.py:515-531) --------------------------------
    const int n_real = d.N - 2;
    for (int e = tid; e < 4 * n_real; e += nt) {
        const int i = e >> 2, c = e & 3;
        const uint64_t km = R.out_kmer[g.nbase + i];
        int js = -1, jp = -1;
        uint32_t s = visited_find(W, g, succ_kmer(km, c, T.kmask));
        if (s != KM_NO_SLOT && W.hflag[g.hbase + s]) js = S.newidx[W.hval[g.hbase + s]];
        if (js == i) js = -1;                       // `if i != j` (:530)
        s = visited_find(W, g, pred_kmer(km, c, k));
        if (s != KM_NO_SLOT && W.hflag[g.hbase + s]) jp = S.newidx[W.hval[g.hbase + s]];
        if (jp == i) jp = -1;
        S.succ[e] = js;
        S.pred[e] = jp;
    }
    for (int i = tid; i < d.N; i += nt) {
        S.eflag[i] = 0x1F;
        S.dist[i] = INFINITY; S.dist2[i] = INFINITY; S.before[i] = -1; S.after[i] = -1;
    }
    ctx.sync();

    pt.mark(1);
    // ---- two shortest-path trees (Graph.py:175-176), concurrently on two warps ------
    {
        const int lane_b = nt > 32 ? 32 : 0;                 // the backward pass's lane
        if (tid == 0) shortest_tree(S, d, true, S.dist, S.before, S.cand);
        if (tid == lane_b) shortest_tree(S, d, false, S.dist2, S.after, S.cand2);
    }
    ctx.sync();
    pt.mark(2);
    if (tid == 0) {
        // ---- strip the reference chain (Graph.py:178-198) ----------------------
        // the only out-edge of the source cap goes to node 0, so node 0 is the one start whose
        // predecessor is the source (`np.where(before == first_node)`, :184)
        if (S.before[0] == d.src) {
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
        sh[1] = 0;   // candidate count
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
        for_each_succ(S, d, a, [&](int b, int slot) {
            if (!(S.eflag[a] & (1u << slot))) return;
            if (!(S.dist2[b] < INFINITY)) return;
            if (a != d.src && S.after[a] == b) {
                const int pa = S.before[a];
                bool earlier;
                if (pa == d.src) earlier = (S.eflag[pa] & 0x10) != 0;
                else {
                    const Slot4 s4 = *reinterpret_cast<const Slot4*>(S.succ + 4 * pa);
                    uint8_t m = 0;
                    for (int c = 0; c < 4; ++c) if (s4.v[c] == a) m |= (uint8_t)(1u << c);
                    earlier = (S.eflag[pa] & m) != 0;
                }
                if (earlier) return;
            }
            const int pos = atomic_addi32(&sh[1], 1);
            if (pos < S.max_cand) { S.ce_a[pos] = a; S.ce_b[pos] = b; }
        });
    }
    ctx.sync();
    const int n_cand = sh[1];
    if (n_cand > S.max_cand || n_cand > S.max_paths) {
        if (tid == 0) {
            atomic_or32(&W.status[t], S.retry ? KM_ST_RETRY_LARGE : (uint32_t)KM_ST_TOO_MANY_COLS);
            R.t_n_paths[t] = 0; R.t_path_first[t] = 0; R.t_n_rows[t] = 0; R.t_row_first[t] = 0;
        }
        ctx.sync();
        return false;
    }
    // length of each unique path incl. both caps: hops to the source + hops to the sink
    for (int c = tid; c < n_cand; c += nt) {
        int len = 0;
        for (int cur = S.ce_a[c]; cur != -1; cur = S.before[cur]) ++len;
        for (int cur = S.ce_b[c]; cur != -1; cur = S.after[cur]) ++len;
        S.ce_len[c] = len;
        S.upath[c] = c;
    }
    ctx.sync();

    pt.mark(4);
    // ---- allocate path ids and pool space (lane 0) ---------------------------------------
    if (tid == 0) {
        int nu = n_cand;
        bool overflow = false;
        int64_t total = 0;
        for (int c = 0; c < nu; ++c) total += S.ce_len[c] - 2;     // caps stripped (MutationFinder.py:562)
        int first = 0;
        if (nu > 0) {
            first = (int)atomic_add64(&R.used[0], (unsigned long long)nu);
            if (first + nu > R.path_cap) overflow = true;
        }
        if (!overflow && nu > 0) {
            int64_t off = (int64_t)atomic_add64(&R.used[1], (unsigned long long)total);
            if (off + total > R.pool_cap) overflow = true;
            else for (int u = 0; u < nu; ++u) {
                const int len = S.ce_len[u] - 2;
                R.path_off[first + u] = off;
                R.path_len[first + u] = len;
                off += len;
            }
        }
        if (overflow) { atomic_or32(&W.status[t], KM_ST_PATH_OVERFLOW); nu = -1; first = 0; }
        R.t_n_paths[t] = nu < 0 ? 0 : nu;
        R.t_path_first[t] = first;
        if (nu < 0) { R.t_n_rows[t] = 0; R.t_row_first[t] = 0; }
        sh[2] = nu;
        sh[3] = first;
    }
    ctx.sync();
    const int nu = sh[2], first = sh[3];
    if (nu < 0) return false;

    pt.mark(5);
    // ---- materialise: one lane per unique path walks its two chains ------------------
    for (int u = tid; u < nu; u += nt) {
        const int c = S.upath[u];
        int32_t* dst = R.pool + R.path_off[first + u];
        int lf = 0;
        for (int cur = S.ce_a[c]; cur != -1; cur = S.before[cur]) ++lf;
        int p = lf - 2;                                          // source cap dropped
        for (int cur = S.ce_a[c]; cur != d.src; cur = S.before[cur]) dst[p--] = cur;
        p = lf - 1;
        for (int cur = S.ce_b[c]; cur != d.snk; cur = S.after[cur]) dst[p++] = cur;
    }
    ctx.sync();

    pt.mark(6);
    // ---- lexicographic order (sorted(set of tuples)): rank by pairwise CTA-parallel compares ----
    if (nu > 1) {
        int* slot = sh + 8;
        for (int u = tid; u < nu; u += nt) S.upath[u] = 0;       // rank of each path
        ctx.sync();
        for (int x = 0; x < nu - 1; ++x)
            for (int y = x + 1; y < nu; ++y) {
                const int32_t* px = R.pool + R.path_off[first + x];
                const int32_t* py = R.pool + R.path_off[first + y];
                const int lx = R.path_len[first + x], ly = R.path_len[first + y], m = lx < ly ? lx : ly;
                if (tid == 0) *slot = m;
                ctx.sync();
                for (int q = tid; q < m; q += nt)
                    if (px[q] != py[q]) { atomic_mini32(slot, q); break; }
                ctx.sync();
                if (tid == 0) {
                    const int q = *slot;
                    const bool x_less = q < m ? px[q] < py[q] : lx < ly;
                    S.upath[x_less ? y : x] += 1;
                }
                ctx.sync();
            }
        if (tid == 0) {
            // permute (offset, length) into rank order; pdiff / ce_len serve as temporaries
            for (int u = 0; u < nu; ++u) {
                const int64_t off = R.path_off[first + u];
                S.pdiff[2 * u] = (int32_t)(off & 0x7FFFFFFF); S.pdiff[2 * u + 1] = (int32_t)(off >> 31);
                S.ce_len[u] = R.path_len[first + u];
            }
            for (int u = 0; u < nu; ++u) {
                R.path_off[first + S.upath[u]] = ((int64_t)S.pdiff[2 * u + 1] << 31) | (int64_t)S.pdiff[2 * u];
                R.path_len[first + S.upath[u]] = S.ce_len[u];
            }
        }
    }
    ctx.sync();
    pt.mark(7);
    return sh[2] >= 0;
}

}  // namespace km
