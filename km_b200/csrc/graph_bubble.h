// Stages 2 and 3 of km_find_batch for the targets whose graph is a SIMPLE BUBBLE, by ONE WARP per target.
//
// Four of five targets of a panel have the graph graph.h calls a simple bubble: the reference chain plus one chain of
// novel nodes that leaves it at node a and rejoins it at node b (or no novel node at all).  For those,
// MutationFinder.graph_analysis (km/utils/MutationFinder.py:496-572) needs neither shortest-path tree, nor the chain
// strip, nor candidate edges (graph.h explains why); what is left -- numbering, overlap edges, the check itself, two
// paths, three rows -- is a string of short phases, each a few hundred cycles of work followed by a barrier.  Run by a
// CTA per target (km_graph_kernel) that string takes 60 k cycles per target and its 23..40 KB of scratch keep 5..8
// targets resident per SM: the pass is bound by how many targets are in flight, not by any unit of the SM.
//
// Here the same steps keep only what a bubble needs -- the k-mers, counts, a 16-bit k-mer -> node map, one successor
// and one predecessor per node: 8 KB for a 256-node target -- and run on one warp, so ~24 targets are in flight per SM
// and every barrier is a __syncwarp.  The bubble test is the one of graph.h, condition for condition; a target that
// fails it is appended to the work list of the CTA-per-target pass of its size class, which runs afterwards.  The rows
// (diff_path_without_overlap, get_name, PathQuant, clusters) are emit_rows of quant.h, instantiated for a warp.
#pragma once
#include "quant.h"

namespace km {

#define KM_BUB_NONE 0xFFFFu

// NODES = node capacity incl. the two caps (the size classes of find_config.h)
template <int NODES>
struct alignas(16) BubbleScratch {
    uint64_t km[NODES];          // canonical k-mers; dead once the paths are spelled (then: solver scratch pointers)
    uint32_t cnt[NODES];         // counts, canonical order
    uint32_t occ[NODES];         // occurrence counters of the solver; numbering: kept novel k-mers (with cand, 8 bytes each)
    int32_t cand[NODES];         // the chain: canonical numbers of the novel nodes in path order
    uint32_t slot[NODES];        // k-mer -> node map: 2 * NODES 16-bit slots (node + 1, 0 = empty), keys verified in km[]
    uint16_t nsucc[NODES];       // reference node: its novel successor; novel node: its (only) successor
    uint16_t npred[NODES];       // a novel predecessor of the node
    uint32_t seen[NODES / 2];    // presence filter in front of the map: 16 bits per node capacity, one bit per k-mer
    uint32_t indeg[NODES / 4];   // in-degrees over the edges that are not reference steps, one byte per node
    uint8_t info[NODES];         // out-degree | reference successors << 3 | bad << 5
    uint8_t last[NODES];         // last base of each node (naming)
    int32_t pdiff[8], grp[12], ce_a[4], ce_b[4], ce_len[4];
    int32_t sh[40];
};

template <int NODES>
KM_HD uint32_t bub_hash(uint64_t key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 37) & (2u * NODES - 1u); }

// (the filter bit of a k-mer: other bits of the same product)
template <int NODES>
KM_HD uint32_t bub_bit(uint64_t key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 18) & (16u * NODES - 1u); }

template <int NODES>
KM_HD int bub_find(const BubbleScratch<NODES>& B, uint64_t key) {
    const uint32_t f = bub_bit<NODES>(key);
    if (!((B.seen[f >> 5] >> (f & 31)) & 1u)) return -1;                // 15 of 16 absent k-mers end here
    uint32_t s = bub_hash<NODES>(key);
    for (;;) {
        const uint32_t v = (load_shared_volatile32(&B.slot[s >> 1]) >> ((s & 1) * 16)) & 0xFFFFu;
        if (v == 0) return -1;
        if (B.km[v - 1] == key) return (int)v - 1;
        s = (s + 1) & (2u * NODES - 1u);
    }
}
// the keys of one target are distinct: claim the first free slot
template <int NODES>
KM_HD void bub_insert(BubbleScratch<NODES>& B, uint64_t key, int idx) {
    uint32_t s = bub_hash<NODES>(key);
    for (;;) {
        uint32_t* w = &B.slot[s >> 1];
        const int sh = (s & 1) * 16;
        uint32_t old = load_shared_volatile32(w);
        while (((old >> sh) & 0xFFFFu) == 0) {
            const uint32_t seen = atomic_cas32(w, old, old | ((uint32_t)(idx + 1) << sh));
            if (seen == old) return;
            old = seen;
        }
        s = (s + 1) & (2u * NODES - 1u);
    }
}

// Target t by one group of threads (all call).  A target that is not a simple bubble is handed to the general pass (the
// scheduler sends here only targets whose walk never branched, so this is rare).  Returns false when it was handed on.
template <int NODES, class Ctx>
KM_HD bool bubble_target(const Ctx& ctx, const TableView& T, const WalkView& W, const ResultView& R, int t, BubbleScratch<NODES>& B) {
    const int k = T.k;
    const TargetGeom g = target_geom(W, t, k);
    const int tid = ctx.tid(), nt = ctx.nt();
    const int n_all = W.n_nodes[t] < g.cap ? W.n_nodes[t] : g.cap;
    const int L = g.L;
    int* sh = B.sh;
    auto hand_on = [&]() { if (tid == 0) defer_to_general(R, W, t); };
    if (n_all + 2 > NODES) { hand_on(); return false; }                 // (the scheduler sends only targets that fit)
    PhaseTimer pt;

    // ---- canonical numbering (graph.h): reference k-mers in place, kept novel k-mers by ascending value ------------------
    uint64_t* tmpk = reinterpret_cast<uint64_t*>(B.occ);                // occ + cand: 8 bytes per node
    uint16_t* tmpq = B.nsucc;
    if (tid == 0) sh[0] = 0;
    ctx.sync();
    for (int q = L + tid; q < n_all; q += nt) {
        if (W.node_slot[g.nbase + q] != KM_NO_SLOT) {                   // the walk marks dropped nodes (walk.h)
            const int pos = atomic_addi32(&sh[0], 1);
            tmpk[pos] = W.node_kmer[g.nbase + q];
            tmpq[pos] = (uint16_t)(q - L);
        }
    }
    for (int i = tid; i < L; i += nt) {
        const uint64_t km = W.node_kmer[g.nbase + i];
        const uint32_t c = W.node_count[g.nbase + i];
        B.km[i] = km; B.cnt[i] = c;
        R.out_kmer[g.nbase + i] = km; R.out_count[g.nbase + i] = c;
    }
    for (int w = tid; w < NODES; w += nt) B.slot[w] = 0u;
    for (int w = tid; w < NODES / 2; w += nt) B.seen[w] = 0u;
    for (int w = tid; w < NODES / 4; w += nt) B.indeg[w] = 0u;
    ctx.sync();
    const int nk = sh[0];
    for (int a = tid; a < nk; a += nt) {
        const uint64_t ka = tmpk[a];
        int rank = 0;
        for (int b = 0; b < nk; ++b) rank += tmpk[b] < ka ? 1 : 0;
        const int i = L + rank;
        const uint32_t c = W.node_count[g.nbase + L + tmpq[a]];
        B.km[i] = ka; B.cnt[i] = c;
        R.out_kmer[g.nbase + i] = ka; R.out_count[g.nbase + i] = c;
    }
    GraphDims d;
    d.L = L; d.N = L + nk + 2; d.src = d.N - 2; d.snk = d.N - 1;
    const int n_real = d.N - 2;
    if (tid == 0) R.t_n[t] = d.N;
    ctx.sync();                                                         // tmpk / tmpq are dead from here on
    pt.mark_warp(40);
    for (int i = tid; i < n_real; i += nt) {
        const uint64_t km = B.km[i];
        bub_insert<NODES>(B, km, i);
        const uint32_t f = bub_bit<NODES>(km);
        atomic_or32(&B.seen[f >> 5], 1u << (f & 31));
        B.npred[i] = KM_BUB_NONE;
    }
    if (tid < 8) sh[24 + tid] = 0;
    ctx.sync();

    pt.mark_warp(41);
    // ---- overlap edges (MutationFinder.py:515-531), kept only as far as the bubble test reads them --------------------------
    // Reference k-mer i + 1 IS a successor of reference k-mer i (consecutive windows of the target; the k-mers of a target
    // are distinct, so it is the only node with that k-mer): that edge is taken as read -- no lookup, no in-degree update
    // (the test below adds it back) -- and only the three other letters are asked for.
    for (int i = tid; i < n_real; i += nt) {
        const uint64_t km = B.km[i];
        const int ref_c = i < L - 1 ? (int)(B.km[i + 1] & 3ull) : -1;
        int od = ref_c >= 0 ? 1 : 0, n_next = od, bad = 0, novel = -1, any = -1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c == ref_c) continue;
            const int j = bub_find<NODES>(B, succ_kmer(km, c, T.kmask));
            if (j < 0 || j == i) continue;                              // `if i != j` (:530)
            od += 1; any = j;
            atomic_add32(&B.indeg[j >> 2], 1u << (8 * (j & 3)));
            if (i >= L) B.npred[j] = (uint16_t)i;
            if (i < L) {
                if (j >= L) novel = j;
                else bad = 1;                                           // an overlap between two reference k-mers that are not neighbours
            }
        }
        od += i == L - 1 ? 1 : 0;                                       // cap edge (:545-551)
        B.info[i] = (uint8_t)(od | (n_next << 3) | (bad << 5));
        const int keep = i < L ? novel : any;
        B.nsucc[i] = keep < 0 ? (uint16_t)KM_BUB_NONE : (uint16_t)keep;
    }
    ctx.sync();
    pt.mark_warp(42);
    // ---- the bubble test of graph.h ---------------------------------------------------------------------------------
    int* bub = sh + 24;                  // [0] branch nodes, [1] a, [2] head, [3] join nodes, [4] b, [5] tail, [6] violations
    for (int i = tid; i < n_real; i += nt) {
        const int inf = B.info[i], od = inf & 7, n_next = (inf >> 3) & 3;
        int bad = inf >> 5;
        const int id = (int)((B.indeg[i >> 2] >> (8 * (i & 3))) & 255u) + (i < L ? 1 : 0);   // + the reference step (or, node 0, the cap edge)
        if (i < L) {
            const int novel_succ = B.nsucc[i] == KM_BUB_NONE ? -1 : (int)B.nsucc[i];
            if (i < L - 1 && n_next != 1) bad = 1;
            if (od == 2 && novel_succ >= 0 && !bad) { atomic_addi32(&bub[0], 1); bub[1] = i; bub[2] = novel_succ; }
            else if (od != 1) bad = 1;
            if (id == 2) {
                const int novel_pred = B.npred[i] == KM_BUB_NONE ? -1 : (int)B.npred[i];
                if (novel_pred >= 0) { atomic_addi32(&bub[3], 1); bub[4] = i; bub[5] = novel_pred; }
                else bad = 1;
            } else if (id != 1) bad = 1;
        } else if (od != 1 || id != 1) bad = 1;                         // a novel node inside a simple chain
        if (bad) bub[6] = 1;
    }
    ctx.sync();
    if (tid == 0) {
        int simple = 0;
        if (!bub[6] && bub[0] == 0 && bub[3] == 0 && nk == 0) simple = 1;                 // the reference alone
        else if (!bub[6] && bub[0] == 1 && bub[3] == 1 && nk >= 1 && nk + 2 <= NODES && (bub[4] - bub[1]) < 50 * (nk + 1)) {
            // follow the chain: it must visit every novel node once and end in the join's predecessor
            int cur = bub[2], n = 0;
            while (cur >= L && n < nk) {
                B.cand[n++] = cur;
                const int nx = B.nsucc[cur] == KM_BUB_NONE ? -1 : (int)B.nsucc[cur];
                if (nx == bub[4] && cur == bub[5]) { cur = -2; break; }
                cur = nx;
            }
            if (cur == -2 && n == nk) simple = 2;
        }
        bub[7] = simple;
        if (simple) atomic_add64(&R.used[7], 1ull);                    // (measurement: how many targets took this path)
    }
    ctx.sync();
    const int simple = sh[24 + 7];
    const int bub_a = sh[24 + 1], bub_b = sh[24 + 4];
    pt.mark_warp(43);
    if (!simple) { hand_on(); return false; }

    // ---- allocate everything this target writes (as graph.h), the two paths, their spelling ----------------------------
    const int len1 = (bub_a + 1) + nk + (L - bub_b);
    if (tid == 0) {
        int nu = simple;
        const int64_t total = (int64_t)L + (simple == 2 ? len1 : 0);
        const unsigned long long a0 = atomic_add64(&R.used[0], (unsigned long long)nu);
        const unsigned long long a1 = atomic_add64(&R.used[1], (unsigned long long)total);
        const unsigned long long a2 = atomic_add64(&R.used[2], 2ull * (unsigned long long)nu);
        const unsigned long long a3 = R.seq_pool ? atomic_add64(&R.used[3], (unsigned long long)(total + (int64_t)nu * (k - 1))) : 0ull;
        int first = (int)a0, first_row = (int)a2;
        const int64_t off = (int64_t)a1, soff = (int64_t)a3;
        const bool overflow = a0 + nu > (unsigned long long)R.path_cap || a1 + total > (unsigned long long)R.pool_cap ||
                              a2 + 2ull * nu > (unsigned long long)R.row_cap ||
                              (R.seq_pool && a3 + total + (int64_t)nu * (k - 1) > (unsigned long long)R.seq_cap);
        if (!overflow) {
            R.path_off[first] = off; R.path_len[first] = L;
            if (simple == 2) { R.path_off[first + 1] = off + L; R.path_len[first + 1] = len1; }
        } else { atomic_or32(&W.status[t], KM_ST_PATH_OVERFLOW); nu = -1; first = 0; }
        R.t_n_paths[t] = nu < 0 ? 0 : nu;
        R.t_path_first[t] = first;
        if (nu <= 0) { R.t_n_rows[t] = 0; R.t_row_first[t] = 0; }
        sh[2] = nu; sh[3] = first; sh[6] = first_row;
        sh[10] = (int)(off & 0x7FFFFFFF); sh[11] = (int)(off >> 31);
        sh[12] = (int)(soff & 0x7FFFFFFF); sh[13] = (int)(soff >> 31);
    }
    ctx.sync();
    const int nu = sh[2], first = sh[3], first_row = sh[6];
    pt.mark_warp(44);
    if (nu < 0) return true;
    {
        // the reference, and 0..a + chain + b..L-1: they ARE in lexicographic order (the chain's first node is numbered
        // >= L > a + 1).  Spelled as MutationFinder.get_seq (:375-403): first k-mer, then the last base of each node.
        const int64_t off = ((int64_t)sh[11] << 31) | (int64_t)sh[10];
        int64_t soff = ((int64_t)sh[13] << 31) | (int64_t)sh[12];
        int32_t* p0 = R.pool + off;
        for (int i = tid; i < L; i += nt) p0[i] = i;
        int32_t* p1 = p0 + L;
        if (simple == 2)
            for (int i = tid; i < len1; i += nt)
                p1[i] = i <= bub_a ? i : (i <= bub_a + nk ? B.cand[i - bub_a - 1] : bub_b + (i - bub_a - 1 - nk));
        if (R.seq_pool) {
            const uint64_t k0 = B.km[0];                               // both paths start at node 0 (a >= 0)
            for (int p = 0; p < nu; ++p) {
                const int len = p == 0 ? L : len1;
                if (tid == 0) R.path_seq_off[first + p] = soff;
                for (int c = tid; c < len + k - 1; c += nt) {
                    int code;
                    if (c < k) code = (int)((k0 >> (2 * (k - 1 - c))) & 3ull);
                    else {
                        const int i = c - k + 1;
                        const int node = p == 0 ? i : (i <= bub_a ? i : (i <= bub_a + nk ? B.cand[i - bub_a - 1] : bub_b + (i - bub_a - 1 - nk)));
                        code = (int)(B.km[node] & 3ull);
                    }
                    R.seq_pool[soff + c] = "ACGT"[code];
                }
                soff += len + k - 1;
            }
        }
        for (int i = tid; i < n_real; i += nt) B.last[i] = (uint8_t)(B.km[i] & 3ull);
        if (tid == 0) {
            // codes for path_view (quant.h): ce_b -2 = the identity path 0..L-1, -3 = the bubble path (a, nk, b in ce_a[0..2])
            B.ce_len[0] = L; B.ce_b[0] = -2;
            if (simple == 2) { B.ce_len[1] = len1; B.ce_b[1] = -3; B.ce_a[0] = bub_a; B.ce_a[1] = nk; B.ce_a[2] = bub_b; }
        }
    }
    ctx.sync();

    pt.mark_warp(45);
    // ---- rows: quant.h on this warp -----------------------------------------------------------------------------------------
    GraphScratch S = {};
    S.occ = reinterpret_cast<int32_t*>(B.occ);
    S.cand = B.cand;
    S.pdiff = B.pdiff; S.grp = B.grp; S.ce_a = B.ce_a; S.ce_b = B.ce_b; S.ce_len = B.ce_len;
    // a cluster of several variants cannot come out of one bubble; should one ever, emit_rows hands the target to the general
    // pass before it touches the solver scratch (size + 1 > max_cols), so these only need to be valid addresses
    S.G = S.V = S.vec = reinterpret_cast<double*>(B.km);
    S.acc = reinterpret_cast<unsigned long long*>(B.km);
    S.cols = reinterpret_cast<PathView*>(B.km); S.members = reinterpret_cast<int32_t*>(B.km);
    S.pcache = nullptr; S.pcache_cap = 0;
    S.maxN = NODES; S.hcap = 0; S.max_cand = 2; S.max_paths = 2; S.max_cols = 2; S.retry = 1;
    emit_rows_prepared<Ctx, false>(ctx, T, W, S, R, t, d, nu, first, first_row, sh, B.last, B.cnt);
    pt.mark_warp(46);
    return true;
}

}  // namespace km
