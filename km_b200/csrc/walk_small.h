// Stage 1 of km_find_batch for targets of ordinary size: the same node discovery as walk.h
// (MutationFinder.__init__ / __extend, km/utils/MutationFinder.py:108-165, in the order-independent
// form of oracle/km_oracle.py:walk_closure), with the whole per-target state of ONE WARP in shared
// memory.  After the first level a walk has one or two live nodes per level and ~31..180 levels
// (SURVEY.md D3/H5), so its cost is the dependent chain per level; here that chain is one round of
// table loads from HBM plus shared-memory bookkeeping -- no visited set, counter or queue in L2.
//
//   * the target is kept 2-bit packed (16 bases per word); reference k-mer i is re-extracted from it
//     whenever a visited-set hit has to be verified, so the set stores only 16-bit node numbers
//   * visited set: open addressing over 16-bit slots (node index + 1), claimed with a 32-bit CAS on the
//     containing word; novel k-mers sit in a side array written before their slot is published
//   * level 0 issues, per reference k-mer, its own lookup and the three successors that leave the
//     reference; the fourth successor IS the next reference k-mer, whose count a neighbouring lane
//     has just fetched (Jellyfish.get_child would ask the table again, Jellyfish.py:61-66)
//   * lanes that reach the same child in one level elect a leader with match.any, the leader inserts,
//     the node number comes back by shuffle; (depth, breaks) is combined with atomicMin as in walk.h
//   * the peel (commit rule, MutationFinder.py:159-163) runs on shared-memory flags
// A target that does not fit (too long, or more novel nodes than KM_WS_NOVEL) is left to the general
// kernel: KM_ST_WALK_DEFER.
#pragma once
#include "walk.h"
#include "find_config.h"

namespace km {

#define KM_WS_HASH 1024      // visited-set slots per target (16 bit each)
#define KM_WS_NOVEL 128      // novel nodes per target
#define KM_WS_MAXL 448       // reference k-mers per target
#define KM_WS_SEQW 32        // packed words: (KM_WS_MAXL + 31 + 15) / 16 + 2 of padding
#define KM_ST_WALK_DEFER 0x20000000u   // internal: redo this target with the general walk kernel
#define KM_WS_KID_REF 255u

struct alignas(16) WalkSmall {
    uint64_t nk[KM_WS_NOVEL];          // novel k-mers, index = node - L
    uint32_t slot[KM_WS_HASH / 2];     // two 16-bit slots per word: node index + 1, 0 = empty
    uint32_t nmeta[KM_WS_NOVEL];       // (depth << 8) | breaks of novel nodes
    uint8_t kid[KM_WS_NOVEL][4];       // each accepted child of a novel node: 0 = none, KM_WS_KID_REF = a reference k-mer
                                       // (never peeled), else novel index + 1
    uint8_t alive[KM_WS_NOVEL];
    uint32_t seq2[KM_WS_SEQW];         // the target, 16 bases per word, first base in the top bits
    int32_t n_nodes;                   // next node index
    uint32_t flags;                    // status bits raised by any lane
};

KM_HD bool walk_small_fits(const TargetGeom& g) { return g.L >= 1 && g.L <= KM_WS_MAXL; }
KM_HD uint8_t ws_kid_code(int idx, int L) { return (uint8_t)(idx < L ? KM_WS_KID_REF : (uint32_t)(idx - L + 1)); }

// reference k-mer i from the packed target (k <= 31)
KM_HD uint64_t ws_ref_kmer(const WalkSmall& M, int i, int k) { return packed_kmer(M.seq2, i, k); }

KM_HD uint64_t ws_node_key(const WalkSmall& M, int idx, int L, int k) { return idx < L ? ws_ref_kmer(M, idx, k) : M.nk[idx - L]; }

KM_HD uint32_t ws_hash(uint64_t key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 40) & (KM_WS_HASH - 1); }

// node index of `key`, or -1
KM_HD int ws_find(const WalkSmall& M, uint64_t key, int L, int k) {
    uint32_t s = ws_hash(key);
    for (;;) {
        const uint32_t v = (load_shared_volatile32(&M.slot[s >> 1]) >> ((s & 1) * 16)) & 0xFFFFu;
        if (v == 0) return -1;
        if (ws_node_key(M, (int)v - 1, L, k) == key) return (int)v - 1;
        s = (s + 1) & (KM_WS_HASH - 1);
    }
}

// Insert node `idx` under `key` unless the key is already there; returns the node that holds the
// key afterwards.  For a novel node the caller has already written M.nk[idx - L].
KM_HD int ws_insert(WalkSmall& M, uint64_t key, int idx, int L, int k) {
    uint32_t s = ws_hash(key);
    for (;;) {
        uint32_t* w = &M.slot[s >> 1];
        const int sh = (s & 1) * 16;
        uint32_t old = load_shared_volatile32(w);
        for (;;) {
            const uint32_t v = (old >> sh) & 0xFFFFu;
            if (v) {
                if (ws_node_key(M, (int)v - 1, L, k) == key) return (int)v - 1;
                break;                                    // another key lives here: next slot
            }
            const uint32_t seen = atomic_cas32(w, old, old | ((uint32_t)(idx + 1) << sh));
            if (seen == old) return idx;
            old = seen;
        }
        s = (s + 1) & (KM_WS_HASH - 1);
    }
}

// Level 0 of every walk, flat over the reference k-mers of ALL targets (K3a).  One warp takes 32
// consecutive reference k-mers of one target: each lane fetches the count of its own k-mer
// (MutationFinder.py:111-112) and of the three successors that leave the reference
// (Jellyfish.get_child, Jellyfish.py:61-66); the fourth successor is the next reference k-mer, whose
// count arrives from the neighbouring lane by shuffle (the chunk's last lane fetches it itself).
// This is ~90 % of a panel's lookups and has no dependency between k-mers, so it runs at the
// random-sector rate of HBM instead of inside the latency-bound per-target walks.
// Writes node_kmer, node_count, the four successor counts (node_kid doubles as their store) and,
// for every reference k-mer with an accepted successor off the reference, one entry of the target's
// EXIT LIST: node_slot[0 .. n_kept) = position | accepted letters << 16 | branching << 20 (n_kept
// counts the entries until the walk kernel replaces it with the kept-node count).
template <class Ctx, bool LINKED>
KM_HD void ref_probe_chunk(const Ctx& ctx, const TableView& T, const WalkView& W, const FindParams& P, int t, int i0) {
    const int k = T.k;
    const TargetGeom g = target_geom(W, t, k);
    const int lane = ctx.tid(), nl = ctx.nt();
    const int i = i0 + lane;
    const bool active = i < g.L;
    uint64_t q[5] = {0, 0, 0, 0, 0};
    uint32_t r[5];
    int ref_c = -1;
    uint32_t mask = 0;
    if (active) {
        const uint32_t* words = W.pack + W.pack_off[t];
        const uint64_t v = packed_kmer(words, i, k);
        if (i + 1 < g.L) ref_c = packed_base(words, i + k);
        q[0] = v;
        mask = 1u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            q[1 + c] = succ_kmer(v, c, T.kmask);
            if (c != ref_c || lane == nl - 1) mask |= 2u << c;
        }
    }
    uint32_t n_issued = (uint32_t)popc32(mask);
    if (LINKED) {
        // the k-mer's own record says which of its successors exist at all (neighbour mask, table.h): an absent one is
        // count 0 without a read, so a reference k-mer costs ~1 table read instead of ~4
        uint32_t own = 0, succ = 15u;
        if (active) table_query_links(T, q[0], &own, &succ);
        const uint32_t need = (mask >> 1) & succ;
        r[0] = own;
        // (a successor that exists off the reference is rare -- a variant site: no key is even formed for the others)
#pragma unroll
        for (int c = 0; c < 4; ++c) r[1 + c] = ((need >> c) & 1u) ? table_query(T, q[1 + c]) : 0u;
        n_issued = active ? 1u + (uint32_t)popc32(need) : 0u;
    } else {
        table_query_family_warp<5>(T, family_of_suffix(T, q[0]), q, mask, r);
    }
    const uint32_t next_own = (uint32_t)warp_shfl_down32((int)r[0], 1);
    if (active) {
        if (lane != nl - 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (c == ref_c) r[1 + c] = next_own;      // (no dynamic index: r stays in registers)
        }
        W.node_kmer[g.nbase + i] = q[0];
        W.node_count[g.nbase + i] = r[0];
        uint32_t* cc = W.node_kid + 4 * (g.nbase + i);
#if KM_DEVICE_BUILD
        *reinterpret_cast<uint4*>(cc) = make_uint4(r[1], r[2], r[3], r[4]);          // (node_kid is 256-byte aligned, 16 bytes per node)
#else
        cc[0] = r[1]; cc[1] = r[2]; cc[2] = r[3]; cc[3] = r[4];
#endif
        // Jellyfish.py:61-72 -- Python int * float, then max with the int floor, then >=
        const uint64_t sum = (uint64_t)r[1] + r[2] + r[3] + r[4];
        double thr = (double)sum * P.ratio;
        if (thr < (double)P.count) thr = (double)P.count;
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc |= ((double)r[1 + c] >= thr) ? (1u << c) : 0u;
        const int nkid = popc32(acc);
        // a reference k-mer starts its walk at depth 1 with no branching behind it (MutationFinder.py:115-120)
        bool limit = 1 > P.max_stack;                                        // :140-141
        if (!limit && nkid > 1 && 1 > P.max_break) limit = true;             // :153-156
        if (limit) atomic_or32(&W.status[t], KM_ST_TOUCHED_LIMIT);
        else {
            const uint32_t off_ref = ref_c >= 0 ? acc & ~(1u << ref_c) : acc;
            if (off_ref) {
                const int pos = atomic_addi32(&W.n_kept[t], 1);
                W.node_slot[g.nbase + pos] = (uint32_t)i | (off_ref << 16) | ((nkid > 1 ? 1u : 0u) << 20);
            }
        }
    }
    const uint32_t n = warp_add32(n_issued);          // (at most 5 per lane: 32 bits, one instruction instead of ten shuffles)
    if (lane == 0 && n) atomic_add64(&W.lookups[t], (unsigned long long)n);
}

// The four successor lookups of a novel node.
KM_HD void ws_query_children(const TableView& T, uint64_t kmer, uint64_t (&ck)[4], uint32_t (&cc)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) ck[c] = succ_kmer(kmer, c, T.kmask);
    table_query_family<4>(T, family_of_suffix(T, kmer), ck, 15u, cc);
}

// Children of one level, one successor letter per call: lanes holding the same child k-mer find each
// other with one match.any, the lowest lane inserts (or finds) the node, its number comes back by
// shuffle; (depth, breaks) is combined with atomicMin.  `parent` >= L records the child link used by
// the peel.  All lanes of the warp must call.
template <class Ctx>
KM_HD void ws_child(const Ctx& ctx, const WalkView& W, const TargetGeom& g, WalkSmall& M, int k, int novel_cap, bool has,
                    uint64_t child, uint32_t count, uint32_t child_meta, int parent, int c) {
    const int lane = ctx.tid(), L = g.L;
    const uint64_t mk = has ? child : (0x8000000000000000ull | (uint64_t)lane);
    const uint32_t peers = warp_match64(mk);
    const int leader = ffs32(peers) - 1;
    int idx = -1;
    if (has && leader == lane) {
        idx = ws_find(M, child, L, k);
        if (idx < 0) {
            const int fresh = atomic_addi32(&M.n_nodes, 1);
            if (fresh - L >= novel_cap) { atomic_or32(&M.flags, 1u); idx = -2; }
            else {
                M.nk[fresh - L] = child;
                M.nmeta[fresh - L] = 0xFFFFFFFFu;
                M.alive[fresh - L] = 1;
                fence_block();                      // the k-mer is visible before its slot is
                idx = ws_insert(M, child, fresh, L, k);
                W.node_kmer[g.nbase + fresh] = child;
                W.node_count[g.nbase + fresh] = count;
            }
        }
    }
    idx = warp_shfl32(idx, leader);
    if (has && idx >= 0) {
        if (idx >= L) atomic_min32(&M.nmeta[idx - L], child_meta);
        if (parent >= L) M.kid[parent - L][c] = ws_kid_code(idx, L);
    }
}

#if KM_DEVICE_BUILD
// node index of `key` or -1; *empty = the slot where the probe ended (where the key would go)
KM_HD int ws_find_slot(const WalkSmall& M, uint64_t key, int L, int k, uint32_t* empty) {
    uint32_t s = ws_hash(key);
    for (;;) {
        const uint32_t v = (load_shared_volatile32(&M.slot[s >> 1]) >> ((s & 1) * 16)) & 0xFFFFu;
        if (v == 0) { *empty = s; return -1; }
        if (ws_node_key(M, (int)v - 1, L, k) == key) return (int)v - 1;
        s = (s + 1) & (KM_WS_HASH - 1);
    }
}

// Counts prefetched one level ahead (ws_chain_level): valid on lanes 0..3 for the children of `node`.
struct ChainPrefetch { int node; uint32_t cnt; };

// One level whose frontier is a SINGLE novel node q -- what a walk looks like for most of its life (a
// variant is a chain of novel k-mers).  Same rules as the general level below (Jellyfish.py:61-72,
// MutationFinder.py:140-163), but with one node there is nothing to elect or to combine: every lane
// follows the same control flow on the same shared-memory words, lane 0 does the stores, and no atomic,
// match or fence is needed.  The table is asked one level AHEAD as well: lanes 0..3 fetch the four
// children, lanes 4..19 the sixteen grandchildren in the same round trip, so when the level yields exactly
// one new node the next level starts with its counts already in registers (`pf`) -- one HBM latency per
// two levels of the chain.
KM_HD void ws_chain_level(const TableView& T, const WalkView& W, const FindParams& P, const TargetGeom& g, WalkSmall& M,
                          int novel_cap, int q, ChainPrefetch& pf, uint32_t& st, unsigned& nlook) {
    const int lane = threadIdx.x & 31, L = g.L, k = T.k;
    const uint32_t meta = M.nmeta[q - L];
    const int depth = (int)(meta >> 8), breaks = (int)(meta & 255u);
    if (depth > P.max_stack) {                                                   // MutationFinder.py:140-141
        st |= KM_ST_TOUCHED_LIMIT;
        if (lane < 4) M.kid[q - L][lane] = 0;
        pf.node = -1;
        return;
    }
    const uint64_t parent = M.nk[q - L];
    const int c = lane & 3;
    uint32_t cnt, ahead = 0;
    const bool loaded = pf.node != q;
    if (loaded) {
        const bool live = lane < 20;
        const uint64_t mine = lane < 4 ? parent : succ_kmer(parent, (lane - 4) >> 2, T.kmask);
        const uint64_t ck = succ_kmer(mine, c, T.kmask);
        uint32_t r = 0;
        if (T.lines) r = quad_line_query(T, family_of_suffix(T, mine), T.canonical ? canonical(ck, T.k) : ck, live);
        else if (live) r = table_query(T, ck);
        nlook += live ? 1u : 0u;
        cnt = r; ahead = r;
    } else {
        cnt = pf.cnt;
    }
    // Jellyfish.py:61-72 over the four counts (every lane computes the same)
    uint32_t cc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cc[j] = (uint32_t)__shfl_sync(0xFFFFFFFFu, (int)cnt, j);
    const uint64_t sum = (uint64_t)cc[0] + cc[1] + cc[2] + cc[3];
    double thr = (double)sum * P.ratio;
    if (thr < (double)P.count) thr = (double)P.count;
    uint32_t pass = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) pass |= ((double)cc[j] >= thr) ? (1u << j) : 0u;
    int nb = breaks;
    if (popc32(pass) > 1) {                                                      // MutationFinder.py:153-156
        st |= KM_ST_BRANCHED;
        nb += 1;
        if (nb > P.max_break) { st |= KM_ST_TOUCHED_LIMIT; pass = 0; }
    }
    const uint32_t child_meta = pack_meta(depth + 1, nb);
    int nn = (int)load_shared_volatile32(reinterpret_cast<const uint32_t*>(&M.n_nodes));
    int n_new = 0, new_c = 0, new_idx = 0;
    uint8_t kid[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!((pass >> j) & 1u)) continue;
        const uint64_t child = succ_kmer(parent, j, T.kmask);
        uint32_t slot = 0;
        int idx = ws_find_slot(M, child, L, k, &slot);
        if (idx < 0) {
            if (nn - L >= novel_cap) {                                           // as ws_child: the walk is redone elsewhere
                if (lane == 0) M.flags |= 1u;
                break;
            }
            idx = nn++;
            if (lane == 0) {
                M.nk[idx - L] = child;
                M.nmeta[idx - L] = child_meta;
                M.alive[idx - L] = 1;
                M.slot[slot >> 1] |= (uint32_t)(idx + 1) << ((slot & 1) * 16);
                W.node_kmer[g.nbase + idx] = child;
                W.node_count[g.nbase + idx] = cc[j];
            }
            __syncwarp();
            n_new += 1; new_c = j; new_idx = idx;
        } else if (idx >= L) {
            if (lane == 0 && child_meta < M.nmeta[idx - L]) M.nmeta[idx - L] = child_meta;
            __syncwarp();
        }
        kid[j] = ws_kid_code(idx, L);
    }
    if (lane == 0) {
        M.kid[q - L][0] = kid[0]; M.kid[q - L][1] = kid[1]; M.kid[q - L][2] = kid[2]; M.kid[q - L][3] = kid[3];
        M.n_nodes = nn;
    }
    // the counts of the new node's children came back with this level's loads
    const uint32_t next = (uint32_t)__shfl_sync(0xFFFFFFFFu, (int)ahead, 4 + 4 * new_c + c);
    if (loaded && n_new == 1) { pf.node = new_idx; pf.cnt = next; }
    else pf.node = -1;
}
#endif

// The start of a target's walk, by one warp: clears the set, registers the reference k-mers (phase 1) and inserts the
// successors that leave the reference from the exit list ref_probe_chunk wrote (level 0).  Returns false when the
// target is finished already (malformed: its status is written and nothing is walked).
template <class Ctx>
KM_HD bool ws_begin(const Ctx& ctx, const TableView& T, const WalkView& W, const FindParams& P, int t, const TargetGeom& g,
                    WalkSmall& M, int novel_cap, PhaseTimer& pt) {
    const int k = T.k;
    const int lane = ctx.tid(), nl = ctx.nt();
    const int L = g.L;
    const int n_exits = W.n_kept[t];               // entries of the exit list (ref_probe_chunk)
    uint32_t st = 0;
    (void)P;

    // ---- set-up: clear the set, pack the target ---------------------------------------------
    for (int s = lane; s < KM_WS_HASH / 2; s += nl) M.slot[s] = 0u;
    {
        const int64_t w0 = W.pack_off[t];
        const int nw = (int)(W.pack_off[t + 1] - w0);
        for (int w = lane; w < KM_WS_SEQW; w += nl) M.seq2[w] = w < nw ? W.pack[w0 + w] : 0u;
    }
    if (W.pre_bad[t]) st |= KM_ST_BAD_BASE;
    if (lane == 0) {
        M.n_nodes = L; M.flags = 0;
        if (n_exits > 1 || (n_exits == 1 && popc32((W.node_slot[g.nbase] >> 16) & 15u) > 1)) atomic_or32(&W.status[t], KM_ST_BRANCHED);
    }
    ctx.sync();

    pt.mark_warp(32);
    // ---- phase 1 (MutationFinder.py:111-112): register the reference k-mers (their counts and those of
    // their successors were fetched by ref_probe_chunk) --------------------------------------------------
    for (int i = lane; i < L; i += nl)
        if (ws_insert(M, ws_ref_kmer(M, i, k), i, L, k) != i) st |= KM_ST_DUP_KMER;         // common.py:55-59
    // a malformed target stops here: the host raises before any walk (common.py:55-59)
    if (ctx.sync_or((st & (KM_ST_BAD_BASE | KM_ST_DUP_KMER)) != 0)) {
        if (st) atomic_or32(&W.status[t], st);
        ctx.sync();
        if (lane == 0) { W.n_nodes[t] = L; W.n_kept[t] = 0; }
        return false;
    }
    ctx.sync();
    pt.mark_warp(33);

    // ---- level 0: the successors that leave the reference, from the exit list -------------------------
    for (int base = 0; base < n_exits; base += nl) {
        const int e = base + lane;
        const bool active = e < n_exits;
        uint32_t entry = 0;
        uint64_t kmer = 0;
        const uint32_t* rc = W.node_kid;
        if (active) {
            entry = W.node_slot[g.nbase + e];
            kmer = ws_ref_kmer(M, (int)(entry & 0xFFFFu), k);
            rc = W.node_kid + 4 * (g.nbase + (int64_t)(entry & 0xFFFFu));
        }
        const uint32_t child_meta = pack_meta(2, (int)((entry >> 20) & 1u));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool has = active && ((entry >> (16 + c)) & 1u);
            ws_child(ctx, W, g, M, k, novel_cap, has, succ_kmer(kmer, c, T.kmask), has ? rc[c] : 0u, child_meta, -1, c);
        }
    }
    ctx.sync();
    pt.mark_warp(34);
    return true;
}

// The end of a target's walk, by one warp: the peel (commit rule, MutationFinder.py:159-163) over the n_all nodes the
// levels discovered, then the per-target results.  `st` / `nlook`: lane-private status bits and lookup counts.
// Returns false when the target was deferred to the general kernel (more novel nodes than fit here).
template <class Ctx>
KM_HD bool ws_finish(const Ctx& ctx, const WalkView& W, const FindParams& P, int t, const TargetGeom& g, WalkSmall& M,
                     int novel_cap, int n_all, uint32_t st, unsigned nlook, PhaseTimer& pt) {
    const int lane = ctx.tid(), nl = ctx.nt();
    const int L = g.L;
    if (load_shared_volatile32(&M.flags) & 1u) {
        // more novel nodes than fit here: a capacity the host can raise, or the general kernel's job
        if (lane == 0) {
            if (novel_cap < KM_WS_NOVEL) { atomic_or32(&W.status[t], KM_ST_NODE_OVERFLOW); W.n_nodes[t] = g.cap + 1; W.n_kept[t] = 0; }
            else atomic_or32(&W.status[t], KM_ST_WALK_DEFER);
        }
        return novel_cap < KM_WS_NOVEL;
    }

    // ---- phase 3: peel novel nodes with no surviving accepted child (commit rule, :159-163) -----------
    int changed = 1;
    while (changed) {
        int mine = 0;
        for (int q = L + lane; q < n_all; q += nl) {
            if (!load_shared_volatile8(&M.alive[q - L])) continue;
            bool ok = false;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t kd = M.kid[q - L][c];
                if (kd && (kd == KM_WS_KID_REF || load_shared_volatile8(&M.alive[kd - 1]))) ok = true;
            }
            if (!ok) { M.alive[q - L] = 0; mine = 1; }
        }
        changed = ctx.sync_or(mine);
    }

    pt.mark_warp(36);
    // ---- results: kept / dropped per node, counts, status ---------------------------------------------
    int kept = 0;
    for (int q = L + lane; q < n_all; q += nl) {
        const bool a = M.alive[q - L] != 0;
        W.node_slot[g.nbase + q] = a ? 0u : KM_NO_SLOT;
        kept += a ? 1 : 0;
    }
    kept = (int)warp_sum64((unsigned long long)kept);
    st = warp_or32(st);
    nlook = (unsigned)warp_sum64((unsigned long long)nlook);
    if (lane == 0) {
        const int total = L + kept;
        W.n_nodes[t] = n_all;
        W.n_kept[t] = total;
        if (total > P.max_node) st |= KM_ST_NODE_LIMIT;                          // MutationFinder.py:143-148
        // the status is final here (the probe kernel is done, the other lanes' bits are in `st`): the walk's internal hint
        // goes to the scheduler's code instead of the status, and the scheduler need not read five arrays per target
        const uint32_t all = atomic_or32(&W.status[t], st) | st;
        if (W.sched_code) {
            const int cap_all = n_all < g.cap ? n_all : g.cap;
            W.sched_code[t] = (uint16_t)(sched_code_of(all, cap_all, total, KM_TINY_NODES, KM_SMALL_NODES) + 1u);
            if (all & KM_ST_BRANCHED) W.status[t] = all & ~KM_ST_BRANCHED;
        }
        if (nlook) atomic_add64(&W.lookups[t], nlook);
    }
    pt.mark_warp(37);
    return true;
}

// One warp walks target t.  Returns false when the target was deferred to the general kernel.
template <class Ctx>
KM_HD bool walk_small_target(const Ctx& ctx, const TableView& T, const WalkView& W, const FindParams& P, int t, WalkSmall& M) {
    const int k = T.k;
    const TargetGeom g = target_geom(W, t, k);
    const int lane = ctx.tid(), nl = ctx.nt();
    const int L = g.L;
    const int novel_cap = g.cap - L < KM_WS_NOVEL ? g.cap - L : KM_WS_NOVEL;
    unsigned nlook = 0;
    uint32_t st = 0;
    (void)k; (void)nl; (void)lane;

    PhaseTimer pt;
    if (!ws_begin(ctx, T, W, P, t, g, M, novel_cap, pt)) return true;

    // ---- later levels: novel nodes only, level-synchronous ---------------------------------------------
    int lo = L;
    int hi = (int)load_shared_volatile32(reinterpret_cast<const uint32_t*>(&M.n_nodes));
    if (hi - L > novel_cap) hi = L + novel_cap;
#if KM_DEVICE_BUILD
    ChainPrefetch pf;
    pf.node = -1; pf.cnt = 0;
#endif
    while (lo < hi && !(load_shared_volatile32(&M.flags) & 1u)) {
#if KM_DEVICE_BUILD
        if (hi - lo == 1) ws_chain_level(T, W, P, g, M, novel_cap, lo, pf, st, nlook);
        // four lanes per frontier node, one successor letter each: one lookup per lane
        else for (int base = lo; base < hi; base += 8) {
            st |= KM_ST_BRANCHED;                      // (two live nodes in one level)
            const int q = base + (lane >> 2), c = lane & 3;
            bool expand = false;
            uint64_t ck = 0;
            uint32_t cnt = 0, meta = 0;
            uint64_t parent = 0;
            if (q < hi) {
                meta = M.nmeta[q - L];
                if ((int)(meta >> 8) > P.max_stack) st |= KM_ST_TOUCHED_LIMIT;                // MutationFinder.py:140-141
                else {
                    expand = true;
                    parent = M.nk[q - L];
                    ck = succ_kmer(parent, c, T.kmask);
                    nlook += 1;
                }
                M.kid[q - L][c] = 0;
            }
            if (T.lines) cnt = quad_line_query(T, family_of_suffix(T, parent), T.canonical ? canonical(ck, T.k) : ck, expand);
            else if (expand) cnt = table_query(T, ck);
            // Jellyfish.py:61-72 over the group's four counts
            unsigned long long sum = cnt;
            sum += __shfl_xor_sync(0xFFFFFFFFu, sum, 1);
            sum += __shfl_xor_sync(0xFFFFFFFFu, sum, 2);
            double thr = (double)sum * P.ratio;
            if (thr < (double)P.count) thr = (double)P.count;
            bool pass = expand && (double)cnt >= thr;
            const int nkid = __popc((__ballot_sync(0xFFFFFFFFu, pass) >> (lane & ~3)) & 15u);
            int nb = (int)(meta & 255u);
            if (nkid > 1) {                                                                     // MutationFinder.py:153-156
                nb += 1;
                if (nb > P.max_break) { if (pass) st |= KM_ST_TOUCHED_LIMIT; pass = false; }
            }
            ws_child(ctx, W, g, M, k, novel_cap, pass, ck, cnt, pack_meta((int)(meta >> 8) + 1, nb), q, c);
        }
#else
        for (int base = lo; base < hi; base += nl) {
            if (hi - lo > 1) st |= KM_ST_BRANCHED;
            const int q = base + lane;
            const bool active = q < hi;
            bool pass[4] = {false, false, false, false};
            uint64_t ck[4] = {0, 0, 0, 0};
            uint32_t cc[4] = {0, 0, 0, 0};
            uint32_t child_meta = 0;
            if (active) {
                const uint32_t meta = M.nmeta[q - L];
                const int depth = (int)(meta >> 8), breaks = (int)(meta & 255u);
                if (depth > P.max_stack) st |= KM_ST_TOUCHED_LIMIT;                // MutationFinder.py:140-141
                else {
                    ws_query_children(T, M.nk[q - L], ck, cc);
                    nlook += 4;
                    // Jellyfish.py:61-72 -- Python int * float, then max with the int floor, then >=
                    const uint64_t sum = (uint64_t)cc[0] + cc[1] + cc[2] + cc[3];
                    double thr = (double)sum * P.ratio;
                    if (thr < (double)P.count) thr = (double)P.count;
                    int nkid = 0;
                    for (int c = 0; c < 4; ++c) { pass[c] = (double)cc[c] >= thr; nkid += pass[c] ? 1 : 0; }
                    int nb = breaks;
                    if (nkid > 1) {                                                 // MutationFinder.py:153-156
                        st |= KM_ST_BRANCHED;
                        nb = breaks + 1;
                        if (nb > P.max_break) { st |= KM_ST_TOUCHED_LIMIT; pass[0] = pass[1] = pass[2] = pass[3] = false; }
                    }
                    child_meta = pack_meta(depth + 1, nb);
                }
                M.kid[q - L][0] = M.kid[q - L][1] = M.kid[q - L][2] = M.kid[q - L][3] = 0;
            }
            for (int c = 0; c < 4; ++c) ws_child(ctx, W, g, M, k, novel_cap, active && pass[c], ck[c], cc[c], child_meta, q, c);
        }
#endif
        ctx.sync();
        pt.mark_warp(35);
        lo = hi;
        const int n = (int)load_shared_volatile32(reinterpret_cast<const uint32_t*>(&M.n_nodes));
        hi = n - L > novel_cap ? L + novel_cap : n;
    }
    return ws_finish(ctx, W, P, t, g, M, novel_cap, hi, st, nlook, pt);
}

}  // namespace km
