// TEST INFRASTRUCTURE ONLY -- single-lane host build of the device stage functions.
//
// Compiles km_b200/csrc/{table,walk,graph,quant}.h with g++ (KM_HOST_EMU: tid 0 of 1, barriers
// no-ops, atomics plain) so the kernel LOGIC can be compared with the oracle on a machine
// without a GPU (tests/test_emu_pipeline.py).  Never loaded by km_b200/: the product library
// contains no host copy of these functions and fails with KM_E_NOGPU when there is no device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../km_b200/csrc/quant.h"
#include "../../km_b200/csrc/walk_small.h"
#include "../../km_b200/csrc/synth.h"

using namespace km;

struct EmuTable {
    TableView v;
    std::vector<Bucket> store;      // sector buckets, or family lines (4 buckets' worth of bytes each)
};

extern "C" {

void* emu_table_create_layout(int k, int canonical, uint64_t capacity, int lines) {
    EmuTable* t = new EmuTable();
    uint64_t nb = capacity < 64 ? 64 : capacity;
    if (lines) {
        nb = (capacity * 2 + 4) / 5 < 64 ? 64 : (capacity * 2 + 4) / 5;
        t->store.resize(nb * 4 + 4);
        // 128-byte alignment of the first line
        char* raw = (char*)t->store.data();
        char* al = raw + ((128 - ((uintptr_t)raw & 127)) & 127);
        LineSlot* sl = (LineSlot*)al;
        for (uint64_t i = 0; i < nb * KM_LINE_SLOTS; ++i) { sl[i].key = KM_EMPTY_KEY; sl[i].count = 0; sl[i].pad = 0; }
        t->v.buckets = (Bucket*)al; t->v.n_buckets = nb; t->v.k = k; t->v.canonical = canonical; t->v.kmask = kmer_mask(k);
        t->v.n_shards = 1; t->v.my_shard = 0; for (auto& sp : t->v.shard) sp = nullptr; t->v.shard[0] = t->v.buckets;
        t->v.lines = 1;
        return t;
    }
    t->v.lines = 0;
    t->store.resize(nb);
    for (auto& b : t->store) { b.key[0] = b.key[1] = KM_EMPTY_KEY; b.count[0] = b.count[1] = 0; b.pad[0] = b.pad[1] = 0; }
    t->v.buckets = t->store.data(); t->v.n_buckets = nb; t->v.k = k; t->v.canonical = canonical; t->v.kmask = kmer_mask(k);
    t->v.n_shards = 1; t->v.my_shard = 0; for (auto& sp : t->v.shard) sp = nullptr; t->v.shard[0] = t->v.buckets;
    return t;
}
void* emu_table_create(int k, int canonical, uint64_t capacity) {
    const char* e = getenv("KM_TABLE_LINES");
    return emu_table_create_layout(k, canonical, capacity, e && *e && *e != '0');
}
void emu_table_free(void* h) { delete (EmuTable*)h; }
int emu_table_insert(void* h, const uint64_t* keys, const uint32_t* counts, uint64_t n, int mode) {
    EmuTable* t = (EmuTable*)h;
    for (uint64_t i = 0; i < n; ++i) if (table_insert(t->v, keys[i] & t->v.kmask, counts[i], mode) < 0) return -1;
    return 0;
}
int emu_table_synthetic(void* h, uint64_t seed, uint64_t n) {
    EmuTable* t = (EmuTable*)h;
    for (uint64_t i = 0; i < n; ++i) { uint64_t key = synth_key(seed, i, t->v.k); if (table_insert(t->v, key, synth_count(key), KM_INSERT_KEEP) < 0) return -1; }
    return 0;
}
uint32_t emu_query(void* h, uint64_t fwd) { return table_query(((EmuTable*)h)->v, fwd); }
uint64_t emu_revcomp(uint64_t v, int k) { return revcomp(v, k); }
uint64_t emu_synth_key(uint64_t seed, uint64_t i, int k) { return synth_key(seed, i, k); }
uint32_t emu_synth_count(uint64_t key) { return synth_count(key); }
int emu_row_size(void) { return (int)sizeof(Row); }

// One target through walk -> graph -> rows.  Outputs go to caller buffers; returns status bits,
// or -1 when an output buffer is too small.
int emu_find_target(void* h, const uint8_t* codes, int len, double ratio, int64_t count, int steps, int branchs, int nodes,
                    int extra, int32_t* out_n, uint64_t* out_kmer, uint32_t* out_count, int node_cap_out,
                    int32_t* out_n_paths, int32_t* out_path_len, int32_t* out_pool, int path_cap, int pool_cap,
                    int32_t* out_n_rows, Row* out_rows, int row_cap, uint64_t* out_lookups,
                    int small_nodes, int small_cand, int small_paths, int small_cols, int32_t* out_pass) {
    EmuTable* t = (EmuTable*)h;
    const int k = t->v.k;
    const int L = len - k + 1 > 0 ? len - k + 1 : 0;
    const int cap = L + extra;
    uint32_t H = 64;
    while (H < 2u * (uint32_t)cap + 1024u) H <<= 1;
    int64_t seq_off[2] = {0, len}, node_off[2] = {0, cap}, hash_off[2] = {0, (int64_t)H};
    std::vector<uint64_t> node_kmer(cap), hkey(H);
    std::vector<uint32_t> node_count(cap), node_slot(cap), node_kid(4 * (size_t)cap), hval(H), hmeta(H);
    std::vector<uint8_t> hflag(H);
    int32_t n_nodes = 0, n_kept = 0;
    uint32_t status = 0;
    unsigned long long lookups = 0;
    WalkView W{};
    W.n_targets = 1; W.codes = codes; W.seq_off = seq_off; W.node_off = node_off; W.hash_off = hash_off;
    W.node_kmer = node_kmer.data(); W.node_count = node_count.data(); W.node_slot = node_slot.data(); W.node_kid = node_kid.data();
    W.hkey = hkey.data(); W.hval = hval.data(); W.hmeta = hmeta.data(); W.hflag = hflag.data();
    W.n_nodes = &n_nodes; W.n_kept = &n_kept; W.status = &status; W.lookups = &lookups;
    // what km_encode_kernel prepares once per upload
    const int n_words = (len + 15) / 16 + 2;
    std::vector<uint32_t> pack((size_t)n_words, 0u);
    uint8_t pre_bad = 0;
    for (int pos = 0; pos < len; ++pos) {
        uint32_t c = codes[pos];
        if (c > 3) { pre_bad = 1; c = 0; }
        pack[(size_t)(pos >> 4)] |= c << (2 * (15 - (pos & 15)));
    }
    int64_t pack_off[2] = {0, n_words};
    W.pack = pack.data(); W.pack_off = pack_off; W.pre_bad = &pre_bad;
    FindParams P; P.ratio = ratio; P.count = count; P.max_stack = steps; P.max_break = branchs; P.max_node = nodes;
    CtaCtx ctx;
    // the two walk kernels: shared-memory walk first, the general one for what it defers
    bool walked = false;
    {
        const TargetGeom g = target_geom(W, 0, k);
        if (walk_small_fits(g)) {
            W.chunk_target = nullptr; W.chunk_start = nullptr; W.n_chunks = 0;
            for (int i0 = 0; i0 < g.L; i0 += ctx.nt()) ref_probe_chunk<CtaCtx, false>(ctx, t->v, W, P, 0, i0);     // km_ref_probe_kernel
            static WalkSmall M;
            memset(&M, 0x5A, sizeof(M));
            walked = walk_small_target(ctx, t->v, W, P, 0, M);
        }
    }
    if (!walked) { status = 0; lookups = 0; n_kept = 0; walk_target(ctx, t->v, W, P, 0); }
    status &= ~KM_ST_BRANCHED;                 // the scheduler's hint (km_schedule_kernel clears it)
    *out_lookups = lookups;
    *out_n = 0; *out_n_paths = 0; *out_n_rows = 0;
    if (status & (KM_ST_BAD_BASE | KM_ST_DUP_KMER | KM_ST_NODE_OVERFLOW | KM_ST_NODE_LIMIT | KM_ST_TOO_SHORT)) return (int)status;

    std::vector<uint64_t> okmer(cap);
    std::vector<uint32_t> ocount(cap);
    std::vector<int64_t> path_off(path_cap);
    int32_t t_n = 0, t_np = 0, t_pf = 0, t_nr = 0, t_rf = 0;
    unsigned long long used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ResultView R;
    R.flags = getenv("KM_NO_REFINE_JUMP") ? KM_RESULT_NO_REFINE_JUMP : 0;
    R.t_n = &t_n; R.t_n_paths = &t_np; R.t_path_first = &t_pf; R.t_n_rows = &t_nr; R.t_row_first = &t_rf;
    R.out_kmer = okmer.data(); R.out_count = ocount.data();
    R.path_off = path_off.data(); R.path_len = out_path_len; R.pool = out_pool; R.path_cap = path_cap; R.pool_cap = pool_cap;
    R.rows = out_rows; R.row_cap = row_cap; R.used = used;
    R.seq_pool = nullptr; R.path_seq_off = nullptr; R.seq_cap = 0;
    int32_t sched_order[3] = {0, 0, 0}, sched_count[4] = {0, 0, 0, 0};
    R.sched_order = sched_order; R.sched_count = sched_count;

    // the two passes of km_graph_kernel: small capacities with deferral, then the general ones
    auto run_pass = [&](int maxcap, int max_cand, int max_paths, int max_cols, int retry) -> bool {
        // the product's own layout + carving, overlays included (compact = the shared-memory pass)
        const ScratchLayout SL = make_layout(maxcap, max_cand, max_paths, max_cols, retry);
        std::vector<char> arena(SL.stride + 64, (char)0x5A);
        char* basep = arena.data() + ((64 - ((uintptr_t)arena.data() & 63)) & 63);
        const GraphScratch S = carve(SL, basep, retry);
        int sh[32] = {0};
        GraphDims d;
        if (!graph_target(ctx, t->v, W, S, R, 0, &d, sh)) return false;
        emit_rows(ctx, t->v, W, S, R, 0, d, sh[2], sh[3], sh[6], sh);
        return true;
    };
    const int n_all = n_nodes < cap ? n_nodes : cap;
    const bool fits = n_all <= small_nodes - 2 && n_kept + 2 <= small_nodes;
    bool done = false;
    if (fits) done = run_pass(small_nodes - 2, small_cand, small_paths, small_cols, 1) && !(status & KM_ST_RETRY_LARGE);
    if (!done) {
        status &= ~KM_ST_RETRY_LARGE;
        run_pass(cap, KM_MAX_PATHS, KM_MAX_PATHS, KM_MAX_COLS, 0);
    }
    *out_pass = done ? 1 : 2;
    if (getenv("KM_EMU_VERBOSE")) fprintf(stderr, "emu: pass %d simple %llu\n", *out_pass, used[7]);
    if (t_n - 2 > node_cap_out) return -1;
    *out_n = t_n;
    memcpy(out_kmer, okmer.data(), sizeof(uint64_t) * (size_t)(t_n - 2));
    memcpy(out_count, ocount.data(), sizeof(uint32_t) * (size_t)(t_n - 2));
    *out_n_paths = t_np;
    *out_n_rows = t_nr;
    // paths were bump-allocated from offset 0 in order of materialisation, then re-ordered: hand
    // the offsets back through the first ints of a side channel -> compact them here instead
    std::vector<int32_t> compact;
    std::vector<int32_t> lens;
    for (int p = t_pf; p < t_pf + t_np; ++p) { compact.insert(compact.end(), out_pool + path_off[p], out_pool + path_off[p] + out_path_len[p]); lens.push_back(out_path_len[p]); }
    memcpy(out_pool, compact.data(), sizeof(int32_t) * compact.size());
    for (int p = 0; p < t_np; ++p) out_path_len[p] = lens[p];
    // rows of the last pass start at t_rf and refer to path ids t_pf..: rebase both to 0
    for (int r = 0; r < t_nr; ++r) { out_rows[r] = out_rows[t_rf + r]; out_rows[r].path_id -= t_pf; }
    return (int)status;
}

// quant.h solve3_exact on its own: acc = lower triangle of G at [a * 3 + b] (b <= a) + h at [9..11]; 1 = solved
int emu_solve3(const unsigned long long* acc, double* x) { return solve3_exact(acc, x) ? 1 : 0; }

}  // extern "C"
