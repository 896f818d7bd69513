// C ABI of libkm_b200.so, part 3: find_mutation over a batch of targets (km/tools/find_mutation.py:47-58) --
// the layout of a batch in HBM, its upload, the launches (find_launch.h) and the fetch with its capacity-retry
// loop.  Host orchestration only; every piece of arithmetic runs in the kernels.
#include "plan.h"
#include "find_launch.h"

thread_local void (*g_trace_mark)(void*, const char*, int) = nullptr;
thread_local void* g_trace_obj = nullptr;
thread_local int g_trace_sub = -1;

int plan_layout(km_plan* p) {
    km_table* t = p->t;
    const int k = t->k, n = p->n;
    int maxcap = 1;
    if (p->layout_reusable) {
        // the lane's previous plan had the very same target lengths and default capacities: its offset and chunk vectors
        // (borrowed through plan_swap_vecs) ARE this plan's
        maxcap = p->maxcap;
        p->layout_reusable = false;                     // (a capacity retry changes `extra`: it lays out afresh)
    } else {
        p->node_off.assign(n + 1, 0);
        p->hash_off.assign(n + 1, 0);
        for (int i = 0; i < n; ++i) {
            const int64_t len = p->seq_off[i + 1] - p->seq_off[i];
            const int L = (int)std::max<int64_t>(0, len - k + 1);
            const int cap = L + p->extra[i];
            maxcap = std::max(maxcap, cap);
            p->node_off[i + 1] = p->node_off[i] + cap;
            p->hash_off[i + 1] = p->hash_off[i] + pow2_at_least(2 * (uint64_t)cap + 256);
        }
        p->pack_off.assign(n + 1, 0);
        for (int i = 0; i < n; ++i) p->pack_off[i + 1] = p->pack_off[i] + (p->seq_off[i + 1] - p->seq_off[i] + 15) / 16 + 2;
        p->chunk_target.clear(); p->chunk_start.clear();
        for (int i = 0; i < n; ++i) {
            const int L = (int)std::max<int64_t>(0, p->seq_off[i + 1] - p->seq_off[i] - k + 1);
            for (int s0 = 0; s0 < L; s0 += 32) { p->chunk_target.push_back(i); p->chunk_start.push_back(s0); }
        }
        p->maxcap = maxcap;
    }
    const int64_t n_pack = p->pack_off[n];
    const size_t n_chunks = p->chunk_target.size();
    const int64_t n_node = p->n_node = p->node_off[n], n_hash = p->n_hash = p->hash_off[n], n_code = p->n_code = p->seq_off[n];
    p->grid_tiny = std::max(1, std::min(n, t->sm_count * KM_GRAPH_TINY_GRID));
    p->grid_graph = std::max(1, std::min(n, t->sm_count * KM_GRAPH_SMALL_GRID));
    p->grid_large = std::max(1, std::min(n, t->sm_count * 2));
    p->grid_bubble_tiny = std::max(1, std::min((n + KM_BUBBLE_SLOTS - 1) / KM_BUBBLE_SLOTS, t->sm_count * KM_BUBBLE_TINY_MINB));
    p->grid_bubble_small = std::max(1, std::min((n + KM_BUBBLE_SLOTS - 1) / KM_BUBBLE_SLOTS, t->sm_count * KM_BUBBLE_SMALL_MINB));
    const ScratchLayout L0 = make_layout(maxcap, KM_MAX_PATHS, KM_MAX_PATHS, KM_MAX_COLS, 0);
    const int32_t path_cap = p->path_cap, row_cap = p->row_cap;
    const int64_t pool_cap = p->pool_cap, seq_cap = p->seq_cap;

    size_t need = 4096;
    auto acc = [&](size_t bytes) { need = align_up(need, 256) + bytes; };
    acc(n_code); acc(8 * (n + 1)); acc(8 * (n + 1)); acc(8 * (n + 1)); acc(4 * n_chunks); acc(4 * n_chunks);
    acc(4 * (size_t)n_pack); acc(8 * (n + 1)); acc(n);
    acc(8 * n_node); acc(4 * n_node); acc(4 * n_node); acc(16 * n_node);          // node arrays
    acc(8 * n_hash); acc(4 * n_hash); acc(4 * n_hash); acc(n_hash);                // visited sets
    acc(4 * n); acc(4 * n); acc(4 * n); acc(8 * n);                                // n_nodes n_kept status lookups
    for (int i = 0; i < 5; ++i) acc(4 * n);                                        // per-target result ints
    acc(2 * (size_t)n);                                                            // the walk's codes for the scheduler
    acc(8 * n_node); acc(4 * n_node);                                              // canonical nodes
    acc(8 * (size_t)path_cap); acc(4 * (size_t)path_cap); acc(8 * (size_t)path_cap); acc(4 * (size_t)pool_cap);
    acc(sizeof(Row) * (size_t)row_cap); acc((size_t)seq_cap); acc(64); acc(64); acc(20 * (size_t)n + 64);
    acc(L0.stride * (size_t)p->grid_large);
    const size_t n_name = p->fmt ? (size_t)p->fmt_name_off[n] : 0;
    if (p->fmt) {
        // room for the text: rows carry two sequences of about the target's length each
        p->text_cap = 16 * n_code + 512ll * n + (int64_t)row_cap * (int64_t)(p->fmt_db.size() + 64) + (1 << 16);
        acc(n_name); acc(8 * (size_t)(n + 1)); acc(p->fmt_db.size() + 1);
        acc(4 * (size_t)row_cap); acc(4 * (size_t)row_cap); acc((size_t)KM_FMT_ROW_BYTES * (size_t)row_cap);
        acc(8 * (size_t)n); acc(8 * (size_t)(n + 1)); acc((size_t)p->text_cap); acc(64);
    }
    if (int rc = p->dev->reserve(need + 16384)) return rc;
    p->dev->reset();
    Arena& A = *p->dev;
    WalkView& W = p->W;
    W.n_targets = n;
    W.codes = A.take<uint8_t>(n_code);
    // codes .. chunk_start are taken in the order (and with the alignment) plan_upload uses for its pinned staging
    // block, so the whole input goes up with ONE copy
    W.seq_off = A.take<int64_t>(n + 1); W.node_off = A.take<int64_t>(n + 1); W.hash_off = A.take<int64_t>(n + 1);
    W.pack_off = A.take<int64_t>(n + 1);
    W.chunk_target = A.take<int32_t>(n_chunks); W.chunk_start = A.take<int32_t>(n_chunks); W.n_chunks = (int)n_chunks;
    const char* input_end = (const char*)(W.chunk_start + n_chunks);
    if (p->fmt) {
        char* dn = A.take<char>(n_name);
        int64_t* dno = A.take<int64_t>(n + 1);
        char* ddb = A.take<char>(p->fmt_db.size() + 1);
        p->F.names = dn; p->F.name_off = dno; p->F.db_name = ddb; p->F.db_len = (int)p->fmt_db.size();
        input_end = ddb + p->fmt_db.size() + 1;
    }
    p->upload_bytes = (size_t)(input_end - (const char*)W.codes);
    W.pack = A.take<uint32_t>((size_t)n_pack); W.pre_bad = A.take<uint8_t>(n);
    W.node_kmer = A.take<uint64_t>(n_node); W.node_count = A.take<uint32_t>(n_node);
    W.node_slot = A.take<uint32_t>(n_node); W.node_kid = A.take<uint32_t>(4 * n_node);
    W.hkey = A.take<uint64_t>(n_hash); W.hval = A.take<uint32_t>(n_hash); W.hmeta = A.take<uint32_t>(n_hash);
    W.hflag = A.take<uint8_t>(n_hash);
    // per-target state and result ints are contiguous so one memset clears them
    p->state0 = A.take<char>(0);
    W.n_nodes = A.take<int32_t>(n); W.n_kept = A.take<int32_t>(n); W.status = A.take<uint32_t>(n);
    W.lookups = A.take<unsigned long long>(n);
    ResultView& R = p->R;
    R.t_n = A.take<int32_t>(n); R.t_n_paths = A.take<int32_t>(n); R.t_path_first = A.take<int32_t>(n);
    R.t_n_rows = A.take<int32_t>(n); R.t_row_first = A.take<int32_t>(n);
    // the pool cursors and the formatter's flags / total sit in the same block: one memset, one copy back
    R.used = A.take<unsigned long long>(8);     // [0..3] pool cursors, [4..6] work counters of the three graph passes
    p->F.flags = A.take<uint32_t>(16);          // [0] flags, [2..3] total bytes of text (64 bit)
    W.walk_cursor = A.take<uint32_t>(4);
    p->state_bytes = (size_t)(A.take<char>(0) - p->state0);
    // (cleared with the state, but not part of what comes back: the memset covers it, the copy does not)
    W.sched_code = A.take<uint16_t>(n);
    p->clear_bytes = (size_t)(A.take<char>(0) - p->state0);
    R.out_kmer = A.take<uint64_t>(n_node); R.out_count = A.take<uint32_t>(n_node);
    R.path_off = A.take<int64_t>(path_cap); R.path_len = A.take<int32_t>(path_cap);
    p->d_path_seq_off = A.take<int64_t>(path_cap);
    R.pool = A.take<int32_t>(pool_cap); R.path_cap = path_cap; R.pool_cap = pool_cap;
    R.rows = A.take<Row>(row_cap); R.row_cap = row_cap;
    p->d_seq_pool = A.take<char>(seq_cap);
    R.seq_pool = p->d_seq_pool; R.path_seq_off = p->d_path_seq_off; R.seq_cap = seq_cap;
    R.sched_order = A.take<int32_t>(5 * (size_t)n); R.sched_count = A.take<int32_t>(8);
    p->SL = L0;
    p->SL.base = A.take<char>(L0.stride * (size_t)p->grid_large);
    if (p->fmt) {
        p->F.row_len = A.take<int32_t>(row_cap); p->F.row_pos = A.take<int32_t>(row_cap);
        p->F.row_num = A.take<char>((size_t)KM_FMT_ROW_BYTES * (size_t)row_cap);
        p->F.t_len = A.take<int64_t>(n); p->F.t_off = A.take<int64_t>(n + 1);
        p->F.text = A.take<char>((size_t)p->text_cap); p->F.text_cap = p->text_cap;
    }
    { const char* e = getenv("KM_NO_REFINE_JUMP");
      R.flags = ((p->prm.flags & KM_FIND_NO_REFINE_JUMP) || (e && *e && *e != '0')) ? KM_RESULT_NO_REFINE_JUMP : 0; }
    p->P.ratio = p->prm.ratio; p->P.count = p->prm.count; p->P.max_stack = p->prm.steps;
    p->P.max_break = p->prm.branchs; p->P.max_node = p->prm.nodes;
    return 0;
}

// the host half of the upload: every input of the batch into ONE pinned staging block
int plan_stage(km_plan* p, cudaStream_t s) {
    const int n = p->n;
    const size_t n_chunks = p->chunk_target.size();
    const size_t n_name = p->fmt ? (size_t)p->fmt_name_off[n] : 0;
    if (int rc = p->pin->reserve((size_t)p->n_code + 40 * (size_t)(n + 1) + 8 * n_chunks + n_name + p->fmt_db.size() + 16384)) return rc;
    p->pin->reset();
    uint8_t* h_codes = p->pin->take<uint8_t>(p->n_code);
    int64_t* h_seq_off = p->pin->take<int64_t>(n + 1);
    int64_t* h_node_off = p->pin->take<int64_t>(n + 1);
    int64_t* h_hash_off = p->pin->take<int64_t>(n + 1);
    int64_t* h_pack_off = p->pin->take<int64_t>(n + 1);
    memcpy(h_pack_off, p->pack_off.data(), 8 * (n + 1));
    int32_t* h_ct = p->pin->take<int32_t>(n_chunks);
    int32_t* h_cs = p->pin->take<int32_t>(n_chunks);
    if (n_chunks) { memcpy(h_ct, p->chunk_target.data(), 4 * n_chunks); memcpy(h_cs, p->chunk_start.data(), 4 * n_chunks); }
    memcpy(h_codes, p->targets_ext ? p->targets_ext : p->targets.data(), p->n_code);   // letters; km_encode_kernel turns them into codes on the device
    memcpy(h_seq_off, p->seq_off.data(), 8 * (n + 1));
    memcpy(h_node_off, p->node_off.data(), 8 * (n + 1));
    memcpy(h_hash_off, p->hash_off.data(), 8 * (n + 1));
    CU(cudaEventRecord(p->ev[0], s));
    const char* h_end = (const char*)(h_cs + n_chunks);
    if (p->fmt) {
        char* hn = p->pin->take<char>(n_name);
        int64_t* hno = p->pin->take<int64_t>(n + 1);
        char* hdb = p->pin->take<char>(p->fmt_db.size() + 1);
        memcpy(hn, p->fmt_names, n_name);
        memcpy(hno, p->fmt_name_off, 8 * (size_t)(n + 1));
        memcpy(hdb, p->fmt_db.c_str(), p->fmt_db.size() + 1);
        h_end = hdb + p->fmt_db.size() + 1;
    }
    if ((size_t)(h_end - (const char*)h_codes) != p->upload_bytes)
        return fail(KM_E_ARG, "internal: staging block and device input block differ in layout");
    p->h_stage = h_codes;
    p->bytes_h2d = (unsigned long long)p->n_code + 32ull * (n + 1) + 8ull * n_chunks + (p->fmt ? n_name + 8ull * (n + 1) + p->fmt_db.size() : 0ull);
    return 0;
}

// the CUDA half of the upload: one copy of the staged block, then the packing kernel
int plan_upload_enqueue(km_plan* p, cudaStream_t s) {
    CU(cudaMemcpyAsync((void*)p->W.codes, p->h_stage, p->upload_bytes, cudaMemcpyHostToDevice, s));
    CU(km_launch_encode(p->W, s));
    return 0;
}

int plan_upload(km_plan* p, cudaStream_t s) {
    if (int rc = plan_stage(p, s)) return rc;
    return plan_upload_enqueue(p, s);
}

// memsets + the two kernels, asynchronously on `s`
int plan_launch(km_plan* p, cudaStream_t s) {
    km_table* t = p->t;
    if (p->n == 0) return 0;
    const bool timed = !p->fmt || p->trace_events;          // km_find_text enqueues as little as it can: no per-phase events
    // (inside a capture a phase event becomes an EXTERNAL record node: it is recorded when the graph runs and can be timed)
    auto mark = [&](cudaEvent_t ev) { return cudaEventRecordWithFlags(ev, s, p->capturing ? cudaEventRecordExternal : cudaEventRecordDefault); };
    if (timed) CU(mark(p->ev[1]));
    CU(cudaMemsetAsync(p->state0, 0, p->clear_bytes, s));
    CU(km_launch_ref_probe(t->view(), p->W, p->P, s));
    if (timed) CU(mark(p->ev[6]));
    CU(km_launch_walks(t->view(), p->W, p->P, s));
    if (timed) CU(mark(p->ev[2]));
    // shared-memory passes first, then the general pass for large or deferred targets
    CU(km_launch_schedule(p->W, p->R, s));
    // The passes run side by side (each pass's tail fills with the others' CTAs): the simple bubbles of the two size classes,
    // a small group of threads per target (graph_bubble.h), and -- for the targets whose walk branched -- the two
    // shared-memory classes of the CTA-per-target pass; then the general pass for what any of them handed on.
    const bool bubbles = km_bubble_pass_enabled();
    CU(cudaEventRecord(p->fork, s));
    // The CTA-per-target classes go FIRST: with the bubble pass in front of them they hold the few targets whose walk
    // branched, among them the one or two whose wide cluster keeps a lane busy for 0.1-0.2 ms -- launched behind the
    // bubble kernels (persistent CTAs that fill every SM until their lists are empty) such a target only STARTED when the
    // bubbles were done, and the pass took both times one after the other.
    CU(km_launch_graph(0, 0, p->grid_tiny, t->view(), p->W, p->SL, p->R, s));
    CU(cudaStreamWaitEvent(p->side, p->fork, 0));
    CU(km_launch_graph(1, 1, p->grid_graph, t->view(), p->W, p->SL, p->R, p->side));
    CU(cudaEventRecord(p->join, p->side));
    if (bubbles) {
        CU(cudaStreamWaitEvent(p->side2, p->fork, 0));
        CU(km_launch_bubble(1, p->grid_bubble_small, t->view(), p->W, p->R, p->side2));
        CU(cudaEventRecord(p->join2, p->side2));
        CU(cudaStreamWaitEvent(p->side3, p->fork, 0));
        CU(km_launch_bubble(0, p->grid_bubble_tiny, t->view(), p->W, p->R, p->side3));
        CU(cudaEventRecord(p->join3, p->side3));
    }
    CU(cudaStreamWaitEvent(s, p->join, 0));
    if (bubbles) { CU(cudaStreamWaitEvent(s, p->join2, 0)); CU(cudaStreamWaitEvent(s, p->join3, 0)); }
    CU(km_launch_graph(2, 2, p->grid_large, t->view(), p->W, p->SL, p->R, s));
    if (timed) CU(mark(p->ev[3]));
    p->n_launches += bubbles ? 8 : 6;
    if (p->fmt) {
        CU(km_launch_format(p->W, p->R, p->F, t->k, s));
        if (p->trace_events) CU(mark(p->ev[4]));
        p->n_launches += 3;
    }
    p->launched = true;
    return 0;
}

// D2H of per-target ints, then exactly the used extents.  `want_graph` also brings back the
// node arrays and index paths (needed by the MutationFinder attribute views and the parity tests;
// the TSV formatter only needs rows + spelled paths).
int plan_download(km_plan* p, cudaStream_t s, km_result* res, bool want_graph, bool head_only) {
    km_table* t = p->t;
    const int n = p->n;
    const WalkView& W = p->W;
    const ResultView& R = p->R;
    res->n_targets = n; res->k = t->k;
    // the per-target state and result ints sit back to back on the device (plan_layout): ONE copy brings the
    // block into pinned memory and the result's arrays are views into it
    if (int rc = res->head.reserve(t->pool, p->state_bytes + 1024)) return rc;
    Span<char> blk = res->head.take<char>(p->state_bytes);
    auto view = [&](const void* dev_ptr) { return blk.data() + ((const char*)dev_ptr - p->state0); };
    res->status.p = (uint32_t*)view(W.status); res->status.n = (size_t)n;
    res->n_nodes.p = (int32_t*)view(R.t_n); res->n_nodes.n = (size_t)n;
    res->path_count.p = (int32_t*)view(R.t_n_paths); res->path_count.n = (size_t)n;
    res->path_first.p = (int32_t*)view(R.t_path_first); res->path_first.n = (size_t)n;
    res->row_count.p = (int32_t*)view(R.t_n_rows); res->row_count.n = (size_t)n;
    res->row_first.p = (int32_t*)view(R.t_row_first); res->row_first.n = (size_t)n;
    res->lookups.p = (unsigned long long*)view(W.lookups); res->lookups.n = (size_t)n;
    res->used.p = (unsigned long long*)view(R.used); res->used.n = 8;
    unsigned long long* used = res->used.data();
    const uint32_t* fmt_info = (const uint32_t*)view(p->F.flags);     // device text: [0] flags, [2..3] total bytes
    if (n) CU(cudaMemcpyAsync(blk.data(), p->state0, p->state_bytes, cudaMemcpyDeviceToHost, s));
    else memset(blk.data(), 0, p->state_bytes);
    if (head_only) {
        CU(km_wait_stream(s, p->wait_ev));
        long long total = 0;
        memcpy(&total, fmt_info + 2, 8);
        res->dev_text_len = total; res->dev_text_flags = fmt_info[0];
        res->has_graph = false;
        res->bytes_h2d = p->bytes_h2d;
        res->bytes_d2h = (unsigned long long)p->state_bytes;
        res->text_len = -1; res->text.reset(); res->fmt_key.clear();
        return 0;
    }
    CU(cudaStreamSynchronize(s));
    const size_t n_paths = std::min<unsigned long long>(used[0], p->path_cap), n_pool = std::min<unsigned long long>(used[1], p->pool_cap);
    const size_t n_rows = std::min<unsigned long long>(used[2], p->row_cap), n_seq = std::min<unsigned long long>(used[3], p->seq_cap);
    const size_t n_node = want_graph ? (size_t)p->n_node : 0, n_pool_c = want_graph ? n_pool : 0;
    if (int rc = res->body.reserve(t->pool, 20 * n_paths + sizeof(Row) * n_rows + n_seq + 4 * n_pool_c + 12 * n_node + 64 * 10)) return rc;
    res->path_off = res->body.take<int64_t>(n_paths); res->path_len = res->body.take<int32_t>(n_paths);
    res->path_seq_off = res->body.take<int64_t>(n_paths);
    res->rows = res->body.take<km_row>(n_rows); res->seq_pool = res->body.take<char>(n_seq);
    res->path_pool = res->body.take<int32_t>(n_pool_c);
    res->node_kmer = res->body.take<uint64_t>(n_node); res->node_count = res->body.take<uint32_t>(n_node);
    res->node_off = p->node_off;
    if (n_paths) {
        CU(cudaMemcpyAsync(res->path_off.data(), R.path_off, 8 * n_paths, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(res->path_len.data(), R.path_len, 4 * n_paths, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(res->path_seq_off.data(), p->d_path_seq_off, 8 * n_paths, cudaMemcpyDeviceToHost, s));
    }
    if (n_rows) CU(cudaMemcpyAsync(res->rows.data(), R.rows, sizeof(Row) * n_rows, cudaMemcpyDeviceToHost, s));
    if (n_seq) CU(cudaMemcpyAsync(res->seq_pool.data(), p->d_seq_pool, n_seq, cudaMemcpyDeviceToHost, s));
    res->has_graph = want_graph;
    if (want_graph) {
        if (n_pool) CU(cudaMemcpyAsync(res->path_pool.data(), R.pool, 4 * n_pool, cudaMemcpyDeviceToHost, s));
        if (n_node) {
            CU(cudaMemcpyAsync(res->node_kmer.data(), R.out_kmer, 8 * n_node, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(res->node_count.data(), R.out_count, 4 * n_node, cudaMemcpyDeviceToHost, s));
        }
    }
    res->bytes_h2d = p->bytes_h2d;
    res->bytes_d2h = (unsigned long long)p->state_bytes + 20ull * n_paths + sizeof(Row) * n_rows + n_seq +
                     (want_graph ? 4ull * n_pool + 12ull * p->n_node : 0ull);
    res->text_len = -1; res->text.reset(); res->fmt_key.clear();
    if (!p->fmt) CU(cudaEventRecord(p->ev[4], s));
    CU(cudaStreamSynchronize(s));
    float ms;
    if (n && !p->fmt) {
        CU(cudaEventElapsedTime(&ms, p->ev[0], p->ev[1])); res->ms_h2d = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[1], p->ev[2])); res->ms_walk = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[2], p->ev[3])); res->ms_graph = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[3], p->ev[4])); res->ms_d2h = ms;
        CU(cudaEventElapsedTime(&ms, p->ev[0], p->ev[4])); res->ms_total = ms;
    }
    return 0;
}

// a plan on a lane borrows the lane's host vectors (and hands them back, km_find_text) for their capacity
// hand the plan's vectors back to its lane, remembering whether they still describe the default layout
void plan_return_vecs(km_plan* p, km_table::Lane* lane) {
    const int32_t extra0 = p->prm.extra_nodes > 0 ? p->prm.extra_nodes : 256;
    lane->vecs_pristine = p->n_retries == 0;
    lane->vecs_k = p->t->k; lane->vecs_extra0 = extra0; lane->vecs_maxcap = p->maxcap;
    plan_swap_vecs(p, lane->vecs);
}

void plan_swap_vecs(km_plan* p, km_table::PlanVecs& v) {
    p->seq_off.swap(v.seq_off); p->node_off.swap(v.node_off); p->hash_off.swap(v.hash_off); p->pack_off.swap(v.pack_off);
    p->chunk_target.swap(v.chunk_target); p->chunk_start.swap(v.chunk_start); p->extra.swap(v.extra);
}

int plan_init(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params, km_plan* p,
                     bool borrow_arena, km_table::Lane* lane) {
    if (lane) plan_swap_vecs(p, lane->vecs);
    p->t = t; p->n = n; p->prm = *params;
    p->stream = lane ? lane->stream : t->stream;
    p->side = lane ? lane->side : t->side;
    p->side2 = lane ? lane->side2 : t->side2; p->side3 = lane ? lane->side3 : t->side3;
    p->join2 = lane ? lane->join2 : t->join2; p->join3 = lane ? lane->join3 : t->join3;
    p->ev = lane ? lane->ev : t->ev;
    p->fork = lane ? lane->fork : t->fork;
    p->join = lane ? lane->join : t->join;
    p->wait_ev = lane ? &lane->wait_ev : nullptr;
    if (p->prm.steps > 60000 || p->prm.branchs > 250) return fail(KM_E_ARG, "steps must be <= 60000 and branchs <= 250");
    const int64_t total = n ? offsets[n] : 0;
    if (!p->targets_ext) p->targets.assign(seqs ? seqs : "", (size_t)total);
    const int32_t extra0 = p->prm.extra_nodes > 0 ? p->prm.extra_nodes : 256;
    // a lane's previous plan leaves its vectors behind: when the target lengths are the same again (a fixed panel, sample
    // after sample) and no capacity was grown, the whole layout is reused
    p->layout_reusable = lane && n > 0 && lane->vecs_k == t->k && lane->vecs_extra0 == extra0 && lane->vecs_pristine &&
                         p->seq_off.size() == (size_t)n + 1 && p->node_off.size() == (size_t)n + 1 && p->extra.size() == (size_t)n &&
                         memcmp(p->seq_off.data(), offsets, sizeof(int64_t) * ((size_t)n + 1)) == 0;
    if (lane) p->maxcap = lane->vecs_maxcap;
    if (!p->layout_reusable) {
        p->seq_off.assign(1, 0);
        if (n) p->seq_off.assign(offsets, offsets + n + 1);
        p->extra.assign((size_t)n, extra0);
    }
    int64_t n_ref = 0;
    for (int i = 0; i < n; ++i) n_ref += std::max<int64_t>(0, offsets[i + 1] - offsets[i] - t->k + 1);
    p->gave_up.assign((size_t)n, 0);
    p->path_cap = std::max(64, 8 * n);
    p->row_cap = std::max(64, 16 * n);
    p->pool_cap = std::max<int64_t>(1 << 16, 6 * (n_ref + 64ll * n));
    p->seq_cap = p->pool_cap + (int64_t)p->path_cap * t->k;
    p->extra_max = std::max(1024, p->prm.nodes + 4 * p->prm.steps + 4096);
    p->own_pin.host = true;
    p->dev = lane ? &lane->dev : borrow_arena ? &t->dev_find : &p->own_dev;
    p->pin = lane ? &lane->pin : borrow_arena ? &t->pin_find : &p->own_pin;
    trace_here("  init: copies");
    if (int rc = plan_layout(p)) return rc;
    trace_here("  init: layout");
    const int rc_up = p->defer_upload ? plan_stage(p, p->stream) : plan_upload(p, p->stream);
    trace_here("  init: upload");
    return rc_up;
}

std::string plan_graph_key(const km_plan* p) {
    std::string k;
    auto add = [&](const void* ptr, size_t n) { k.append((const char*)ptr, n); };
    const TableView T = p->t->view();
    add(&T, sizeof(T)); add(&p->W, sizeof(p->W)); add(&p->R, sizeof(p->R)); add(&p->F, sizeof(p->F)); add(&p->P, sizeof(p->P));
    add(&p->SL, sizeof(p->SL));
    const int64_t misc[] = {p->n, p->grid_tiny, p->grid_graph, p->grid_large, p->grid_bubble_tiny, p->grid_bubble_small, km_bubble_pass_enabled() ? 1 : 0, (int64_t)p->state_bytes, (int64_t)p->upload_bytes,
                            (int64_t)(intptr_t)p->state0, (int64_t)(intptr_t)p->h_stage, (int64_t)(intptr_t)p->stream,
                            (int64_t)(intptr_t)p->side, (int64_t)(intptr_t)p->side2, (int64_t)(intptr_t)p->side3, p->fmt ? 1 : 0};
    add(misc, sizeof(misc));
    return k;
}

// fetch with the capacity-retry loop: targets whose exploration overflowed get 8x the node
// capacity, exhausted pools grow 4x, and the batch is re-run
int plan_fetch(km_plan* p, km_result* res, bool want_graph, bool head_only) {
    km_table* t = p->t;
    for (int attempt = 0; attempt < 12; ++attempt) {
        if (!p->launched) if (int rc = plan_launch(p, p->stream)) return rc;
        if (int rc = plan_download(p, p->stream, res, want_graph, head_only)) return rc;
        bool again = false, pool_over = false;
        for (int i = 0; i < p->n; ++i) {
            if (res->status[i] & KM_ST_NODE_OVERFLOW) {
                // A target that visits more than extra_max nodes off the reference is ITS OWN failure: it keeps
                // KM_ST_NODE_OVERFLOW as its final status (the host raises for it after the rows of the targets
                // before it are printed, like the reference's node limit, MutationFinder.py:143-148) and the rest
                // of the batch completes.  On a relaunch it runs with no capacity, so it overflows at once.
                if (p->gave_up[(size_t)i]) continue;
                if (p->extra[i] >= p->extra_max) { p->gave_up[(size_t)i] = 1; p->extra[i] = 0; continue; }
                p->extra[i] = (int32_t)std::min<int64_t>(p->extra_max, (int64_t)p->extra[i] * 8);
                again = true;
            }
            if (res->status[i] & KM_ST_PATH_OVERFLOW) pool_over = true;
        }
        if (pool_over) {
            p->pool_cap *= 4; p->path_cap *= 4; p->row_cap *= 4;
            p->seq_cap = p->pool_cap + (int64_t)p->path_cap * t->k;
            again = true;
        }
        res->n_launches = p->n_launches; res->n_retries = p->n_retries;
        if (!again) return 0;
        p->n_retries++;
        p->launched = false;
        if (int rc = plan_layout(p)) return rc;
        if (int rc = plan_upload(p, p->stream)) return rc;
    }
    return fail(KM_E_LIMIT, "km_find: capacities still exceeded after 12 attempts");
}

extern "C" int km_find_plan_create(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params,
                                   km_plan** out) {
    if (!t || !out || n < 0 || (n && (!seqs || !offsets)) || !params) return fail(KM_E_ARG, "km_find_plan_create: bad argument");
    CU(cudaSetDevice(t->device));
    if (int rc = km_ensure_linked(t)) return rc;
    km_plan* p = new km_plan();
    if (int rc = plan_init(t, seqs, offsets, n, params, p, false)) { delete p; return rc; }
    CU(cudaStreamSynchronize(t->stream));
    // its own side stream and events (the table's are shared by km_find_batch and by other plans)
    CU(cudaStreamCreateWithFlags(&p->own_side, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->own_side2, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->own_side3, cudaStreamNonBlocking));
    for (auto& e : p->own_ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&p->own_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p->own_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p->own_join2, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p->own_join3, cudaEventDisableTiming));
    p->side = p->own_side; p->ev = p->own_ev; p->fork = p->own_fork; p->join = p->own_join;
    p->side2 = p->own_side2; p->side3 = p->own_side3; p->join2 = p->own_join2; p->join3 = p->own_join3;
    CU(cudaEventRecord(p->ev[0], t->stream));          // (the upload's start mark was taken on the table's event)
    *out = p;
    return 0;
}

extern "C" int km_find_plan_launch(km_plan* p, void* stream) {
    if (!p) return fail(KM_E_ARG, "null plan");
    CU(cudaSetDevice(p->t->device));
    if (int rc = km_ensure_linked(p->t)) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : p->stream;
    // A resident plan is launched again and again (a fixed panel against sample after sample): from its third launch on the
    // sequence -- memset, nine kernels on four streams, their events -- is ONE graph launch; the gaps between dependent
    // kernels (4 us on a stream, 10-14 us across streams through an event) shrink to what the hardware needs.
    // KM_NO_GRAPH=1 = A/B switch.
    static const bool graphs_on = !getenv("KM_NO_GRAPH");
    if (!graphs_on) return plan_launch(p, s);
    const std::string key = plan_graph_key(p);
    if (p->gexec && p->gkey == key) {
        if (cudaGraphLaunch(p->gexec, s) == cudaSuccess) { p->launched = true; p->n_launches += km_bubble_pass_enabled() ? 8 : 6; return 0; }
        cudaGetLastError();
        cudaGraphExecDestroy(p->gexec); p->gexec = nullptr;
    }
    if (p->gkey != key) { p->gkey = key; p->direct_launches = 0; if (p->gexec) { cudaGraphExecDestroy(p->gexec); p->gexec = nullptr; } }
    if (p->direct_launches++ >= 2 && !p->gexec) {
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            p->capturing = true;
            const int rc = plan_launch(p, s);
            p->capturing = false;
            const cudaError_t ce = cudaStreamEndCapture(s, &graph);
            bool ok = !rc && ce == cudaSuccess && graph && cudaGraphInstantiate(&p->gexec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok && cudaGraphLaunch(p->gexec, s) == cudaSuccess) return 0;
            cudaGetLastError();
            if (p->gexec) { cudaGraphExecDestroy(p->gexec); p->gexec = nullptr; }
            p->direct_launches = -1000000;          // do not try again for this layout
        } else cudaGetLastError();
    }
    return plan_launch(p, s);
}

extern "C" int km_find_plan_last_ms(km_plan* p, float* walk_ms, float* graph_ms) {
    if (!p || !p->launched) return fail(KM_E_ARG, "km_find_plan_last_ms: nothing launched");
    CU(cudaSetDevice(p->t->device));
    CU(cudaEventSynchronize(p->ev[3]));
    if (walk_ms) CU(cudaEventElapsedTime(walk_ms, p->ev[1], p->ev[2]));
    if (graph_ms) CU(cudaEventElapsedTime(graph_ms, p->ev[2], p->ev[3]));
    return 0;
}

extern "C" int km_find_plan_kernel_ms(km_plan* p, float* out3) {
    if (!p || !p->launched || !out3) return fail(KM_E_ARG, "km_find_plan_kernel_ms: nothing launched");
    CU(cudaSetDevice(p->t->device));
    CU(cudaEventSynchronize(p->ev[3]));
    CU(cudaEventElapsedTime(&out3[0], p->ev[1], p->ev[6]));    // memsets + reference probe
    CU(cudaEventElapsedTime(&out3[1], p->ev[6], p->ev[2]));    // the two walk kernels
    CU(cudaEventElapsedTime(&out3[2], p->ev[2], p->ev[3]));    // the two graph kernels
    return 0;
}

extern "C" int km_find_plan_fetch(km_plan* p, int want_graph, km_result** out) {
    if (!p || !out) return fail(KM_E_ARG, "null argument");
    CU(cudaSetDevice(p->t->device));
    CU(cudaDeviceSynchronize());         // launches may have gone to a caller's stream
    km_result* res = new km_result();
    res->targets = p->targets;
    res->seq_off = p->seq_off;
    if (int rc = plan_fetch(p, res, want_graph != 0)) { delete res; return rc; }
    *out = res;
    return 0;
}

extern "C" void km_find_plan_free(km_plan* p) {
    if (!p) return;
    cudaSetDevice(p->t->device);
    p->own_dev.release();
    p->own_pin.release();
    for (auto& e : p->own_ev) if (e) cudaEventDestroy(e);
    if (p->own_fork) cudaEventDestroy(p->own_fork);
    if (p->own_join) cudaEventDestroy(p->own_join);
    if (p->own_join2) cudaEventDestroy(p->own_join2);
    if (p->own_join3) cudaEventDestroy(p->own_join3);
    if (p->gexec) cudaGraphExecDestroy(p->gexec);
    if (p->own_side) cudaStreamDestroy(p->own_side);
    if (p->own_side2) cudaStreamDestroy(p->own_side2);
    if (p->own_side3) cudaStreamDestroy(p->own_side3);
    delete p;
}

extern "C" int km_find_batch(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params,
                             km_result** out) {
    if (!t || !out || n < 0 || (n && (!seqs || !offsets)) || !params) return fail(KM_E_ARG, "km_find_batch: bad argument");
    CU(cudaSetDevice(t->device));
    if (int rc = km_ensure_linked(t)) return rc;
    km_plan plan;
    if (int rc = plan_init(t, seqs, offsets, n, params, &plan, true)) return rc;
    km_result* res = new km_result();
    res->seq_off = plan.seq_off;
    if (int rc = plan_fetch(&plan, res, (params->flags & KM_FIND_NO_GRAPH) == 0)) { delete res; return rc; }
    res->targets.swap(plan.targets);      // after the fetch: a capacity retry uploads the letters again
    *out = res;
    return 0;
}

extern "C" int km_result_get(const km_result* r, km_result_view* v) {
    if (!r || !v) return fail(KM_E_ARG, "null argument");
    memset(v, 0, sizeof(*v));
    v->n_targets = r->n_targets; v->n_paths = (int32_t)r->path_off.size(); v->n_rows = (int32_t)r->rows.size(); v->k = r->k;
    if (!r->parts.empty()) {      // a km_find_text result: the text and the per-target status are what it holds
        v->status = r->all_status.data();
        v->ms_h2d = r->ms_h2d; v->ms_walk = r->ms_walk; v->ms_graph = r->ms_graph; v->ms_d2h = r->ms_d2h; v->ms_total = r->ms_total;
        v->n_launches = r->n_launches; v->n_retries = r->n_retries; v->has_graph = 0;
        v->bytes_h2d = r->bytes_h2d; v->bytes_d2h = r->bytes_d2h;
        return 0;
    }
    v->status = r->status.data(); v->n_nodes = r->n_nodes.data(); v->node_off = r->node_off.data();
    v->node_kmer = r->node_kmer.data(); v->node_count = r->node_count.data();
    v->path_first = r->path_first.data(); v->path_count = r->path_count.data();
    v->path_off = r->path_off.data(); v->path_len = r->path_len.data(); v->path_pool = r->path_pool.data();
    v->row_first = r->row_first.data(); v->row_count = r->row_count.data(); v->rows = r->rows.data();
    v->lookups = reinterpret_cast<const uint64_t*>(r->lookups.data());
    v->ms_h2d = r->ms_h2d; v->ms_walk = r->ms_walk; v->ms_graph = r->ms_graph; v->ms_d2h = r->ms_d2h; v->ms_total = r->ms_total;
    v->n_launches = r->n_launches; v->n_retries = r->n_retries; v->has_graph = r->has_graph ? 1 : 0;
    v->reserved = r->used.size() >= 8 ? (int32_t)r->used[7] : 0;       // targets whose graph was a simple bubble (graph.h)
    v->bytes_h2d = r->bytes_h2d; v->bytes_d2h = r->bytes_d2h;
    return 0;
}

extern "C" void km_result_free(km_result* r) { delete r; }

