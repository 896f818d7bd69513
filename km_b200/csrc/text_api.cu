// C ABI of libkm_b200.so, part 4: the text `km find_mutation` prints.  The host formatter (format_rows_of,
// put_fixed, nat_cmp: PathQuant.Path.__str__ PathQuant.py:37-49, MutationFinder.get_paths :813-833,
// common.natsortkey common.py:95-116) serves km_result_format_* and is the twin of the device formatter in
// format.h; km_find_text pipelines a batch as sub-batches whose text is formatted on the device.
#include "plan.h"
#include "find_launch.h"

// ---- row formatting (PathQuant.Path.__str__, MutationFinder.get_paths) ------------------------
static const char* TYPE_NAME[6] = {"Reference", "Substitution", "ITD", "Indel", "Insertion", "Deletion"};

// "%.{prec}f" of a double, digit for digit what Python / glibc print (the exact binary value rounded
// half-to-even at the last printed digit), without snprintf: |v| = m * 2^e with m < 2^53, so
// m * 10^prec fits 64 bits for prec <= 3 and the rounding is decided on integers.
static char* put_uint(char* o, unsigned long long v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}
static char* put_int(char* o, long long v) {
    if (v < 0) { *o++ = '-'; return put_uint(o, 0ull - (unsigned long long)v); }
    return put_uint(o, (unsigned long long)v);
}
static char* put_fixed(char* o, double v, int prec) {
    if (std::isnan(v)) { memcpy(o, "nan", 3); return o + 3; }          // Python prints nan without a sign
    if (std::isinf(v)) { if (v < 0) *o++ = '-'; memcpy(o, "inf", 3); return o + 3; }
    const double a = fabs(v);
    if (prec > 3 || a >= 4503599627370496.0) return o + snprintf(o, 400, "%.*f", prec, v);
    if (std::signbit(v)) *o++ = '-';
    static const unsigned long long P10[4] = {1ull, 10ull, 100ull, 1000ull};
    int e;
    const double fr = frexp(a, &e);                                    // a = fr * 2^e, fr in [0.5, 1)
    unsigned long long q = 0;
    if (a != 0.0) {
        const unsigned long long m = (unsigned long long)ldexp(fr, 53);   // exact 53-bit integer
        const int e2 = e - 53;                                            // a = m * 2^e2, e2 < 0 here
        const unsigned long long scaled = m * P10[prec];
        const int sh = -e2;
        if (sh <= 0) q = scaled << (-sh);
        else if (sh >= 64) q = 0;
        else {
            q = scaled >> sh;
            const unsigned long long rem = scaled & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
            if (rem > half || (rem == half && (q & 1ull))) ++q;
        }
    }
    const unsigned long long ip = q / P10[prec], fp = q % P10[prec];
    o = put_uint(o, ip);
    if (prec > 0) {
        *o++ = '.';
        for (int d = prec - 1; d >= 0; --d) *o++ = (char)('0' + (fp / P10[d]) % 10);
    }
    return o;
}

// common.natsortkey (common.py:95-116) on two strings without building the token lists:
// re.split('([0-9]+)', key) alternates text / digit runs starting and ending with a (possibly empty)
// text chunk; text compares lower-cased, digit runs as integers, and a list that is a prefix of the
// other sorts first.
static int nat_cmp(const char* a, size_t na, const char* b, size_t nb) {
    size_t i = 0, j = 0;
    for (;;) {
        // text chunks
        for (;;) {
            const bool ea = i >= na || isdigit((unsigned char)a[i]), eb = j >= nb || isdigit((unsigned char)b[j]);
            if (ea || eb) { if (ea != eb) return ea ? -1 : 1; break; }
            const int ca = tolower((unsigned char)a[i]), cb = tolower((unsigned char)b[j]);
            if (ca != cb) return ca < cb ? -1 : 1;
            ++i; ++j;
        }
        const bool enda = i >= na, endb = j >= nb;
        if (enda || endb) return enda == endb ? 0 : (enda ? -1 : 1);
        // digit runs as integers of any length
        size_t i2 = i, j2 = j;
        while (i2 < na && isdigit((unsigned char)a[i2])) ++i2;
        while (j2 < nb && isdigit((unsigned char)b[j2])) ++j2;
        size_t ia = i, jb = j;
        while (ia + 1 < i2 && a[ia] == '0') ++ia;
        while (jb + 1 < j2 && b[jb] == '0') ++jb;
        if (i2 - ia != j2 - jb) return i2 - ia < j2 - jb ? -1 : 1;
        const int c = memcmp(a + ia, b + jb, i2 - ia);
        if (c) return c < 0 ? -1 : 1;
        i = i2; j = j2;
    }
}

struct FmtRow {
    const km_row* w;
    const char* name; uint32_t name_len;       // variant name ("" for Reference)
    const char* line; uint32_t line_len;
};

// the rows of target tg: text into `arena` (unsorted), then sorted as MutationFinder.get_paths does
// (:825-829) and appended to `out`
static void format_rows_of(const km_result* r, int32_t tg, const char* db_name, size_t db_len, const char* qn, size_t qn_len,
                           std::vector<char>& arena, std::vector<FmtRow>& rows, std::vector<char>& out) {
    const int k = r->k;
    const char* tseq = r->targets.data() + r->seq_off[tg];
    const int nrow = r->row_count[tg];
    rows.clear();
    if (nrow <= 0) return;
    // upper bound of this target's text
    size_t need = 0;
    for (int i = 0; i < nrow; ++i) {
        const km_row& w = r->rows[r->row_first[tg] + i];
        need += db_len + qn_len + 256 + (size_t)(w.del_len + w.ins_len) + (size_t)(w.var_end - w.var_begin + k) +
                (size_t)(w.ref_end - w.ref_begin + k);
    }
    if (arena.size() < need) arena.resize(need + need / 2);
    char* o = arena.data();
    for (int i = 0; i < nrow; ++i) {
        const km_row& w = r->rows[r->row_first[tg] + i];
        const char* pseq = r->seq_pool.data() + r->path_seq_off[w.path_id];
        FmtRow fr;
        fr.w = &w;
        fr.line = o;
        memcpy(o, db_name, db_len); o += db_len; *o++ = '\t';
        memcpy(o, qn, qn_len); o += qn_len; *o++ = '\t';
        const size_t tl = strlen(TYPE_NAME[w.type]);
        memcpy(o, TYPE_NAME[w.type], tl); o += tl; *o++ = '\t';
        fr.name = o;
        if (w.type != 0) {      // "{}\t{}:{}:{}" (MutationFinder.py:483-488); Reference -> "Reference\t"
            o = put_int(o, w.name_start); *o++ = ':';
            for (int j = 0; j < w.del_len; ++j) *o++ = (char)tolower((unsigned char)tseq[w.del_begin + j + k - 1]);
            *o++ = '/';
            memcpy(o, pseq + w.ins_begin + k - 1, (size_t)w.ins_len); o += w.ins_len;
            *o++ = ':'; o = put_int(o, w.name_end);
        }
        fr.name_len = (uint32_t)(o - fr.name);
        *o++ = '\t';
        o = put_fixed(o, w.rvaf, 3); *o++ = '\t';
        o = put_fixed(o, w.expr, 1); *o++ = '\t';
        o = put_int(o, (long long)w.min_cov); *o++ = '\t';
        o = put_int(o, w.start_off); *o++ = '\t';
        if (w.var_end > w.var_begin) { const size_t n = (size_t)(w.var_end - w.var_begin + k - 1); memcpy(o, pseq + w.var_begin, n); o += n; }
        *o++ = '\t';
        o = put_fixed(o, w.ref_expr, 1); *o++ = '\t';
        if (w.ref_end > w.ref_begin) { const size_t n = (size_t)(w.ref_end - w.ref_begin + k - 1); memcpy(o, tseq + w.ref_begin, n); o += n; }
        *o++ = '\t';
        if (w.kind == 0) { memcpy(o, "vs_ref", 6); o += 6; }
        else { memcpy(o, "cluster ", 8); o += 8; o = put_int(o, w.cluster_id); memcpy(o, " n=", 3); o += 3; o = put_int(o, w.cluster_n); }
        *o++ = '\n';
        fr.line_len = (uint32_t)(o - fr.line);
        rows.push_back(fr);
    }
    // key = natsortkey(*info.split(' '), query, variant_name, type, min_coverage, rev_ix=[0]) (:825-829):
    // info is "vs_ref" or "cluster <i> n=<j>"; the first word compares REVERSED (vs_ref rows first),
    // then the words (numbers as numbers), the query (equal inside a target), the variant name, the
    // type, Min_coverage; a key that is a prefix of the other sorts first (vs_ref has one word, a
    // cluster three, but those never tie on the first word).
    if (nrow > 1) {
        auto less = [&](const FmtRow& x, const FmtRow& y) {
            const km_row& a = *x.w; const km_row& b = *y.w;
            if (a.kind != b.kind) return a.kind < b.kind;                       // "vs_ref" > "cluster", reversed
            if (a.kind != 0) {
                if (a.cluster_id != b.cluster_id) return a.cluster_id < b.cluster_id;
                if (a.cluster_n != b.cluster_n) return a.cluster_n < b.cluster_n;
            }
            int c = nat_cmp(x.name, x.name_len, y.name, y.name_len);
            if (c) return c < 0;
            c = nat_cmp(TYPE_NAME[a.type], strlen(TYPE_NAME[a.type]), TYPE_NAME[b.type], strlen(TYPE_NAME[b.type]));
            if (c) return c < 0;
            // Min_coverage prints as a decimal integer; the counts are never negative
            return a.min_cov < b.min_cov;
        };
        if (nrow <= 16) {                   // stable insertion sort: no temporary buffer for the usual 2-3 rows
            for (int i = 1; i < nrow; ++i) {
                FmtRow cur = rows[(size_t)i];
                int j = i;
                while (j > 0 && less(cur, rows[(size_t)j - 1])) { rows[(size_t)j] = rows[(size_t)j - 1]; --j; }
                rows[(size_t)j] = cur;
            }
        } else {
            std::stable_sort(rows.begin(), rows.end(), less);
        }
    }
    for (const FmtRow& f : rows) out.insert(out.end(), f.line, f.line + f.line_len);
}

extern "C" int64_t km_result_format_target(const km_result* r, int32_t tg, const char* db_name, const char* query_name, char* buf,
                                           int64_t buf_len) {
    if (!r || tg < 0 || tg >= r->n_targets || !db_name || !query_name) { fail(KM_E_ARG, "km_result_format_target: bad argument"); return -1; }
    std::vector<char> arena, text;
    std::vector<FmtRow> rows;
    format_rows_of(r, tg, db_name, strlen(db_name), query_name, strlen(query_name), arena, rows, text);
    const int64_t need = (int64_t)text.size();
    if (buf && need < buf_len) { memcpy(buf, text.data(), text.size()); buf[need] = 0; }
    return need;
}

// rows of targets [lo, hi) in order, appended to `out`
static void format_range(const km_result* r, int lo, int hi, const char* db_name, const char* names, const int64_t* name_off,
                         std::vector<char>& out) {
    const size_t db_len = strlen(db_name);
    std::vector<char> arena;
    std::vector<FmtRow> rows;
    size_t guess = 0;
    for (int t = lo; t < hi; ++t) guess += (size_t)r->row_count[t] * (size_t)(2 * (r->seq_off[t + 1] - r->seq_off[t]) + 160 + db_len);
    out.reserve(out.size() + guess);
    for (int t = lo; t < hi; ++t)
        format_rows_of(r, t, db_name, db_len, names + name_off[t], (size_t)(name_off[t + 1] - name_off[t]), arena, rows, out);
}

// Builds (once) the text of all targets in target order on host threads: every thread formats a
// contiguous range of targets into its own buffer, the pieces are then copied side by side.
static int build_text(const km_result* r, const char* db_name, const char* names, const int64_t* name_off, int32_t threads) {
    const int n = r->n_targets;
    {
        // the text at hand was made for (db_name, names)?  Compared in place: the key of a 10,000-target panel is 140 KB, and
        // this is on the path of every km_find_text + km_result_text pair
        const size_t dl = strlen(db_name), nb = n ? (size_t)name_off[n] : 0;
        const std::string& k = r->fmt_key;
        if (r->text_len >= 0 && k.size() == dl + 1 + nb && memcmp(k.data(), db_name, dl) == 0 && k[dl] == '\0' &&
            (nb == 0 || memcmp(k.data() + dl + 1, names, nb) == 0))
            return 0;
    }
    std::string key = std::string(db_name) + '\0' + (n ? std::string(names, (size_t)name_off[n]) : std::string());
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::max(1, n / 64)));
    std::vector<std::vector<char>> piece((size_t)nt);
    auto work = [&](int w) {
        const int lo = (int)((int64_t)n * w / nt), hi = (int)((int64_t)n * (w + 1) / nt);
        format_range(r, lo, hi, db_name, names, name_off, piece[(size_t)w]);
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nt; ++i) pool.emplace_back(work, i);
        for (auto& th : pool) th.join();
    }
    int64_t total = 0;
    std::vector<int64_t> at((size_t)nt);
    for (int i = 0; i < nt; ++i) { at[(size_t)i] = total; total += (int64_t)piece[(size_t)i].size(); }
    r->text.reset((size_t)total + 1);
    char* dst = r->text.get();
    auto copy = [&](int w) { if (!piece[(size_t)w].empty()) memcpy(dst + at[(size_t)w], piece[(size_t)w].data(), piece[(size_t)w].size()); };
    if (nt == 1) copy(0);
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nt; ++i) pool.emplace_back(copy, i);
        for (auto& th : pool) th.join();
    }
    dst[total] = 0;
    r->text_len = total;
    r->fmt_key.swap(key);
    return 0;
}

extern "C" int64_t km_result_format_all(const km_result* r, const char* db_name, const char* names, const int64_t* name_off,
                                        int32_t threads, char* buf, int64_t buf_len) {
    if (!r || !db_name || (r->n_targets && (!names || !name_off))) { fail(KM_E_ARG, "km_result_format_all: bad argument"); return -1; }
    build_text(r, db_name, names, name_off, threads);
    const int64_t need = r->text_len;
    if (buf && need < buf_len) memcpy(buf, r->text.get(), (size_t)need + 1);
    return need;
}

extern "C" int64_t km_result_text(const km_result* r, const char* db_name, const char* names, const int64_t* name_off,
                                  int32_t threads, const char** text) {
    if (!r || !db_name || !text || (r->n_targets && (!names || !name_off))) { fail(KM_E_ARG, "km_result_text: bad argument"); return -1; }
    build_text(r, db_name, names, name_off, threads);
    *text = r->text.get();
    return r->text_len;
}

// ---- pipelined batch -> text ---------------------------------------------------------------------
// A small persistent pool of host threads (thread creation costs more than formatting a sub-batch).
struct HostPool {
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::vector<std::thread> workers;
    bool stop = false;
    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i)
            workers.emplace_back([this] {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (stop && q.empty()) return;
                        job = std::move(q.front());
                        q.pop_front();
                    }
                    job();
                }
            });
    }
    void submit(std::function<void()> f) { { std::lock_guard<std::mutex> g(m); q.push_back(std::move(f)); } cv.notify_one(); }
    ~HostPool() { { std::lock_guard<std::mutex> g(m); stop = true; } cv.notify_all(); for (auto& w : workers) w.join(); }
};
static HostPool& host_pool() {
    static HostPool pool((int)std::max(2u, std::min(64u, std::thread::hardware_concurrency())));
    return pool;
}
struct Latch {
    std::mutex m; std::condition_variable cv; int left;
    explicit Latch(int n) : left(n) {}
    void done() { std::lock_guard<std::mutex> g(m); if (--left == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return left == 0; }); }
};

// km find_mutation for a whole batch, host buffers in, text out, as ONE call: the batch is cut into
// sub-batches that are all enqueued at once on their own streams; while the GPU works on the later
// ones the host formats the rows of the earlier ones (pool threads), so copies, kernels and text
// building overlap.  The text equals km_find_batch + km_result_format_all.
// KM_TRACE=1: host-clock timeline of km_find_text on stderr (measurement aid)
struct Trace {
    bool on, device;         // device (KM_TRACE=2): per-phase CUDA events on every lane as well (and no graph replay)
    std::chrono::steady_clock::time_point t0;
    std::mutex m;
    std::vector<std::tuple<const char*, int, double>> ev;
    Trace() : on(getenv("KM_TRACE") != nullptr), device(on && atoi(getenv("KM_TRACE")) >= 2), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what, int sub = -1) {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::lock_guard<std::mutex> g(m);
        ev.emplace_back(what, sub, ms);
    }
    ~Trace() {
        if (!on) return;
        for (auto& e : ev) fprintf(stderr, "[km_trace] %8.3f ms  %s %d\n", std::get<2>(e), std::get<0>(e), std::get<1>(e));
    }
};

extern "C" int km_find_text(km_table* t, const char* seqs, const int64_t* offsets, int32_t n, const km_find_params* params,
                            const char* db_name, const char* names, const int64_t* name_off, int32_t n_sub, km_result** out) {
    if (!t || !out || n < 0 || (n && (!seqs || !offsets || !names || !name_off)) || !params || !db_name)
        return fail(KM_E_ARG, "km_find_text: bad argument");
    CU(cudaSetDevice(t->device));
    if (n == 0) {                          // an empty batch: empty text, no status (offsets / names may be NULL)
        km_result* res = new km_result();
        res->n_targets = 0; res->k = t->k; res->has_graph = false;
        res->parts.emplace_back(new km_result());
        res->text.reset(1);
        res->text.get()[0] = 0;
        res->text_len = 0;
        res->fmt_key = std::string(db_name) + '\0';
        *out = res;
        return 0;
    }
    if (int rc = km_ensure_linked(t)) return rc;
    Trace tr;
    // (measured, 10,000 targets: 4 and 6 sub-batches tie on one GPU -- 1.407 / 1.408 ms -- and 4 wins when eight ranks share a
    // 32-core box, 1.91 vs 2.18 ms: fewer pool threads per process; profiles/r2m_nsub.txt, r2o_nsub_n8.txt)
    if (n_sub <= 0) n_sub = n >= 4096 ? 4 : n >= 1024 ? 2 : 1;
    n_sub = std::max(1, std::min(n_sub, std::max(1, n)));
    while ((int)t->lanes.size() < n_sub) {
        std::unique_ptr<km_table::Lane> L(new km_table::Lane());
        L->pin.host = true;
        // earlier sub-batches run at higher stream priority: they finish one after the other instead of all
        // together at the end, so the host can format the first while the GPU works on the rest
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));          // lo = least urgent (0), hi = most urgent (negative)
        // KM_LANE_PRIORITIES (experiments): 0 = all lanes alike, 2 = only lane 0 urgent; default: one level per lane
        static const int prio_mode = getenv("KM_LANE_PRIORITIES") ? atoi(getenv("KM_LANE_PRIORITIES")) : 1;
        const int lane_no = (int)t->lanes.size();
        const int prio = prio_mode == 0 ? prio_lo : prio_mode == 2 ? (lane_no == 0 ? prio_hi : prio_lo) : std::min(prio_lo, prio_hi + lane_no);
        CU(cudaStreamCreateWithPriority(&L->stream, cudaStreamNonBlocking, prio));
        CU(cudaStreamCreateWithPriority(&L->side, cudaStreamNonBlocking, prio));
        CU(cudaStreamCreateWithPriority(&L->side2, cudaStreamNonBlocking, prio));
        CU(cudaStreamCreateWithPriority(&L->side3, cudaStreamNonBlocking, prio));
        for (auto& e : L->ev) CU(cudaEventCreate(&e));
        CU(cudaEventCreateWithFlags(&L->fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&L->join, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&L->join2, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&L->join3, cudaEventDisableTiming));
        t->lanes.push_back(std::move(L));
    }
    // sub-batches balanced by sequence length (contiguous ranges)
    std::vector<int> cut(1, 0);
    const int64_t total = n ? offsets[n] - offsets[0] : 0;
    // Shares of the sub-batches: 2 : 3 : 4 : 5 ... -- the first (most urgent streams) is the smallest, so the sub-batches finish
    // one after the other and each text is copied while the GPU still works on the rest; with equal shares they finish
    // together and the copies queue up at the end (10,000 targets, 4 sub-batches: 1.28 -> 1.18 ms, profiles/r3o_sub_weights.txt).
    // KM_SUB_WEIGHTS=1,1,1,1 overrides (experiments).
    std::vector<double> cum((size_t)n_sub + 1, 0.0);
    {
        std::vector<double> w((size_t)n_sub, 1.0);
        for (int c = 0; c < n_sub; ++c) w[(size_t)c] = 2.0 + (double)c;
        if (const char* e = getenv("KM_SUB_WEIGHTS")) {
            size_t i = 0;
            for (const char* q = e; *q && i < w.size(); ++i) { w[i] = std::max(0.01, atof(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
        }
        for (int c = 0; c < n_sub; ++c) cum[(size_t)c + 1] = cum[(size_t)c] + w[(size_t)c];
    }
    for (int c = 1; c < n_sub; ++c) {
        const int64_t want = offsets[0] + (int64_t)((double)total * (cum[(size_t)c] / cum[(size_t)n_sub]));
        int i = (int)(std::lower_bound(offsets, offsets + n + 1, want) - offsets);
        i = std::max(cut.back(), std::min(i, n));
        cut.push_back(i);
    }
    cut.push_back(n);
    km_result* res = new km_result();
    res->n_targets = n; res->k = t->k; res->has_graph = false;
    km_find_params prm = *params;
    prm.flags |= KM_FIND_NO_GRAPH;
    if (!getenv("KM_HOST_FORMAT")) {
        // ---- the text is formatted on the device (format.h) --------------------------------------------------
        // One pool task per sub-batch: set-up + upload + launches, a small fetch (per-target status, length of
        // the text), then -- once the lengths of the sub-batches before it are known -- ONE copy of its text
        // straight to its place in the result's pinned buffer.  The host formats nothing and joins nothing.
        const std::string db(db_name);
        std::vector<int64_t> caps((size_t)n_sub);
        int64_t cap_total = 1;
        for (int c = 0; c < n_sub; ++c) {
            const int lo = cut[(size_t)c], hi = cut[(size_t)c + 1];
            const int64_t rows = std::max(64, 16 * (hi - lo));
            caps[(size_t)c] = 16 * (offsets[hi] - offsets[lo]) + 512ll * (hi - lo) + rows * (int64_t)(db.size() + 64) + (1 << 16);
            cap_total += caps[(size_t)c];
        }
        if (int rc = res->text.pin.reserve(t->pool, (size_t)cap_total)) { delete res; return rc; }
        char* final_text = res->text.pin.base;
        std::vector<std::unique_ptr<km_plan>> plans((size_t)n_sub);
        std::vector<std::vector<int64_t>> offs((size_t)n_sub), noffs((size_t)n_sub);
        std::vector<std::vector<char>> spill((size_t)n_sub);            // text that did not go straight to its place
        std::vector<char> spilled((size_t)n_sub, 0);
        res->parts.resize((size_t)n_sub);
        std::vector<int> rcs((size_t)n_sub, 0);
        std::vector<std::string> errs((size_t)n_sub);
        std::vector<long long> lens((size_t)n_sub, -1);
        std::mutex lm; std::condition_variable lcv;
        Latch latch(n_sub);
        const int device = t->device;
        std::atomic<int> next_lane(0);
        for (int c = 0; c < n_sub; ++c) {
            host_pool().submit([=, &db, &next_lane, &cut, &plans, &offs, &noffs, &spill, &spilled, &rcs, &errs, &lens, &lm, &lcv, &latch, &prm, &tr] {
                const int lo = cut[(size_t)c], hi = cut[(size_t)c + 1];
                auto publish = [&](long long len) { { std::lock_guard<std::mutex> g(lm); lens[(size_t)c] = len; } lcv.notify_all(); };
                auto fail_all = [&](int rc) { rcs[(size_t)c] = rc; errs[(size_t)c] = g_err; publish(0); latch.done(); };
                if (cudaSetDevice(device) != cudaSuccess) { fail(KM_E_CUDA, "cudaSetDevice failed"); return fail_all(KM_E_CUDA); }
                auto& o = offs[(size_t)c]; auto& no = noffs[(size_t)c];
                o.resize((size_t)(hi - lo) + 1); no.resize((size_t)(hi - lo) + 1);
                for (int i = lo; i <= hi; ++i) { o[(size_t)(i - lo)] = offsets[i] - offsets[lo]; no[(size_t)(i - lo)] = name_off[i] - name_off[lo]; }
                plans[(size_t)c].reset(new km_plan());
                km_plan* p = plans[(size_t)c].get();
                p->fmt = true; p->fmt_names = names + name_off[lo]; p->fmt_name_off = no.data(); p->fmt_db = db;
                p->targets_ext = seqs + offsets[lo];
                // sub-batch c always runs on lane c: the same streams (earlier sub-batches more urgent), the same arenas --
                // and therefore, for a batch of the same layout, the same enqueue, which is what makes the graph replay below
                const int lane_ix = c;
                (void)next_lane;
                g_trace_obj = &tr; g_trace_sub = c;
                g_trace_mark = tr.on ? +[](void* o, const char* w, int sub) { static_cast<Trace*>(o)->mark(w, sub); } : nullptr;
                tr.mark("task start", c);
                p->defer_upload = true;
                if (int rc = plan_init(t, seqs + offsets[lo], o.data(), hi - lo, &prm, p, false, t->lanes[(size_t)lane_ix].get())) return fail_all(rc);
                tr.mark("plan_init", c);
                {
                    // Enqueued directly this is ~17 driver calls per sub-batch, ~6 us each whether one thread issues them or
                    // six do at once.  When the lane saw the very same layout in the previous call, the sequence is captured
                    // into a CUDA graph (once) and from then on replayed with ONE call.
                    km_table::Lane* lane = t->lanes[(size_t)lane_ix].get();
                    static const bool graphs_on = !getenv("KM_NO_GRAPH");
                    const bool use_graph = graphs_on && !tr.device;
                    p->trace_events = tr.device;
                    const std::string key = use_graph ? plan_graph_key(p) : std::string();
                    int rc = 0;
                    bool done = false;
                    if (use_graph && lane->gexec && lane->gkey == key) {
                        if (cudaGraphLaunch(lane->gexec, p->stream) == cudaSuccess) { done = true; p->launched = true; p->n_launches += km_bubble_pass_enabled() ? 11 : 9; tr.mark("graph replay", c); }
                        else { cudaGetLastError(); cudaGraphExecDestroy(lane->gexec); lane->gexec = nullptr; }
                    } else if (use_graph && lane->last_key == key) {
                        // (one capture at a time: it happens once per layout, and four threads capturing and instantiating at
                        // once is the one thing this call does that profilers and the driver see rarely -- an ncu run of
                        // bench.py once died there)
                        static std::mutex capture_mutex;
                        std::lock_guard<std::mutex> capture_lock(capture_mutex);
                        if (lane->gexec) { cudaGraphExecDestroy(lane->gexec); lane->gexec = nullptr; }
                        cudaGraph_t graph = nullptr;
                        if (cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                            rc = plan_upload_enqueue(p, p->stream);
                            if (!rc) rc = plan_launch(p, p->stream);
                            const cudaError_t ce = cudaStreamEndCapture(p->stream, &graph);
                            if (!rc && ce == cudaSuccess && graph && cudaGraphInstantiate(&lane->gexec, graph, 0) == cudaSuccess) {
                                lane->gkey = key;
                                if (cudaGraphLaunch(lane->gexec, p->stream) == cudaSuccess) { done = true; tr.mark("graph captured", c); }
                            }
                            if (graph) cudaGraphDestroy(graph);
                            if (!done) { cudaGetLastError(); if (lane->gexec) { cudaGraphExecDestroy(lane->gexec); lane->gexec = nullptr; } rc = 0; p->launched = false; }
                        } else cudaGetLastError();
                    }
                    if (use_graph) lane->last_key = key;
                    if (!done) {
                        rc = plan_upload_enqueue(p, p->stream);
                        if (!rc) rc = plan_launch(p, p->stream);
                    }
                    if (rc) return fail_all(rc);
                }
                p->defer_upload = false;
                tr.mark("plan_launch", c);
                std::unique_ptr<km_result> part(new km_result());
                part->seq_off = p->seq_off;
                if (int rc = plan_fetch(p, part.get(), false, true)) return fail_all(rc);
                tr.mark("head fetched", c);
                km_result* pr = part.get();
                res->parts[(size_t)c] = std::move(part);
                long long len = pr->dev_text_len;
                const bool host_format = pr->dev_text_flags != 0;
                if (host_format) {        // the device declined (capacity, or a number it does not print): rows come back, host formats
                    if (int rc = plan_download(p, p->stream, pr, false)) return fail_all(rc);
                    pr->targets.assign(p->targets_ext, (size_t)p->n_code);
                    format_range(pr, 0, pr->n_targets, db.c_str(), names + name_off[lo], no.data(), spill[(size_t)c]);
                    len = (long long)spill[(size_t)c].size();
                }
                publish(len);
                long long at = 0;
                {
                    std::unique_lock<std::mutex> lk(lm);
                    lcv.wait(lk, [&] { for (int j = 0; j < c; ++j) if (lens[(size_t)j] < 0) return false; return true; });
                    for (int j = 0; j < c; ++j) at += lens[(size_t)j];
                }
                const bool fits = at + len + 1 <= cap_total;
                if (host_format) {
                    if (fits) { memcpy(final_text + at, spill[(size_t)c].data(), (size_t)len); spill[(size_t)c].clear(); }
                    else spilled[(size_t)c] = 1;
                } else if (len) {
                    char* dst = final_text + at;
                    if (!fits) { spill[(size_t)c].resize((size_t)len); dst = spill[(size_t)c].data(); spilled[(size_t)c] = 1; }
                    if (cudaMemcpyAsync(dst, p->F.text, (size_t)len, cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
                        km_wait_stream(p->stream, p->wait_ev) != cudaSuccess) {
                        fail(KM_E_CUDA, "copy of the text failed: %s", cudaGetErrorString(cudaGetLastError()));
                        rcs[(size_t)c] = KM_E_CUDA; errs[(size_t)c] = g_err;
                    }
                    pr->bytes_d2h += (unsigned long long)len;
                }
                if (tr.device) cudaEventRecord(p->ev[5], p->stream);
                tr.mark("text placed", c);
                plan_return_vecs(p, t->lanes[(size_t)lane_ix].get());
                latch.done();
            });
        }
        tr.mark("all submitted");
        // (while the pool works: the key km_result_text recognises this text by)
        res->fmt_key = std::string(db_name) + '\0' + (n ? std::string(names, (size_t)name_off[n]) : std::string());
        latch.wait();
        tr.mark("all placed");
        if (tr.device) {
            // device timeline: every lane's phase events against the earliest upload
            cudaDeviceSynchronize();
            static const char* what[] = {"upload", "probe", "walks", "graph passes", "format", "text copied"};
            static const int evi[] = {0, 1, 6, 2, 3, 4, 5};
            int first = 0;
            for (int c = 1; c < n_sub; ++c) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, t->lanes[(size_t)first]->ev[0], t->lanes[(size_t)c]->ev[0]) == cudaSuccess && ms < 0.f) first = c;
            }
            for (int c = 0; c < n_sub; ++c) {
                fprintf(stderr, "[km_trace] device, sub-batch %d:", c);
                for (int j = 0; j < 7; ++j) {
                    float ms = 0.f;
                    if (cudaEventElapsedTime(&ms, t->lanes[(size_t)first]->ev[0], t->lanes[(size_t)c]->ev[evi[j]]) != cudaSuccess) { cudaGetLastError(); ms = -1.f; }
                    if (j == 0) fprintf(stderr, " start %.3f", ms); else fprintf(stderr, " | %s done %.3f", what[j - 1], ms);
                }
                fprintf(stderr, " ms\n");
            }
        }
        for (int c = 0; c < n_sub; ++c)
            if (rcs[(size_t)c]) {
                const int rc = rcs[(size_t)c];
                fail(rc, "%s", errs[(size_t)c].c_str());
                delete res;
                return rc;
            }
        long long len = 0;
        bool any_spill = false;
        for (int c = 0; c < n_sub; ++c) { len += lens[(size_t)c]; any_spill |= spilled[(size_t)c] != 0; }
        if (any_spill) {
            // (only after capacity retries grew a sub-batch beyond the estimate) assemble in a buffer of the exact size
            PinBlock exact;
            if (int rc = exact.reserve(t->pool, (size_t)len + 1)) { delete res; return rc; }
            long long at = 0;
            for (int c = 0; c < n_sub; ++c) {
                const long long l = lens[(size_t)c];
                if (spilled[(size_t)c]) memcpy(exact.base + at, spill[(size_t)c].data(), (size_t)l);
                else if (at + l + 1 <= cap_total) memcpy(exact.base + at, final_text + at, (size_t)l);
                at += l;
            }
            res->text.pin.drop();
            res->text.pin.pool = exact.pool; res->text.pin.base = exact.base; res->text.pin.cap = exact.cap;
            exact.base = nullptr;
        }
        res->text.get()[len] = 0;
        res->text_len = len;
        for (auto& part : res->parts) {
            res->all_status.insert(res->all_status.end(), part->status.data(), part->status.data() + part->status.size());
            res->ms_h2d += part->ms_h2d; res->ms_walk += part->ms_walk; res->ms_graph += part->ms_graph; res->ms_d2h += part->ms_d2h;
            res->n_launches += part->n_launches; res->n_retries += part->n_retries;
            res->bytes_h2d += part->bytes_h2d; res->bytes_d2h += part->bytes_d2h;
        }
        *out = res;
        tr.mark("done");
        return 0;
    }
    // ---- KM_HOST_FORMAT: rows come back, host threads format them -------------------------------------------
    // One pool task per sub-batch: layout + upload + launch, then the fetch (which waits for that sub-batch's
    // stream only), then its rows go to the pool in slices.  Nothing on the pool waits for another pool task; the
    // caller waits for the last slice.  Enqueueing from several threads at once keeps the host off the critical
    // path: done one after the other the six set-ups alone took as long as all the kernels.
    const int n_slice = std::max(4, std::min(16, (int)host_pool().workers.size() / 2));
    std::vector<std::vector<char>> piece((size_t)n_sub * n_slice);
    std::vector<std::unique_ptr<km_plan>> plans((size_t)n_sub);
    std::vector<std::vector<int64_t>> offs((size_t)n_sub), noffs((size_t)n_sub);
    res->parts.resize((size_t)n_sub);
    std::vector<int> rcs((size_t)n_sub, 0);
    std::vector<std::string> errs((size_t)n_sub);
    Latch latch(n_sub * n_slice);
    const int device = t->device;
    std::atomic<int> next_lane(0);
    for (int c = 0; c < n_sub; ++c) {
        host_pool().submit([=, &next_lane, &cut, &piece, &plans, &offs, &noffs, &rcs, &errs, &latch, &prm, &tr] {
            const int lo = cut[(size_t)c], hi = cut[(size_t)c + 1];
            auto fail_all = [&](int rc) { rcs[(size_t)c] = rc; errs[(size_t)c] = g_err; for (int j = 0; j < n_slice; ++j) latch.done(); };
            if (cudaSetDevice(device) != cudaSuccess) { fail(KM_E_CUDA, "cudaSetDevice failed"); return fail_all(KM_E_CUDA); }
            auto& o = offs[(size_t)c]; auto& no = noffs[(size_t)c];
            o.resize((size_t)(hi - lo) + 1); no.resize((size_t)(hi - lo) + 1);
            for (int i = lo; i <= hi; ++i) { o[(size_t)(i - lo)] = offsets[i] - offsets[lo]; no[(size_t)(i - lo)] = name_off[i] - name_off[lo]; }
            plans[(size_t)c].reset(new km_plan());
            km_plan* p = plans[(size_t)c].get();
            tr.mark("task start", c);
            // lanes are handed out in the order the tasks get here: the first one to enqueue has the most urgent streams
            const int lane_ix = next_lane.fetch_add(1);
            if (int rc = plan_init(t, seqs + offsets[lo], o.data(), hi - lo, &prm, p, false, t->lanes[(size_t)lane_ix].get())) return fail_all(rc);
            tr.mark("plan_init", c);
            if (int rc = plan_launch(p, p->stream)) return fail_all(rc);
            tr.mark("plan_launch", c);
            std::unique_ptr<km_result> part(new km_result());
            part->seq_off = p->seq_off;
            if (int rc = plan_fetch(p, part.get(), false)) return fail_all(rc);
            tr.mark("plan_fetch", c);
            part->targets.swap(p->targets);
            km_result* pr = part.get();
            res->parts[(size_t)c] = std::move(part);
            const char* nm = names + name_off[lo];
            const int64_t* nop = no.data();
            const int m = pr->n_targets;
            for (int j = 0; j < n_slice; ++j) {
                std::vector<char>* dst = &piece[(size_t)c * n_slice + (size_t)j];
                const int a = (int)((int64_t)m * j / n_slice), b = (int)((int64_t)m * (j + 1) / n_slice);
                host_pool().submit([pr, db_name, nm, nop, a, b, dst, &latch, &tr, c] {
                    *dst = piece_cache().get(0);
                    format_range(pr, a, b, db_name, nm, nop, *dst);
                    tr.mark("slice", c);
                    latch.done();
                });
            }
        });
    }
    tr.mark("all submitted");
    latch.wait();
    tr.mark("formatted");
    for (int c = 0; c < n_sub; ++c)
        if (rcs[(size_t)c]) {
            const int rc = rcs[(size_t)c];
            fail(rc, "%s", errs[(size_t)c].c_str());
            for (auto& pc : piece) piece_cache().put(std::move(pc));
            delete res;
            return rc;
        }
    int64_t len = 0;
    std::vector<int64_t> at_of;
    for (auto& pc : piece) { at_of.push_back(len); len += (int64_t)pc.size(); }
    res->text.reset((size_t)len + 1);
    {
        Latch joined((int)piece.size());
        char* dst = res->text.get();
        for (size_t c = 0; c < piece.size(); ++c) {
            std::vector<char>* pc = &piece[c];
            const int64_t at = at_of[c];
            host_pool().submit([pc, dst, at, &joined] {
                if (!pc->empty()) memcpy(dst + at, pc->data(), pc->size());
                piece_cache().put(std::move(*pc));
                joined.done();
            });
        }
        joined.wait();
    }
    tr.mark("joined");
    for (auto& part : res->parts) {
        res->all_status.insert(res->all_status.end(), part->status.data(), part->status.data() + part->status.size());
        res->ms_h2d += part->ms_h2d; res->ms_walk += part->ms_walk; res->ms_graph += part->ms_graph; res->ms_d2h += part->ms_d2h;
        res->n_launches += part->n_launches; res->n_retries += part->n_retries;
        res->bytes_h2d += part->bytes_h2d; res->bytes_d2h += part->bytes_d2h;
    }
    res->text.get()[len] = 0;
    res->text_len = len;
    res->fmt_key = std::string(db_name) + '\0' + (n ? std::string(names, (size_t)name_off[n]) : std::string());
    *out = res;
    tr.mark("done");
    return 0;
}

extern "C" int km_debug_format_fixed(double v, int prec, char* buf64) {
    if (!buf64 || prec < 0 || prec > 3) return fail(KM_E_ARG, "km_debug_format_fixed: bad argument");
    char* e = put_fixed(buf64, v, prec);
    *e = 0;
    return (int)(e - buf64);
}
extern "C" int km_debug_nat_cmp(const char* a, const char* b) { return nat_cmp(a, strlen(a), b, strlen(b)); }

