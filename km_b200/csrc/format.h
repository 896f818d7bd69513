// The text `km find_mutation` prints, built ON THE DEVICE: PathQuant.Path.__str__ (PathQuant.py:37-49) for
// every row and MutationFinder.get_paths' sort (MutationFinder.py:813-833, common.natsortkey common.py:95-116)
// inside every target.  km_find_text uses it so that the host neither formats nor joins anything: the text
// of a sub-batch comes back with one copy, straight into its place in the caller's buffer.  (With the rows
// formatted by host threads, eight ranks on one box spent 14 ms of CPU per 10,000-target panel each and the
// end-to-end rate stopped scaling at ~11 M targets/s; the kernels themselves scale linearly.)
//
// Three launches: measure (one lane per row: exact length of the line, sort rank inside the target), scan
// (exclusive prefix of the per-target lengths), write (one warp per target; numbers by every lane, sequences
// copied 32 bytes at a time).  The same formatter runs in both, behind a counting and a storing emitter.
// It is the device twin of format_rows_of / put_fixed / nat_cmp in api.cu; tests compare the two byte by byte.
#pragma once
#include "graph.h"

namespace km {

struct FormatView {
    const char* db_name; int db_len;
    const char* names; const int64_t* name_off;      // query names of the targets
    int32_t* row_len;        // [row_cap] length of the row's line
    int32_t* row_pos;        // [row_cap] offset of the line inside its target's block (sorted order)
    char* row_num;           // [row_cap * KM_FMT_ROW_BYTES] the row's numbers as text (measure writes, write copies)
    int64_t* t_len;          // [n] bytes of the target's block
    int64_t* t_off;          // [n + 1] exclusive prefix; t_off[n] = total
    char* text; int64_t text_cap;
    uint32_t* flags;         // [0] bit 0: text_cap exceeded, bit 1: a value the device formatter does not print;
                             // [2..3] total bytes of text (64 bit), for the host
};

// per-row scratch: 9 numeric fields of up to 24 characters + their lengths
#define KM_FMT_FIELDS 9
#define KM_FMT_FIELD_BYTES 24
#define KM_FMT_ROW_BYTES 256

// (the device code below is compiled by format_kernels.cu only; the other translation units see FormatView)
#if KM_DEVICE_BUILD && defined(KM_FORMAT_KERNELS)

// decimal digits of v straight into buf (no temporary: a local array indexed by a loop counter lives in local
// memory, and every digit then costs a round trip to L1); returns their number
__device__ __forceinline__ int fmt_uint(char* buf, unsigned long long v) {
    if (v <= 0xFFFFFFFFull) {
        unsigned int x = (unsigned int)v;
        const int n = x < 10u ? 1 : x < 100u ? 2 : x < 1000u ? 3 : x < 10000u ? 4 : x < 100000u ? 5 : x < 1000000u ? 6 :
                      x < 10000000u ? 7 : x < 100000000u ? 8 : x < 1000000000u ? 9 : 10;
        for (int i = n - 1; i >= 0; --i) { buf[i] = (char)('0' + (int)(x % 10u)); x /= 10u; }
        return n;
    }
    int n = 0;
    for (unsigned long long t = v; t; t /= 10ull) ++n;            // (rare: 64-bit division is emulated)
    for (int i = n - 1; i >= 0; --i) { buf[i] = (char)('0' + (int)(v % 10ull)); v /= 10ull; }
    return n;
}
__device__ __forceinline__ int fmt_int(char* buf, long long v) {
    if (v < 0) { buf[0] = '-'; return 1 + fmt_uint(buf + 1, 0ull - (unsigned long long)v); }
    return fmt_uint(buf, (unsigned long long)v);
}
// "%.{prec}f" (prec <= 3): the exact binary value rounded half-to-even at the last printed digit, decided on
// integers (|v| = m * 2^e, m < 2^53, m * 10^prec fits 64 bits).  Returns -1 for magnitudes >= 2^52.
__device__ __forceinline__ int fmt_fixed(char* buf, double v, int prec) {
    if (isnan(v)) { buf[0] = 'n'; buf[1] = 'a'; buf[2] = 'n'; return 3; }
    int n = 0;
    if (isinf(v)) { if (v < 0) buf[n++] = '-'; buf[n++] = 'i'; buf[n++] = 'n'; buf[n++] = 'f'; return n; }
    const double a = fabs(v);
    if (a >= 4503599627370496.0) return -1;
    if (signbit(v)) buf[n++] = '-';
    const unsigned long long p10 = prec == 0 ? 1ull : prec == 1 ? 10ull : prec == 2 ? 100ull : 1000ull;
    unsigned long long q = 0;
    if (a != 0.0) {
        int e;
        const double fr = frexp(a, &e);
        const unsigned long long m = (unsigned long long)ldexp(fr, 53);
        const int sh = 53 - e;                       // a = m * 2^-sh
        const unsigned long long scaled = m * p10;
        if (sh <= 0) q = scaled << (-sh);
        else if (sh >= 64) q = 0;
        else {
            q = scaled >> sh;
            const unsigned long long rem = scaled & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
            if (rem > half || (rem == half && (q & 1ull))) ++q;
        }
    }
    unsigned long long ip; unsigned int fp;
    if (q <= 0xFFFFFFFFull) { const unsigned int q32 = (unsigned int)q, p32 = (unsigned int)p10; ip = q32 / p32; fp = q32 % p32; }
    else { ip = q / p10; fp = (unsigned int)(q % p10); }
    n += fmt_uint(buf + n, ip);
    if (prec > 0) {
        buf[n++] = '.';
        for (unsigned int d = (unsigned int)p10 / 10u; d; d /= 10u) { buf[n++] = (char)('0' + (int)((fp / d) % 10u)); }
    }
    return n;
}

__device__ const char km_type_name[6][16] = {"Reference", "Substitution", "ITD", "Indel", "Insertion", "Deletion"};
__device__ const int km_type_len[6] = {9, 12, 3, 5, 9, 8};

// emitters: where the formatter's output goes
struct CountEmit {                       // one lane, nothing stored
    long long n = 0;
    __device__ __forceinline__ void small(const char*, int len) { n += len; }
    __device__ __forceinline__ void ch(char) { n += 1; }
    __device__ __forceinline__ void raw(const char*, int len) { n += len; }
    __device__ __forceinline__ void codes(const uint8_t*, int len, bool) { n += len; }
};
struct WarpEmit {                        // the whole warp on one row: all lanes call with the same arguments
    char* dst; long long n = 0; int lane;
    __device__ __forceinline__ void small(const char* buf, int len) { if (lane < len) dst[n + lane] = buf[lane]; n += len; }   // len <= 32
    __device__ __forceinline__ void ch(char c) { if (lane == 0) dst[n] = c; n += 1; }
    // four loads in flight per lane: a copy is a chain of global round trips otherwise
    __device__ __forceinline__ void raw(const char* src, int len) {
        for (int i = lane; i < len; i += 128) {
            char v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = i + 32 * u < len ? src[i + 32 * u] : (char)0;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i + 32 * u < len) dst[n + i + 32 * u] = v[u];
        }
        n += len;
    }
    __device__ __forceinline__ void codes(const uint8_t* src, int len, bool lower) {
        const char* alphabet = lower ? "acgt" : "ACGT";
        for (int i = lane; i < len; i += 128) {
            uint8_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = i + 32 * u < len ? src[i + 32 * u] : (uint8_t)0;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i + 32 * u < len) dst[n + i + 32 * u] = alphabet[v[u] & 3];
        }
        n += len;
    }
};

// The numbers of a row as text.  The measure pass (one lane per row) prints them once -- into the row's scratch --
// and the write pass (a whole warp per row) only copies: printing them there again would be the same work done
// by 32 lanes at once.  Fields: 0 name_start, 1 name_end, 2 rVAF, 3 expression, 4 min_cov, 5 start_off,
// 6 ref_expression, 7 cluster id, 8 cluster size.
struct PrintNums {
    char* store;                 // the row's scratch, or nullptr
    bool ok = true;
    __device__ __forceinline__ int get(int field, const Row& w, char* buf, const char** out) {
        int n;
        char* dst = store ? store + KM_FMT_FIELD_BYTES * field : buf;       // (the longest: '-' + 16 digits + ".ddd" = 21)
        switch (field) {
            case 0: n = fmt_int(dst, w.name_start); break;
            case 1: n = fmt_int(dst, w.name_end); break;
            case 2: n = fmt_fixed(dst, w.rvaf, 3); break;
            case 3: n = fmt_fixed(dst, w.expr, 1); break;
            case 4: n = fmt_int(dst, (long long)w.min_cov); break;
            case 5: n = fmt_int(dst, w.start_off); break;
            case 6: n = fmt_fixed(dst, w.ref_expr, 1); break;
            case 7: n = fmt_int(dst, w.cluster_id); break;
            default: n = fmt_int(dst, w.cluster_n); break;
        }
        if (n < 0) { ok = false; n = 0; }
        if (store) store[KM_FMT_FIELD_BYTES * KM_FMT_FIELDS + field] = (char)n;
        *out = dst;
        return n;
    }
};
struct StoredNums {
    const char* store;
    bool ok = true;
    __device__ __forceinline__ int get(int field, const Row&, char*, const char** out) {
        *out = store + KM_FMT_FIELD_BYTES * field;
        return (int)store[KM_FMT_FIELD_BYTES * KM_FMT_FIELDS + field];
    }
};

// one line: "{db}\t{query}\t{type}\t{name}\t{rVAF:.3f}\t{expr:.1f}\t{min_cov}\t{start_off}\t{seq}\t{ref_expr:.1f}\t{ref_seq}\t{info}\n"
// Returns false if a number could not be printed here.
template <class E, class N>
__device__ __forceinline__ bool format_row(E& e, N& nums, const FormatView& F, const WalkView& W, const ResultView& R, int k, const Row& w) {
    const int t = w.target;
    const uint8_t* tcodes = W.codes + W.seq_off[t];
    const char* pseq = R.seq_pool + R.path_seq_off[w.path_id];
    char buf[40];
    const char* s;
    int n;
    e.raw(F.db_name, F.db_len); e.ch('\t');
    e.raw(F.names + F.name_off[t], (int)(F.name_off[t + 1] - F.name_off[t])); e.ch('\t');
    e.small(km_type_name[w.type], km_type_len[w.type]); e.ch('\t');
    if (w.type != 0) {                 // "{}:{}/{}:{}" (MutationFinder.py:483-488); Reference -> empty name
        n = nums.get(0, w, buf, &s); e.small(s, n); e.ch(':');
        e.codes(tcodes + w.del_begin + k - 1, w.del_len, true); e.ch('/');
        e.raw(pseq + w.ins_begin + k - 1, w.ins_len); e.ch(':');
        n = nums.get(1, w, buf, &s); e.small(s, n);
    }
    e.ch('\t');
    n = nums.get(2, w, buf, &s); e.small(s, n); e.ch('\t');
    n = nums.get(3, w, buf, &s); e.small(s, n); e.ch('\t');
    n = nums.get(4, w, buf, &s); e.small(s, n); e.ch('\t');
    n = nums.get(5, w, buf, &s); e.small(s, n); e.ch('\t');
    if (w.var_end > w.var_begin) e.raw(pseq + w.var_begin, w.var_end - w.var_begin + k - 1);
    e.ch('\t');
    n = nums.get(6, w, buf, &s); e.small(s, n); e.ch('\t');
    if (w.ref_end > w.ref_begin) e.codes(tcodes + w.ref_begin, w.ref_end - w.ref_begin + k - 1, false);
    e.ch('\t');
    if (w.kind == 0) e.small("vs_ref", 6);
    else {
        e.small("cluster ", 8); n = nums.get(7, w, buf, &s); e.small(s, n);
        e.small(" n=", 3); n = nums.get(8, w, buf, &s); e.small(s, n);
    }
    e.ch('\n');
    return nums.ok;
}

// natsortkey of the variant name without spelling it: the name is "" (Reference) or
// "<start>:<deleted, lower case>/<inserted>:<end>", whose token list is ["", start, ":del/ins:", end, ""] -- the
// bases hold no digits.  Text compares lower-cased, numbers as integers, a prefix sorts first.
__device__ __forceinline__ int name_text_at(const WalkView& W, const ResultView& R, int k, const Row& w, int i) {
    // character i of ":<del>/<ins>:" lower-cased, -1 past its end
    if (i == 0) return ':';
    i -= 1;
    if (i < w.del_len) return "acgt"[W.codes[W.seq_off[w.target] + w.del_begin + k - 1 + i] & 3];
    i -= w.del_len;
    if (i == 0) return '/';
    i -= 1;
    if (i < w.ins_len) {
        const char c = R.seq_pool[R.path_seq_off[w.path_id] + w.ins_begin + k - 1 + i];
        return c >= 'A' && c <= 'Z' ? c + 32 : c;
    }
    i -= w.ins_len;
    return i == 0 ? ':' : -1;
}
__device__ __forceinline__ int name_cmp(const WalkView& W, const ResultView& R, int k, const Row& a, const Row& b) {
    const bool ea = a.type == 0, eb = b.type == 0;
    if (ea || eb) return ea == eb ? 0 : (ea ? -1 : 1);
    if (a.name_start != b.name_start) return a.name_start < b.name_start ? -1 : 1;      // (never negative: start + k + offset)
    // the vs_ref row and the cluster row of one variant spell the same bases: no need to walk them
    const bool same_text = a.target == b.target && a.path_id == b.path_id && a.del_begin == b.del_begin && a.del_len == b.del_len &&
                           a.ins_begin == b.ins_begin && a.ins_len == b.ins_len;
    if (!same_text) for (int i = 0;; ++i) {
        const int ca = name_text_at(W, R, k, a, i), cb = name_text_at(W, R, k, b, i);
        if (ca != cb) return ca < cb ? -1 : 1;           // -1 (end) sorts first: the shorter text is a prefix
        if (ca < 0) break;
    }
    if (a.name_end != b.name_end) return a.name_end < b.name_end ? -1 : 1;
    return 0;
}
__device__ __forceinline__ int type_cmp(int ta, int tb) {           // natsort of the type names (no digits): lower-cased text
    if (ta == tb) return 0;
    const char* a = km_type_name[ta]; const char* b = km_type_name[tb];
    for (int i = 0;; ++i) {
        int ca = i < km_type_len[ta] ? a[i] : -1, cb = i < km_type_len[tb] ? b[i] : -1;
        if (ca >= 'A' && ca <= 'Z') ca += 32;
        if (cb >= 'A' && cb <= 'Z') cb += 32;
        if (ca != cb) return ca < cb ? -1 : 1;
        if (ca < 0) return 0;
    }
}
// the order of MutationFinder.get_paths (:825-829): "vs_ref" rows first (first word reversed), clusters by
// number then size, then variant name, type, Min_coverage
__device__ __forceinline__ bool row_less(const WalkView& W, const ResultView& R, int k, const Row& a, const Row& b) {
    if (a.kind != b.kind) return a.kind < b.kind;
    if (a.kind != 0) {
        if (a.cluster_id != b.cluster_id) return a.cluster_id < b.cluster_id;
        if (a.cluster_n != b.cluster_n) return a.cluster_n < b.cluster_n;
    }
    int c = name_cmp(W, R, k, a, b);
    if (c) return c < 0;
    c = type_cmp(a.type, b.type);
    if (c) return c < 0;
    return a.min_cov < b.min_cov;
}

// one WARP per target: lane r measures row r and finds its place among the target's rows
__global__ void __launch_bounds__(128) km_format_measure_kernel(WalkView W, ResultView R, FormatView F, int k) {
    const int t = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= W.n_targets) return;
    int nrow = R.t_n_rows[t];
    const int first = R.t_row_first[t];
    long long total = 0;
    bool ok = true;
    // rows or spelled paths that did not fit their pools (the host grows the pools and runs the batch again)
    if (first < 0 || nrow < 0 || (long long)first + nrow > (long long)R.row_cap) { nrow = 0; ok = false; }
    for (int r = lane; r < nrow; r += 32) {
        const Row w = R.rows[first + r];
        if (w.path_id < 0 || w.path_id >= R.path_cap || R.path_seq_off[w.path_id] < 0 || w.type < 0 || w.type > 5 ||
            (w.type != 0 && (w.name_start < 0 || w.name_end < 0))) { ok = false; F.row_len[first + r] = 0; continue; }
        CountEmit e;
        PrintNums nums;
        nums.store = F.row_num + (size_t)KM_FMT_ROW_BYTES * (size_t)(first + r);
        ok &= format_row(e, nums, F, W, R, k, w);
        F.row_len[first + r] = (int32_t)e.n;
        total += e.n;
    }
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
    ok = __all_sync(0xFFFFFFFFu, ok);
    if (!ok) { if (lane == 0) { atomicOr(F.flags, 2u); F.t_len[t] = 0; } return; }
    __syncwarp();
    for (int r = lane; r < nrow; r += 32) {
        const Row w = R.rows[first + r];
        long long pos = 0;
        for (int j = 0; j < nrow; ++j) {
            if (j == r) continue;
            const Row o = R.rows[first + j];
            const bool before = row_less(W, R, k, o, w) || (!row_less(W, R, k, w, o) && j < r);      // stable
            if (before) pos += F.row_len[first + j];
        }
        F.row_pos[first + r] = (int32_t)pos;
    }
    if (lane == 0) F.t_len[t] = total;
}

// exclusive prefix of t_len into t_off (one CTA: every thread sums a contiguous stretch, the 1024 partial sums
// are scanned with shuffles, every thread writes its stretch back)
__global__ void __launch_bounds__(1024) km_format_scan_kernel(FormatView F, int n) {
    __shared__ long long warp_tot[32];
    const int per = (n + 1023) / 1024;
    const int lo = (int)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    long long mine = 0;
    for (int i = lo; i < hi; ++i) mine += F.t_len[i];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        long long w = warp_tot[lane], wi = w;
        for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += v; }
        warp_tot[lane] = wi - w;                                  // exclusive prefix of the warp totals
        if (lane == 31) {
            F.t_off[n] = wi;
            *reinterpret_cast<long long*>(F.flags + 2) = wi;
            if (wi > F.text_cap) atomicOr(F.flags, 1u);
        }
    }
    __syncthreads();
    long long at = warp_tot[wid] + incl - mine;
    for (int i = lo; i < hi; ++i) { F.t_off[i] = at; at += F.t_len[i]; }
}

// one WARP per target writes its rows, each at its place
__global__ void __launch_bounds__(128) km_format_write_kernel(WalkView W, ResultView R, FormatView F, int k) {
    const int t = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= W.n_targets) return;
    if (*F.flags) return;                            // the text does not fit / cannot be printed here: the host formats this batch
    const int nrow = R.t_n_rows[t], first = R.t_row_first[t];
    char* block = F.text + F.t_off[t];
    for (int r = 0; r < nrow; ++r) {
        const Row w = R.rows[first + r];
        WarpEmit e;
        e.dst = block + F.row_pos[first + r];
        e.lane = lane;
        StoredNums nums;
        nums.store = F.row_num + (size_t)KM_FMT_ROW_BYTES * (size_t)(first + r);
        format_row(e, nums, F, W, R, k, w);
    }
}

#endif  // KM_DEVICE_BUILD && KM_FORMAT_KERNELS

}  // namespace km
