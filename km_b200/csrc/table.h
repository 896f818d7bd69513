// GPU k-mer count table: the replacement for Jellyfish's QueryMerFile behind
// km/utils/Jellyfish.py (qf[mer] -> count, 0 when absent; Jellyfish.py:53).
//
// Layout in HBM: an array of 32-byte buckets, 32-byte aligned -- exactly one DRAM sector,
// the smallest unit HBM3e delivers -- each holding two (key, count) records:
//     +0  key[0]  u64      +16 count[0] u32     +24 8 bytes unused
//     +8  key[1]  u64      +20 count[1] u32
//     +24 hop u64: bit d (0..62) = "a key whose HOME is this bucket lives in bucket home+1+d";
//                  bit 63 = "... lives further away" (never seen at load 0.5; linear scan)
// A lookup reads ONE sector with one 256-bit load in the common case -- for an ABSENT key too: the
// hop word of the home bucket names every other bucket that holds one of its keys (hopscotch-style),
// so a full home bucket with hop == 0 answers "absent" at once, and otherwise exactly the named
// buckets are read (92 % / 8 % one / two dependent reads at load 0.47; without the hop word a
// full bucket forwarded to its neighbour 26 % of the time and chains of 3-4 dependent reads were
// common -- that, not bandwidth, was what a walk level waited for).  Insertion is linear probing at
// bucket granularity from the home bucket; placing a key away from home sets the home's hop bit.
// Bucket = mulhi64(mix64(key), n_buckets): no power-of-two constraint, so a 2e9-key table can sit
// at any load factor the HBM budget allows.  Empty key = ~0 (never a valid canonical k-mer for k <= 31).
#pragma once
#include "kmer.h"

namespace km {

#define KM_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define KM_GOLDEN_T 0x9E3779B97F4A7C15ull
#define KM_BUCKET_SLOTS 2

struct alignas(32) Bucket {
    uint64_t key[KM_BUCKET_SLOTS];
    uint32_t count[KM_BUCKET_SLOTS];
    uint32_t pad[2];          // the 64-bit hop word (see above), low half first
};
// The hop word, refined: bits 0..46 hop distances 1..47, bit 47 "further away", and -- new -- bits 48..55 / 56..63 the
// NEIGHBOUR MASK of the key in slot 0 / slot 1: bit c (0..3) = "the k-mer key[1:] + c is in the table", bit 4 + c =
// "c + key[:-1] is in the table" (letters A C G T, on the strand the canonical key is written in).  MutationFinder asks
// for a k-mer's count and at once for its four successors (Jellyfish.get_child, Jellyfish.py:61-66), and three of those
// four are absent on every reference k-mer outside a variant: with the mask riding along with the k-mer's own record,
// "absent" is known without touching memory -- the reference-probe kernel reads ~1 line per reference k-mer instead of
// ~4 (every random 32-byte read costs a 128-byte DRAM line; that traffic was 4.6x the algorithmic bytes).  The masks are
// written by km_table_link_kernel after the table's content has changed (TableView::linked says they are current);
// a set bit is always verified by the lookup it allows, a clear bit IS the answer "count 0".
#define KM_HOP_DIST 47
#define KM_HOP_FAR (1ull << 47)
#define KM_HOP_MASK ((1ull << 48) - 1ull)
#define KM_LINK_SHIFT(slot) (48 + 8 * (slot))

#define KM_MAX_SHARDS 8

// One process holds ONE shard (n_shards == 1: the whole table).  In cohort mode the table is
// hash-sharded over the GPUs of a box: shard[r] is rank r's bucket array mapped into this process
// through CUDA IPC, so a probe of a remote key is an ordinary 32-byte load that travels over
// NVLink/NVSwitch from inside the same kernels -- no exchange step, no extra launch.
struct TableView {
    Bucket* buckets;          // this process's shard
    uint64_t n_buckets;       // buckets per shard
    uint64_t kmask;
    int k;
    int canonical;
    int n_shards;
    int my_shard;
    const Bucket* shard[KM_MAX_SHARDS];   // shard[my_shard] == buckets; others null until peers are attached
    int lines;                // 0: sector buckets (above); 1: family lines (below), n_buckets counts 128-byte lines
    int linked;               // 1: the neighbour masks in the hop words are current (km_table_link); sector layout, one shard
    int route;                // cohort: 1 = an insert goes to the key's OWNER shard, wherever it is (peer atomics over NVLink:
                              // every rank gives its own part of the stream); 0 = only owned keys are kept (every rank
                              // streams everything)
};

KM_HD uint64_t key_hash(uint64_t key) { return mix64(key + KM_GOLDEN_T); }
// owner of a key: the top bits of its hash; the bucket inside the shard comes from a remix, so the
// two are independent.  With one shard the mapping is the plain multiply-shift of the hash.
KM_HD int shard_of_hash(uint64_t h, int n_shards) { return n_shards > 1 ? (int)mulhi64(h, (uint64_t)n_shards) : 0; }
KM_HD uint64_t bucket_of_hash(uint64_t h, int n_shards, uint64_t n_buckets) {
    return mulhi64(n_shards > 1 ? h * 0xD6E8FEB86659FD93ull : h, n_buckets);
}
// where a key lives: base of its shard + bucket index inside it
KM_HD const Bucket* locate(const TableView& t, uint64_t key, uint64_t* b) {
    const uint64_t h = key_hash(key);
    *b = bucket_of_hash(h, t.n_shards, t.n_buckets);
    return t.n_shards > 1 ? t.shard[shard_of_hash(h, t.n_shards)] : t.buckets;
}

#if KM_DEVICE_BUILD
// One 32-byte sector with ONE 256-bit load (sm_100 has ld.global.v4.u64).  Measured on B200
// (tools/probes/peer_gather.cu): the memory system serves ~36 G random REQUESTS/s whatever their
// size, so a bucket fetched as two 16-byte loads -- the L1-bypassing kind does not merge them -- tops
// out at 18 G buckets/s, one 32-byte load at 36 G/s (and 6.6 vs 3.3 G/s from a peer GPU over NVLink).
KM_HD void load_bucket(const Bucket* b, uint64_t& k0, uint64_t& k1, uint32_t& c0, uint32_t& c1, uint64_t& hop) {
    uint64_t cc;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(k1), "=l"(cc), "=l"(hop) : "l"(b));
    c0 = (uint32_t)cc; c1 = (uint32_t)(cc >> 32);
}
#else
KM_HD void load_bucket(const Bucket* b, uint64_t& k0, uint64_t& k1, uint32_t& c0, uint32_t& c1, uint64_t& hop) {
    k0 = b->key[0]; k1 = b->key[1]; c0 = b->count[0]; c1 = b->count[1];
    hop = (uint64_t)b->pad[0] | ((uint64_t)b->pad[1] << 32);
}
#endif
KM_HD void load_bucket(const Bucket* b, uint64_t& k0, uint64_t& k1, uint32_t& c0, uint32_t& c1) {
    uint64_t hop;
    load_bucket(b, k0, k1, c0, c1, hop);
}

// The rest of a lookup once the HOME bucket `b` of `key` has been read (k0..hop): the key is in the home
// bucket, in one of the buckets its hop word names, or absent.
KM_HD uint32_t finish_lookup(const Bucket* base, uint64_t n_buckets, uint64_t b, uint64_t key, uint64_t k0, uint64_t k1,
                             uint32_t c0, uint32_t c1, uint64_t hop) {
    if (k0 == key) return c0;
    if (k1 == key) return c1;
    hop &= KM_HOP_MASK;                          // (the top 16 bits are the slots' neighbour masks)
    if (hop & KM_HOP_FAR) {                      // a key of this home sits 48+ buckets away: linear scan from there
        uint64_t bb = (b + KM_HOP_DIST + 1) % n_buckets;
        for (uint64_t tries = 0; tries < n_buckets; ++tries) {
            uint64_t f0, f1; uint32_t d0, d1;
            load_bucket(base + bb, f0, f1, d0, d1);
            if (f0 == key) return d0;
            if (f1 == key) return d1;
            if (f0 == KM_EMPTY_KEY || f1 == KM_EMPTY_KEY) break;
            if (++bb == n_buckets) bb = 0;
        }
        hop &= ~KM_HOP_FAR;
    }
    while (hop) {
        const int d = ffs64(hop) - 1;
        hop &= hop - 1;
        uint64_t bb = b + 1 + (uint64_t)d;
        if (bb >= n_buckets) bb -= n_buckets;
        uint64_t f0, f1; uint32_t d0, d1;
        load_bucket(base + bb, f0, f1, d0, d1);
        if (f0 == key) return d0;
        if (f1 == key) return d1;
    }
    return 0;
}

// ---- family lines (TableView::lines == 1) ---------------------------------------------------------------
// Measured on B200: a random 32-byte read costs a whole 128-byte DRAM line (ncu: 4 L2 sectors, 129 B of DRAM
// traffic per load, whatever cudaLimitMaxL2FetchGranularity says), and that line traffic -- not the request
// count -- is what caps random probes.  So the table is laid out such that the k-mers the walk asks for
// TOGETHER share a line: a line = 8 slots of (u64 key, u32 count, 4 spare bytes), addressed by a canonical
// (k-1)-mer M, and holds the k-mers that contain M as their first or last k-1 bases -- the four successors
// M+c and the four predecessors c+M.  Every k-mer is therefore stored twice (under its prefix and under
// its suffix (k-1)-mer).  A reference k-mer X and its four successors X[1:]+c all live in the line of
// canonical(X[1:]): get_child plus the k-mer's own count is ONE line instead of five sectors.  A line that
// is full forwards to the next one (linear probing at line granularity).  Inside a line a k-mer prefers the
// 32-byte sector chosen by its own hash, so an isolated query usually reads a single sector.
#define KM_LINE_SLOTS 8
// the two copies of a k-mer are told apart by the top bit of the stored key (k <= 31 uses 62 bits): the
// copy under the suffix family carries it.  Without it an insertion whose two families share a line would
// find its own first copy and count twice.
#define KM_COPY_BIT 0x8000000000000000ull
#define KM_KEY_OF(stored) ((stored) & ~KM_COPY_BIT)
struct alignas(16) LineSlot { uint64_t key; uint32_t count; uint32_t pad; };
struct alignas(128) Line { LineSlot s[KM_LINE_SLOTS]; };

KM_HD uint64_t sub_canon(const TableView& t, uint64_t v) {         // v: a (k-1)-mer
    return t.canonical ? canonical(v, t.k - 1) : v;
}
// the family in which a forward k-mer appears as a SUCCESSOR (shares its first k-1 bases with its siblings)
KM_HD uint64_t family_of_prefix(const TableView& t, uint64_t fwd) { return sub_canon(t, (fwd & t.kmask) >> 2); }
// the family in which a forward k-mer appears as the PARENT of its successors / as a predecessor
KM_HD uint64_t family_of_suffix(const TableView& t, uint64_t fwd) { return sub_canon(t, fwd & (t.kmask >> 2)); }
KM_HD int preferred_sector(uint64_t key) { return (int)((key * 0x9E3779B97F4A7C15ull) >> 62); }

KM_HD const Line* locate_line(const TableView& t, uint64_t fam, uint64_t* idx) {
    const uint64_t h = key_hash(fam ^ 0x5851F42D4C957F2Dull);
    *idx = bucket_of_hash(h, t.n_shards, t.n_buckets);
    const Bucket* base = t.n_shards > 1 ? t.shard[shard_of_hash(h, t.n_shards)] : t.buckets;
    return reinterpret_cast<const Line*>(base);
}
KM_HD int family_owner(const TableView& t, uint64_t fam) { return shard_of_hash(key_hash(fam ^ 0x5851F42D4C957F2Dull), t.n_shards); }

#if KM_DEVICE_BUILD
KM_HD void load_sector(const LineSlot* p, uint64_t& k0, uint32_t& c0, uint64_t& k1, uint32_t& c1) {
    uint64_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(a), "=l"(k1), "=l"(b) : "l"(p));
    c0 = (uint32_t)a; c1 = (uint32_t)b;
}
#else
KM_HD void load_sector(const LineSlot* p, uint64_t& k0, uint32_t& c0, uint64_t& k1, uint32_t& c1) {
    k0 = p[0].key; c0 = p[0].count; k1 = p[1].key; c1 = p[1].count;
}
#endif

// counts of up to N canonical keys that all belong to one family, starting at line `idx` of `base`: the whole
// line is fetched (four 32-byte loads in flight), then the rare forwarding to the next line
template <int N>
KM_HD void line_find_from(const TableView& t, const Line* base, uint64_t idx, const uint64_t (&key)[N], uint32_t pending,
                          uint32_t (&out)[N]) {
    while (pending) {
        const LineSlot* p = base[idx].s;
        uint64_t k[KM_LINE_SLOTS];
        uint32_t c[KM_LINE_SLOTS];
#pragma unroll
        for (int h = 0; h < 4; ++h) load_sector(p + 2 * h, k[2 * h], c[2 * h], k[2 * h + 1], c[2 * h + 1]);
        bool has_empty = false;
#pragma unroll
        for (int s = 0; s < KM_LINE_SLOTS; ++s) {
            has_empty |= k[s] == KM_EMPTY_KEY;
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (KM_KEY_OF(k[s]) == key[i] && (pending & (1u << i))) { out[i] = c[s]; pending &= ~(1u << i); }
        }
        if (has_empty) return;
        if (++idx == t.n_buckets) idx = 0;
    }
}
template <int N>
KM_HD void line_find(const TableView& t, uint64_t fam, const uint64_t (&key)[N], uint32_t mask, uint32_t (&out)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = 0;
    uint64_t idx;
    const Line* base = locate_line(t, fam, &idx);
    line_find_from<N>(t, base, idx, key, mask, out);
}

#if KM_DEVICE_BUILD
// The same for a whole warp, every lane with its own family: lanes fetch each other's lines FOUR LANES PER
// LINE (lane j of a quad loads sector j), so a line is one coalesced 128-byte request instead of four
// 32-byte ones, and the sectors are handed back to the owning lane by shuffle.  Four passes cover the 32
// lanes; the loads of all passes are issued before any is consumed.  All 32 lanes must call.
template <int N>
__device__ __forceinline__ void warp_line_find(const TableView& t, uint64_t fam, const uint64_t (&key)[N], uint32_t mask,
                                               uint32_t (&out)[N]) {
    const int lane = threadIdx.x & 31, quad = lane >> 2, sec = lane & 3;
    uint64_t idx;
    const Line* base = locate_line(t, fam, &idx);
    const unsigned long long mine = (unsigned long long)(base[idx].s);
    uint64_t k0[4], k1[4], cc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const LineSlot* line = (const LineSlot*)__shfl_sync(0xFFFFFFFFu, mine, 8 * p + quad);
        uint32_t c0, c1;
        load_sector(line + 2 * sec, k0[p], c0, k1[p], c1);
        cc[p] = (uint64_t)c0 | ((uint64_t)c1 << 32);
    }
    uint64_t K[KM_LINE_SLOTS];
    uint32_t C[KM_LINE_SLOTS];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int src = 4 * (lane & 7) + s;
            const uint64_t a = __shfl_sync(0xFFFFFFFFu, (unsigned long long)k0[p], src);
            const uint64_t b = __shfl_sync(0xFFFFFFFFu, (unsigned long long)k1[p], src);
            const uint64_t c = __shfl_sync(0xFFFFFFFFu, (unsigned long long)cc[p], src);
            if ((lane >> 3) == p) { K[2 * s] = a; K[2 * s + 1] = b; C[2 * s] = (uint32_t)c; C[2 * s + 1] = (uint32_t)(c >> 32); }
        }
    }
    uint32_t pending = mask;
    bool has_empty = false;
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = 0;
#pragma unroll
    for (int s = 0; s < KM_LINE_SLOTS; ++s) {
        has_empty |= K[s] == KM_EMPTY_KEY;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (KM_KEY_OF(K[s]) == key[i] && (pending & (1u << i))) { out[i] = C[s]; pending &= ~(1u << i); }
    }
    // a full line that lacks a wanted key forwards to the next line: rare, finished by the lane on its own
    if (pending && !has_empty) line_find_from<N>(t, base, idx + 1 == t.n_buckets ? 0 : idx + 1, key, pending, out);
}
#endif

// one isolated k-mer: its preferred sector first, the rest of the line only if that sector is full without it
KM_HD uint32_t line_lookup_one(const TableView& t, uint64_t fam, uint64_t key) {
    uint64_t idx;
    const Line* base = locate_line(t, fam, &idx);
    const int pref = preferred_sector(key);
    for (;;) {
        const LineSlot* p = base[idx].s;
        uint64_t k0, k1; uint32_t c0, c1;
        load_sector(p + 2 * pref, k0, c0, k1, c1);
        if (k0 != KM_EMPTY_KEY && KM_KEY_OF(k0) == key) return c0;
        if (k1 != KM_EMPTY_KEY && KM_KEY_OF(k1) == key) return c1;
        if (k0 == KM_EMPTY_KEY || k1 == KM_EMPTY_KEY) return 0;     // insertion fills the preferred sector first
        bool has_empty = false;
#pragma unroll
        for (int h = 1; h < 4; ++h) {
            load_sector(p + 2 * ((pref + h) & 3), k0, c0, k1, c1);
            if (k0 != KM_EMPTY_KEY && KM_KEY_OF(k0) == key) return c0;
            if (k1 != KM_EMPTY_KEY && KM_KEY_OF(k1) == key) return c1;
            has_empty |= k0 == KM_EMPTY_KEY || k1 == KM_EMPTY_KEY;
        }
        if (has_empty) return 0;
        if (++idx == t.n_buckets) idx = 0;
    }
}

// one copy of the k-mer into the line of `fam` (`key` carries the copy bit): 1 newly inserted, 0 existed, -1 table full
KM_HD int line_insert(const TableView& t, uint64_t fam, uint64_t key, uint32_t count, int mode) {
    uint64_t idx;
    Line* base = const_cast<Line*>(locate_line(t, fam, &idx));
    const int pref = preferred_sector(KM_KEY_OF(key));
    for (uint64_t tries = 0; tries < t.n_buckets; ++tries) {
        LineSlot* p = base[idx].s;
        for (int j = 0; j < KM_LINE_SLOTS; ++j) {
            LineSlot* sl = p + ((2 * pref + j) & (KM_LINE_SLOTS - 1));
            uint64_t cur = load_cg64(&sl->key);
            if (cur == KM_EMPTY_KEY) {
                cur = atomic_cas64(&sl->key, KM_EMPTY_KEY, key);
                if (cur == KM_EMPTY_KEY) {
                    if (mode == 2) atomic_add32(&sl->count, count);
                    else sl->count = count;
                    return 1;
                }
            }
            if (cur == key) {
                if (mode == 2) atomic_add32(&sl->count, count);
                else if (mode == 1) sl->count = count;
                return 0;
            }
        }
        if (++idx == t.n_buckets) idx = 0;
    }
    return -1;
}

// canonical key -> count (0 when absent).  Read-only path: valid only while no kernel is
// inserting into the table.
KM_HD uint32_t table_lookup_key(const TableView& t, uint64_t key) {
    uint64_t b;
    const Bucket* base = locate(t, key, &b);
    uint64_t k0, k1, hop; uint32_t c0, c1;
    load_bucket(base + b, k0, k1, c0, c1, hop);
    return finish_lookup(base, t.n_buckets, b, key, k0, k1, c0, c1, hop);
}

// The same, also telling WHERE the key was found: *at = its bucket, *slot = its slot; false when absent.
KM_HD bool table_find_key(const TableView& t, uint64_t key, const Bucket** at, int* slot, uint32_t* count, uint64_t* word) {
    uint64_t b;
    const Bucket* base = locate(t, key, &b);
    uint64_t k0, k1, hop; uint32_t c0, c1;
    load_bucket(base + b, k0, k1, c0, c1, hop);
    if (k0 == key) { *at = base + b; *slot = 0; *count = c0; *word = hop; return true; }
    if (k1 == key) { *at = base + b; *slot = 1; *count = c1; *word = hop; return true; }
    uint64_t h = hop & KM_HOP_MASK;
    if (h & KM_HOP_FAR) {
        uint64_t bb = (b + KM_HOP_DIST + 1) % t.n_buckets;
        for (uint64_t tries = 0; tries < t.n_buckets; ++tries) {
            uint64_t f0, f1, w; uint32_t d0, d1;
            load_bucket(base + bb, f0, f1, d0, d1, w);
            if (f0 == key) { *at = base + bb; *slot = 0; *count = d0; *word = w; return true; }
            if (f1 == key) { *at = base + bb; *slot = 1; *count = d1; *word = w; return true; }
            if (f0 == KM_EMPTY_KEY || f1 == KM_EMPTY_KEY) break;
            if (++bb == t.n_buckets) bb = 0;
        }
        h &= ~KM_HOP_FAR;
    }
    while (h) {
        const int d = ffs64(h) - 1;
        h &= h - 1;
        uint64_t bb = b + 1 + (uint64_t)d;
        if (bb >= t.n_buckets) bb -= t.n_buckets;
        uint64_t f0, f1, w; uint32_t d0, d1;
        load_bucket(base + bb, f0, f1, d0, d1, w);
        if (f0 == key) { *at = base + bb; *slot = 0; *count = d0; *word = w; return true; }
        if (f1 == key) { *at = base + bb; *slot = 1; *count = d1; *word = w; return true; }
    }
    return false;
}

KM_HD uint32_t rev_nibble(uint32_t x) { return ((x & 1u) << 3) | ((x & 2u) << 1) | ((x & 4u) >> 1) | ((x & 8u) >> 3); }

// Jellyfish.query of a forward-strand k-mer PLUS which of its four successors fwd[1:] + c exist (bit c of *succ), from the
// neighbour mask stored with the key (valid only while TableView::linked).  Returns false when the k-mer itself is
// absent: then nothing is known about its successors (*succ = 15: ask for all of them).
KM_HD bool table_query_links(const TableView& t, uint64_t fwd, uint32_t* count, uint32_t* succ) {
    const uint64_t v = fwd & t.kmask;
    const uint64_t rc = revcomp(v, t.k);
    const bool flip = t.canonical && rc < v;
    const Bucket* at; int slot; uint64_t word;
    if (!table_find_key(t, flip ? rc : v, &at, &slot, count, &word)) { *count = 0; *succ = 15u; return false; }
    const uint32_t m = (uint32_t)(word >> KM_LINK_SHIFT(slot)) & 0xFFu;
    // on the other strand the successors of the query are the predecessors of the stored key, letters complemented
    *succ = flip ? rev_nibble(m >> 4) : (m & 15u);
    return true;
}

// forward-strand packed k-mer -> count: Jellyfish.query (km/utils/Jellyfish.py:47-53)
KM_HD uint32_t table_query(const TableView& t, uint64_t fwd) {
    uint64_t v = fwd & t.kmask;
    const uint64_t key = t.canonical ? canonical(v, t.k) : v;
    if (t.lines) return line_lookup_one(t, family_of_prefix(t, v), key);
    return table_lookup_key(t, key);
}

// N independent first probes in flight per thread (the walk issues the 4 successor
// lookups of Jellyfish.get_child this way), then the rare forwarding loop per query.
template <int N>
KM_HD void table_query_multi(const TableView& T, const uint64_t (&fwd)[N], uint32_t (&out)[N]) {
    uint64_t key[N], b[N], k0[N], k1[N], hop[N];
    uint32_t c0[N], c1[N];
    const Bucket* base[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint64_t v = fwd[i] & T.kmask;
        key[i] = T.canonical ? canonical(v, T.k) : v;
        base[i] = locate(T, key[i], &b[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) load_bucket(base[i] + b[i], k0[i], k1[i], c0[i], c1[i], hop[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = finish_lookup(base[i], T.n_buckets, b[i], key[i], k0[i], k1[i], c0[i], c1[i], hop[i]);
}

// N independent lookups in flight, only where `mask` has the bit set (others return 0).
template <int N>
KM_HD void table_query_masked(const TableView& T, const uint64_t (&fwd)[N], uint32_t mask, uint32_t (&out)[N]) {
    uint64_t key[N], b[N], k0[N], k1[N], hop[N];
    uint32_t c0[N], c1[N];
    const Bucket* base[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint64_t v = fwd[i] & T.kmask;
        key[i] = T.canonical ? canonical(v, T.k) : v;
        base[i] = locate(T, key[i], &b[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (mask & (1u << i)) load_bucket(base[i] + b[i], k0[i], k1[i], c0[i], c1[i], hop[i]);
#pragma unroll
    for (int i = 0; i < N; ++i)
        out[i] = (mask & (1u << i)) ? finish_lookup(base[i], T.n_buckets, b[i], key[i], k0[i], k1[i], c0[i], c1[i], hop[i]) : 0u;
}

// Lookups that belong together (the successors of one parent, plus the parent itself): `fam` is their
// family (family_of_prefix of a successor == family_of_suffix of the parent).  With sector buckets this
// is N independent probes; with family lines it is one line.
template <int N>
KM_HD void table_query_family(const TableView& T, uint64_t fam, const uint64_t (&fwd)[N], uint32_t mask, uint32_t (&out)[N]) {
    if (T.lines) {
        uint64_t key[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { const uint64_t v = fwd[i] & T.kmask; key[i] = T.canonical ? canonical(v, T.k) : v; }
        line_find<N>(T, fam, key, mask, out);
    } else {
        table_query_masked<N>(T, fwd, mask, out);
    }
}

#if KM_DEVICE_BUILD
// The four lanes of a quad each ask for one successor of the same parent (family `fam`): lane j loads sector j
// of the family's line -- one coalesced 128-byte request for the quad -- and every lane then searches the
// eight slots (handed round by shuffle) for its own key.  All 32 lanes must call; quads with `live` false
// skip the load.
__device__ __forceinline__ uint32_t quad_line_query(const TableView& t, uint64_t fam, uint64_t key, bool live) {
    const int lane = threadIdx.x & 31, qbase = lane & 28, sec = lane & 3;
    uint64_t idx = 0;
    const Line* base = nullptr;
    uint64_t k0 = KM_EMPTY_KEY, k1 = KM_EMPTY_KEY, cc = 0;
    if (live) {
        base = locate_line(t, fam, &idx);
        uint32_t c0, c1;
        load_sector(base[idx].s + 2 * sec, k0, c0, k1, c1);
        cc = (uint64_t)c0 | ((uint64_t)c1 << 32);
    }
    uint32_t out = 0;
    bool found = false, has_empty = false;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint64_t a = __shfl_sync(0xFFFFFFFFu, (unsigned long long)k0, qbase + s);
        const uint64_t b = __shfl_sync(0xFFFFFFFFu, (unsigned long long)k1, qbase + s);
        const uint64_t c = __shfl_sync(0xFFFFFFFFu, (unsigned long long)cc, qbase + s);
        has_empty |= a == KM_EMPTY_KEY || b == KM_EMPTY_KEY;
        if (KM_KEY_OF(a) == key) { out = (uint32_t)c; found = true; }
        if (KM_KEY_OF(b) == key) { out = (uint32_t)(c >> 32); found = true; }
    }
    if (live && !found && !has_empty) {          // full line without the key: the next line, on this lane's own
        const uint64_t keys[1] = {key};
        uint32_t r[1] = {0};
        line_find_from<1>(t, base, idx + 1 == t.n_buckets ? 0 : idx + 1, keys, 1u, r);
        out = r[0];
    }
    return out;
}
#endif

// table_query_family for a converged warp (all 32 lanes call; lanes with nothing to ask pass mask 0)
template <int N>
KM_HD void table_query_family_warp(const TableView& T, uint64_t fam, const uint64_t (&fwd)[N], uint32_t mask, uint32_t (&out)[N]) {
#if KM_DEVICE_BUILD
    if (T.lines) {
        uint64_t key[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { const uint64_t v = fwd[i] & T.kmask; key[i] = T.canonical ? canonical(v, T.k) : v; }
        warp_line_find<N>(T, fam, key, mask, out);
        return;
    }
#endif
    table_query_family<N>(T, fam, fwd, mask, out);
}

enum InsertMode { KM_INSERT_KEEP = 0, KM_INSERT_OVERWRITE = 1, KM_INSERT_ADD = 2 };

// One key into the bucket array `base` (this process's shard, or a peer's mapped over NVLink when `sys`): linear
// probing at bucket granularity from `home`; placing a key away from home sets the home's hop bit.  Returns 1 if the
// key was newly inserted, 0 if it already existed, -1 if the shard is full.  With `sys` every atomic has system
// scope -- it is carried out at the home GPU's L2, which is what makes inserts from several GPUs into one shard
// safe; plain loads of a peer's slot may be stale only in the harmless direction (a slot read as empty is then
// claimed with a CAS, which returns what is really there; a slot once filled never changes its key).
KM_HD int bucket_insert(Bucket* base, uint64_t n_buckets, uint64_t home, uint64_t key, uint32_t count, int mode, bool sys) {
    uint64_t b = home;
    for (uint64_t tries = 0; tries < n_buckets; ++tries) {
        Bucket* bk = base + b;
        for (int s = 0; s < KM_BUCKET_SLOTS; ++s) {
            uint64_t cur = sys ? load64_sys(&bk->key[s]) : load_cg64(&bk->key[s]);
            int placed = -1;
            if (cur == KM_EMPTY_KEY) {
                cur = sys ? atomic_cas64_sys(&bk->key[s], KM_EMPTY_KEY, key) : atomic_cas64(&bk->key[s], KM_EMPTY_KEY, key);
                if (cur == KM_EMPTY_KEY) {
                    if (mode == 2) { if (sys) atomic_add32_sys(&bk->count[s], count); else atomic_add32(&bk->count[s], count); }
                    else if (sys) store32_sys(&bk->count[s], count);
                    else bk->count[s] = count;
                    placed = 1;
                }
            }
            if (placed < 0 && cur == key) {
                if (mode == 2) { if (sys) atomic_add32_sys(&bk->count[s], count); else atomic_add32(&bk->count[s], count); }
                else if (mode == 1) { if (sys) store32_sys(&bk->count[s], count); else bk->count[s] = count; }
                placed = 0;
            }
            if (placed >= 0) {
                // away from home: the home bucket's hop word must name this bucket (every inserter of the key sets
                // the same bit, so whoever finishes last leaves it set)
                if (tries) {
                    uint64_t* hop = reinterpret_cast<uint64_t*>(base[home].pad);
                    const uint64_t bit = tries <= KM_HOP_DIST ? 1ull << (tries - 1) : KM_HOP_FAR;
                    if (sys) atomic_or64_sys(hop, bit); else atomic_or64(hop, bit);
                }
                return placed;
            }
        }
        if (++b == n_buckets) b = 0;
    }
    return -1;
}

// Counting (mode add) of a key that is usually there already: ONE load of the home bucket's two keys finds it, one
// reduction adds to it; anything else (absent, displaced) takes bucket_insert.
KM_HD int bucket_count(Bucket* base, uint64_t n_buckets, uint64_t home, uint64_t key, uint32_t count, bool sys) {
#if KM_DEVICE_BUILD
    uint64_t k0, k1;               // both keys of the home bucket with one 128-bit load (at L2: the slot may have just been claimed)
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "l"(base + home) : "memory");
    if (k0 == key || k1 == key) {
        uint32_t* c = &base[home].count[k0 == key ? 0 : 1];
        if (sys) atomic_add32_sys(c, count); else atomic_add32(c, count);
        return 0;
    }
#endif
    return bucket_insert(base, n_buckets, home, key, count, 2, sys);
}

// Returns 1 if the key was newly inserted, 0 if it already existed (or was left to its owner), -1 if the shard
// is full.  Without routing only the owner inserts a key; with routing (TableView::route, peers attached) the key
// goes to its owner's shard from wherever it was seen.
KM_HD int table_insert(const TableView& t, uint64_t key, uint32_t count, int mode) {
    if (t.lines) {
        // two copies: under the k-mer's prefix (k-1)-mer and under its suffix (k-1)-mer; each goes to the
        // shard that owns its family; "newly inserted" is reported for the prefix copy only
        const uint64_t f1 = sub_canon(t, key >> 2), f2 = sub_canon(t, key & (t.kmask >> 2));
        int r1 = 0, r2 = 0;
        if (family_owner(t, f1) == t.my_shard) r1 = line_insert(t, f1, key, count, mode);
        if (f2 != f1 && family_owner(t, f2) == t.my_shard) r2 = line_insert(t, f2, key | KM_COPY_BIT, count, mode);
        return (r1 < 0 || r2 < 0) ? -1 : r1;
    }
    const uint64_t h = key_hash(key);
    const int owner = shard_of_hash(h, t.n_shards);
    const uint64_t home = bucket_of_hash(h, t.n_shards, t.n_buckets);
    if (owner != t.my_shard) {
        if (!t.route) return 0;
        Bucket* base = const_cast<Bucket*>(t.shard[owner]);
        return mode == 2 ? bucket_count(base, t.n_buckets, home, key, count, true) : bucket_insert(base, t.n_buckets, home, key, count, mode, true);
    }
    const bool sys = t.route != 0;                 // peers may be inserting into this shard at the same time
    return mode == 2 ? bucket_count(t.buckets, t.n_buckets, home, key, count, sys) : bucket_insert(t.buckets, t.n_buckets, home, key, count, mode, sys);
}

}  // namespace km
