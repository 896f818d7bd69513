// GPU k-mer count table: the replacement for Jellyfish's QueryMerFile behind
// km/utils/Jellyfish.py (qf[mer] -> count, 0 when absent; Jellyfish.py:53).
//
// Layout in HBM: an array of 32-byte buckets, 32-byte aligned -- exactly one DRAM sector,
// the smallest unit HBM3e delivers -- each holding two (key, count) records:
//     +0  key[0]  u64      +16 count[0] u32     +24 8 bytes unused
//     +8  key[1]  u64      +20 count[1] u32
// A lookup reads ONE sector in the common case (both 16-byte halves are requested by the
// same thread, so the LSU coalesces them into a single 32-byte sector request); only a
// full bucket that does not hold the key forwards to the next bucket (linear probing at
// bucket granularity).  Bucket = mulhi64(mix64(key), n_buckets): no power-of-two
// constraint, so a 2e9-key table can sit at any load factor the HBM budget allows.
// Empty key = ~0 (never a valid canonical k-mer for k <= 31).
#pragma once
#include "kmer.h"

namespace km {

#define KM_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define KM_GOLDEN_T 0x9E3779B97F4A7C15ull
#define KM_BUCKET_SLOTS 2

struct alignas(32) Bucket {
    uint64_t key[KM_BUCKET_SLOTS];
    uint32_t count[KM_BUCKET_SLOTS];
    uint32_t pad[2];
};

struct TableView {
    Bucket* buckets;
    uint64_t n_buckets;
    uint64_t kmask;
    int k;
    int canonical;
};

KM_HD uint64_t bucket_of(const TableView& t, uint64_t key) { return mulhi64(mix64(key + KM_GOLDEN_T), t.n_buckets); }

#if KM_DEVICE_BUILD
// one 32-byte sector as two 128-bit read-only loads that bypass L1 allocation
KM_HD void load_bucket(const Bucket* b, uint64_t& k0, uint64_t& k1, uint32_t& c0, uint32_t& c1) {
    uint32_t pad0, pad1;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "l"(b));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(c0), "=r"(c1), "=r"(pad0), "=r"(pad1) : "l"(reinterpret_cast<const char*>(b) + 16));
}
#else
KM_HD void load_bucket(const Bucket* b, uint64_t& k0, uint64_t& k1, uint32_t& c0, uint32_t& c1) {
    k0 = b->key[0]; k1 = b->key[1]; c0 = b->count[0]; c1 = b->count[1];
}
#endif

// canonical key -> count (0 when absent).  Read-only path: valid only while no kernel is
// inserting into the table.
KM_HD uint32_t table_lookup_key(const TableView& t, uint64_t key) {
    uint64_t b = bucket_of(t, key);
    for (;;) {
        uint64_t k0, k1; uint32_t c0, c1;
        load_bucket(t.buckets + b, k0, k1, c0, c1);
        if (k0 == key) return c0;
        if (k1 == key) return c1;
        if (k0 == KM_EMPTY_KEY || k1 == KM_EMPTY_KEY) return 0;
        if (++b == t.n_buckets) b = 0;
    }
}

// forward-strand packed k-mer -> count: Jellyfish.query (km/utils/Jellyfish.py:47-53)
KM_HD uint32_t table_query(const TableView& t, uint64_t fwd) {
    uint64_t v = fwd & t.kmask;
    return table_lookup_key(t, t.canonical ? canonical(v, t.k) : v);
}

// N independent first probes in flight per thread (the walk issues the 4 successor
// lookups of Jellyfish.get_child this way), then the rare forwarding loop per query.
template <int N>
KM_HD void table_query_multi(const TableView& T, const uint64_t (&fwd)[N], uint32_t (&out)[N]) {
    uint64_t key[N], b[N], k0[N], k1[N];
    uint32_t c0[N], c1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint64_t v = fwd[i] & T.kmask;
        key[i] = T.canonical ? canonical(v, T.k) : v;
        b[i] = bucket_of(T, key[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) load_bucket(T.buckets + b[i], k0[i], k1[i], c0[i], c1[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        uint32_t r = 0;
        for (;;) {
            if (k0[i] == key[i]) { r = c0[i]; break; }
            if (k1[i] == key[i]) { r = c1[i]; break; }
            if (k0[i] == KM_EMPTY_KEY || k1[i] == KM_EMPTY_KEY) break;
            if (++b[i] == T.n_buckets) b[i] = 0;      // rare: full bucket without the key
            load_bucket(T.buckets + b[i], k0[i], k1[i], c0[i], c1[i]);
        }
        out[i] = r;
    }
}

enum InsertMode { KM_INSERT_KEEP = 0, KM_INSERT_OVERWRITE = 1, KM_INSERT_ADD = 2 };

// Returns 1 if the key was newly inserted, 0 if it already existed, -1 if the table is full.
KM_HD int table_insert(const TableView& t, uint64_t key, uint32_t count, int mode) {
    uint64_t b = bucket_of(t, key);
    for (uint64_t tries = 0; tries < t.n_buckets; ++tries) {
        Bucket* bk = t.buckets + b;
        for (int s = 0; s < KM_BUCKET_SLOTS; ++s) {
            uint64_t cur = load_cg64(&bk->key[s]);
            if (cur == KM_EMPTY_KEY) {
                cur = atomic_cas64(&bk->key[s], KM_EMPTY_KEY, key);
                if (cur == KM_EMPTY_KEY) {
                    if (mode == KM_INSERT_ADD) atomic_add32(&bk->count[s], count);
                    else bk->count[s] = count;
                    return 1;
                }
            }
            if (cur == key) {
                if (mode == KM_INSERT_ADD) atomic_add32(&bk->count[s], count);
                else if (mode == KM_INSERT_OVERWRITE) bk->count[s] = count;
                return 0;
            }
        }
        if (++b == t.n_buckets) b = 0;
    }
    return -1;
}

}  // namespace km
