// GPU k-mer count table: the replacement for Jellyfish's QueryMerFile behind
// km/utils/Jellyfish.py (qf[mer] -> count, 0 when absent; Jellyfish.py:53).
//
// Layout in HBM: an array of 32-byte buckets, 32-byte aligned -- exactly one DRAM sector,
// the smallest unit HBM3e delivers -- each holding two (key, count) records:
//     +0  key[0]  u64      +16 count[0] u32     +24 8 bytes unused
//     +8  key[1]  u64      +20 count[1] u32
// A lookup reads ONE sector with one 256-bit load in the common case; only a
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
KM_HD void load_bucket(const Bucket* b, uint64_t& k0, uint64_t& k1, uint32_t& c0, uint32_t& c1) {
    uint64_t cc, pad;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(k1), "=l"(cc), "=l"(pad) : "l"(b));
    c0 = (uint32_t)cc; c1 = (uint32_t)(cc >> 32);
}
#else
KM_HD void load_bucket(const Bucket* b, uint64_t& k0, uint64_t& k1, uint32_t& c0, uint32_t& c1) {
    k0 = b->key[0]; k1 = b->key[1]; c0 = b->count[0]; c1 = b->count[1];
}
#endif

// canonical key -> count (0 when absent).  Read-only path: valid only while no kernel is
// inserting into the table.
KM_HD uint32_t table_lookup_key(const TableView& t, uint64_t key) {
    uint64_t b;
    const Bucket* base = locate(t, key, &b);
    for (;;) {
        uint64_t k0, k1; uint32_t c0, c1;
        load_bucket(base + b, k0, k1, c0, c1);
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
    const Bucket* base[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint64_t v = fwd[i] & T.kmask;
        key[i] = T.canonical ? canonical(v, T.k) : v;
        base[i] = locate(T, key[i], &b[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) load_bucket(base[i] + b[i], k0[i], k1[i], c0[i], c1[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        uint32_t r = 0;
        for (;;) {
            if (k0[i] == key[i]) { r = c0[i]; break; }
            if (k1[i] == key[i]) { r = c1[i]; break; }
            if (k0[i] == KM_EMPTY_KEY || k1[i] == KM_EMPTY_KEY) break;
            if (++b[i] == T.n_buckets) b[i] = 0;      // rare: full bucket without the key
            load_bucket(base[i] + b[i], k0[i], k1[i], c0[i], c1[i]);
        }
        out[i] = r;
    }
}

// N independent lookups in flight, only where `mask` has the bit set (others return 0).
template <int N>
KM_HD void table_query_masked(const TableView& T, const uint64_t (&fwd)[N], uint32_t mask, uint32_t (&out)[N]) {
    uint64_t key[N], b[N], k0[N], k1[N];
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
        if (mask & (1u << i)) load_bucket(base[i] + b[i], k0[i], k1[i], c0[i], c1[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        uint32_t r = 0;
        if (mask & (1u << i)) {
            for (;;) {
                if (k0[i] == key[i]) { r = c0[i]; break; }
                if (k1[i] == key[i]) { r = c1[i]; break; }
                if (k0[i] == KM_EMPTY_KEY || k1[i] == KM_EMPTY_KEY) break;
                if (++b[i] == T.n_buckets) b[i] = 0;      // rare: full bucket without the key
                load_bucket(base[i] + b[i], k0[i], k1[i], c0[i], c1[i]);
            }
        }
        out[i] = r;
    }
}

enum InsertMode { KM_INSERT_KEEP = 0, KM_INSERT_OVERWRITE = 1, KM_INSERT_ADD = 2 };

// Returns 1 if the key was newly inserted, 0 if it already existed (or belongs to another shard),
// -1 if the shard is full.  Only the owner inserts a key.
KM_HD int table_insert(const TableView& t, uint64_t key, uint32_t count, int mode) {
    const uint64_t h = key_hash(key);
    if (shard_of_hash(h, t.n_shards) != t.my_shard) return 0;
    uint64_t b = bucket_of_hash(h, t.n_shards, t.n_buckets);
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
