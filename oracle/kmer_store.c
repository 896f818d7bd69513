/* ORACLE / TEST INFRASTRUCTURE -- CPU k-mer count store.
 *
 * Stands in for the Jellyfish C++ library (QueryMerFile / MerDNA), the one
 * native dependency on km's find_mutation path.  Jellyfish itself is NOT in
 * /root/reference (un-vendored; pyproject.toml:10 pyjellyfish>=1.3.0, CI built
 * 2.2.6 at .travis.yml:20-22), so this restates the behaviour km relies on at
 * its call sites:
 *   km/utils/Jellyfish.py:24-25  QueryMerFile(fn), MerDNA.k()
 *   km/utils/Jellyfish.py:50-53  MerDNA(seq); canonicalize(); jf[kmer] -> count, 0 if absent
 * and is pinned by the raw counts asserted in km/tests/test_main.py:581-652.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline /
 * --impl reference) may load this; km_b200/ never does.
 *
 * It also implements the ANALYTIC synthetic background used by BASELINE.json
 * config 4: key_i = canonical(splitmix64_i(seed) & mask), i < n, whose membership
 * is decided by inverting splitmix64 -- so a 2e9-key background costs no memory
 * on the CPU and can be compared exactly with the GPU table built from the
 * same definition (km_b200/csrc/synth.cuh).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KS_EMPTY 0xFFFFFFFFFFFFFFFFull

typedef struct ks {
    int k;
    int canonical;
    uint64_t mask;      /* 2k low bits */
    uint64_t cap;       /* power of two */
    uint64_t size;
    uint64_t *keys;
    uint32_t *vals;
    int bg_on;
    uint64_t bg_seed, bg_n;
} ks_t;

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline uint64_t unxorshift(uint64_t x, int s) {
    uint64_t r = x;
    for (int i = s; i < 64; i += s) r = x ^ (r >> s);
    return r;
}

static inline uint64_t unmix64(uint64_t z) {
    z = unxorshift(z, 31);
    z *= 0x319642B2D24D8EC3ull; /* inverse of 0x94D049BB133111EB mod 2^64 */
    z = unxorshift(z, 27);
    z *= 0x96DE1B173F119089ull; /* inverse of 0xBF58476D1CE4E5B9 mod 2^64 */
    z = unxorshift(z, 30);
    return z;
}

#define GOLDEN 0x9E3779B97F4A7C15ull
#define GOLDEN_INV 0xF1DE83E19937733Dull /* GOLDEN * GOLDEN_INV == 1 mod 2^64 */

uint64_t ks_revcomp(uint64_t v, int k) {
    v = ~v;
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
    v = ((v >> 8) & 0x00FF00FF00FF00FFull) | ((v & 0x00FF00FF00FF00FFull) << 8);
    v = ((v >> 16) & 0x0000FFFF0000FFFFull) | ((v & 0x0000FFFF0000FFFFull) << 16);
    v = (v >> 32) | (v << 32);
    return v >> (64 - 2 * k);
}

static inline uint64_t canon(const ks_t *s, uint64_t v) {
    if (!s->canonical) return v;
    uint64_t rc = ks_revcomp(v, s->k);
    return rc < v ? rc : v;
}

/* ---- synthetic background definition (shared with the GPU generator) ---- */
uint64_t ks_synth_raw(uint64_t seed, uint64_t i) { return mix64(seed + (i + 1) * GOLDEN); }

uint64_t ks_synth_key(uint64_t seed, uint64_t i, int k) {
    uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t v = ks_synth_raw(seed, i) & mask;
    uint64_t rc = ks_revcomp(v, k);
    return rc < v ? rc : v;
}

/* Zipf-like count in [2, 2^20): Pareto(alpha=1) on the leading-zero count,
 * integer-only so CPU and GPU agree bit for bit. */
uint32_t ks_synth_count(uint64_t key) {
    uint64_t h = mix64(key ^ 0xD6E8FEB86659FD93ull);
    int lz = h ? __builtin_clzll(h) : 64;
    if (lz > 18) lz = 18;
    uint32_t base = 2u << lz;
    return base + ((uint32_t)h & (base - 1));
}

static int bg_member_fwd(const ks_t *s, uint64_t v) {
    int free_bits = 64 - 2 * s->k;
    uint64_t ntop = 1ull << free_bits;
    for (uint64_t top = 0; top < ntop; ++top) {
        uint64_t full = free_bits ? (v | (top << (2 * s->k))) : v;
        uint64_t st = unmix64(full);
        uint64_t ip1 = (st - s->bg_seed) * GOLDEN_INV;
        if (ip1 >= 1 && ip1 <= s->bg_n) return 1;
    }
    return 0;
}

static int bg_member(const ks_t *s, uint64_t key) {
    if (bg_member_fwd(s, key)) return 1;
    if (s->canonical) {
        uint64_t rc = ks_revcomp(key, s->k);
        if (rc != key && bg_member_fwd(s, rc)) return 1;
    }
    return 0;
}

/* ---- open addressing ---- */
static int ks_grow(ks_t *s, uint64_t newcap);

ks_t *ks_create(int k, int canonical, uint64_t capacity_hint) {
    if (k < 1 || k > 32) return NULL;
    ks_t *s = (ks_t *)calloc(1, sizeof(ks_t));
    if (!s) return NULL;
    s->k = k;
    s->canonical = canonical;
    s->mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t cap = 1024;
    while (cap < capacity_hint * 2) cap <<= 1;
    if (ks_grow(s, cap)) { free(s); return NULL; }
    return s;
}

void ks_destroy(ks_t *s) {
    if (!s) return;
    free(s->keys);
    free(s->vals);
    free(s);
}

static inline uint64_t slot_of(const ks_t *s, uint64_t key) { return mix64(key + GOLDEN) & (s->cap - 1); }

static void put(ks_t *s, uint64_t key, uint32_t val, int overwrite) {
    uint64_t i = slot_of(s, key);
    for (;;) {
        if (s->keys[i] == KS_EMPTY) { s->keys[i] = key; s->vals[i] = val; s->size++; return; }
        if (s->keys[i] == key) { if (overwrite) s->vals[i] = val; return; }
        i = (i + 1) & (s->cap - 1);
    }
}

static int ks_grow(ks_t *s, uint64_t newcap) {
    uint64_t *ok = s->keys; uint32_t *ov = s->vals; uint64_t oc = s->cap;
    s->keys = (uint64_t *)malloc(newcap * sizeof(uint64_t));
    s->vals = (uint32_t *)malloc(newcap * sizeof(uint32_t));
    if (!s->keys || !s->vals) return -1;
    memset(s->keys, 0xFF, newcap * sizeof(uint64_t));
    s->cap = newcap;
    s->size = 0;
    for (uint64_t i = 0; i < oc; ++i)
        if (ok && ok[i] != KS_EMPTY) put(s, ok[i], ov[i], 1);
    free(ok); free(ov);
    return 0;
}

/* keys are stored as given (the caller passes canonical keys for a canonical store,
 * exactly like the records of a `jellyfish count -C` database). */
int ks_insert(ks_t *s, const uint64_t *keys, const uint32_t *counts, uint64_t n, int overwrite) {
    for (uint64_t j = 0; j < n; ++j) {
        if ((s->size + 1) * 2 > s->cap && ks_grow(s, s->cap * 2)) return -1;
        put(s, keys[j] & s->mask, counts[j], overwrite);
    }
    return 0;
}

void ks_set_background(ks_t *s, uint64_t seed, uint64_t n) {
    s->bg_on = n > 0; s->bg_seed = seed; s->bg_n = n;
}

uint64_t ks_size(const ks_t *s) { return s->size; }
int ks_k(const ks_t *s) { return s->k; }

/* forward-strand packed k-mer -> count; canonicalises like Jellyfish.query
 * (km/utils/Jellyfish.py:47-53); absent -> 0. */
uint32_t ks_query_packed(const ks_t *s, uint64_t v) {
    uint64_t key = canon(s, v & s->mask);
    uint64_t i = slot_of(s, key);
    for (;;) {
        uint64_t kk = s->keys[i];
        if (kk == key) return s->vals[i];
        if (kk == KS_EMPTY) break;
        i = (i + 1) & (s->cap - 1);
    }
    if (s->bg_on && bg_member(s, key)) return ks_synth_count(key);
    return 0;
}

void ks_query_batch(const ks_t *s, const uint64_t *kmers, uint64_t n, uint32_t *out) {
    for (uint64_t j = 0; j < n; ++j) out[j] = ks_query_packed(s, kmers[j]);
}

/* ASCII k-mer (exactly k chars of ACGT) -> count.  Non-ACGT is "parity unpinned"
 * (SURVEY.md 8c); we return 0xFFFFFFFF so the Python side can raise. */
uint32_t ks_query_ascii(const ks_t *s, const char *seq) {
    uint64_t v = 0;
    for (int i = 0; i < s->k; ++i) {
        uint64_t c;
        switch (seq[i]) {
            case 'A': c = 0; break; case 'C': c = 1; break;
            case 'G': c = 2; break; case 'T': c = 3; break;
            default: return 0xFFFFFFFFu;
        }
        v = (v << 2) | c;
    }
    return ks_query_packed(s, v);
}
