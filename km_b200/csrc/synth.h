// Definition of the synthetic background table of BASELINE.json config 4:
//   key_i   = canonical(mix64(seed + (i+1)*GOLDEN) & mask),  i in [0, n)
//   count   = f(key): Pareto-like integer in [2, 2^20)
// Same definition as oracle/kmer_store.c (ks_synth_key / ks_synth_count) and
// km_b200/synth.py (background_keys / background_count).
#pragma once
#include "kmer.h"

namespace km {

#define KM_GOLDEN 0x9E3779B97F4A7C15ull

KM_HD uint64_t synth_key(uint64_t seed, uint64_t i, int k) {
    return canonical(mix64(seed + (i + 1) * KM_GOLDEN) & kmer_mask(k), k);
}

KM_HD uint32_t synth_count(uint64_t key) {
    uint64_t h = mix64(key ^ 0xD6E8FEB86659FD93ull);
    int lz = clz64(h);
    if (lz > 18) lz = 18;
    uint32_t base = 2u << lz;
    return base + ((uint32_t)h & (base - 1u));
}

}  // namespace km
