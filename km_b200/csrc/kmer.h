// 2-bit k-mer arithmetic: A0 C1 G2 T3, first base in the most significant bits -- the
// encoding of Jellyfish's MerDNA as found in the bundled binary/sorted files (SURVEY.md
// Appendix A).  canonical() replaces MerDNA.canonicalize() (km/utils/Jellyfish.py:50-52).
#pragma once
#include "exec_model.h"

namespace km {

KM_HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

KM_HD uint64_t kmer_mask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// bitwise reverse complement: complement = ~, then reverse the 2-bit groups
KM_HD uint64_t revcomp(uint64_t v, int k) {
    v = ~v;
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
#if KM_DEVICE_BUILD
    // bytes reversed with two PRMTs + a word swap
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __byte_perm(lo, 0, 0x0123);
    hi = __byte_perm(hi, 0, 0x0123);
    v = ((uint64_t)lo << 32) | hi;
#else
    v = __builtin_bswap64(v);
#endif
    return v >> (64 - 2 * k);
}

KM_HD uint64_t canonical(uint64_t v, int k) {
    uint64_t rc = revcomp(v, k);
    return rc < v ? rc : v;
}

// successor / predecessor k-mers (Jellyfish.get_child: seq[1:]+c / c+seq[:-1], Jellyfish.py:63-66)
KM_HD uint64_t succ_kmer(uint64_t v, int c, uint64_t mask) { return ((v << 2) | (uint64_t)c) & mask; }
KM_HD uint64_t pred_kmer(uint64_t v, int c, int k) { return (v >> 2) | ((uint64_t)c << (2 * (k - 1))); }

}  // namespace km
