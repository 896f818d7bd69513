"""ORACLE / TEST INFRASTRUCTURE -- CPU restatement of `jellyfish count -m k -C [-L n]` on a byte stream of reads.

The reference has no counting code: km's workflow delegates it to the external `jellyfish count` binary
(example/run_leucegene.sh:22, `jellyfish count -m 31 -C -L 2 -Q+ ...`), which is absent from /root/reference
(third-party, un-vendored) -- parity UNPINNED; the semantics restated here are the four flags: every k-mer of
every read whose k letters are all ACGT (either case) is counted under its canonical form (-C); -L drops counts
below the bound on output; -Q masks bases below a quality character.  Only tests/ and bench.py (as the checker)
import this."""
import numpy as np

_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _LUT[_c] = _i
    _LUT[_c | 0x20] = _i


def count_stream(stream, k=31, canonical=True, qual=None, min_qual=0):
    """stream: bytes of sequences separated by any non-ACGT byte.  Returns (keys uint64 sorted, counts int64)."""
    buf = np.frombuffer(stream, dtype=np.uint8)
    codes = _LUT[buf]
    if qual is not None and min_qual > 0:
        q = np.frombuffer(qual, dtype=np.uint8)
        codes = np.where(q >= min_qual, codes, 255).astype(np.uint8)
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64)
    bad = (codes > 3).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(bad)])
    valid = (csum[k:] - csum[:-k]) == 0                       # window [i, i+k) holds no invalid byte
    c64 = (codes & 3).astype(np.uint64)
    fwd = np.zeros(n, dtype=np.uint64)
    for j in range(k):                                        # k passes over the array: fine for the megabytes a test counts
        fwd = (fwd << np.uint64(2)) | c64[j:j + n]
    fwd = fwd[valid]
    if canonical:
        from km_b200.synth import canonical as canon
        fwd = canon(fwd, k)
    keys, counts = np.unique(fwd, return_counts=True)
    return keys, counts.astype(np.int64)
