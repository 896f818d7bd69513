"""Synthetic workload of BASELINE.json config 4: a panel of targets with
planted SNV / insertion / deletion / tandem-duplication variants, the k-mer counts a
sample carrying them would produce, and the definition of the pseudo-random
background table (built on device by km_table_build_synthetic, decided analytically
by the CPU oracle).

Host-side numpy only; nothing here is on the timed path.
"""
import numpy as np

MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
TABLE_SEED = 20240001
PANEL_SEED = 20240002
QUERY_SEED = 20240003
GOLDEN = 0x9E3779B97F4A7C15


def encode(seq):
    """ASCII -> 2-bit codes (A0 C1 G2 T3); any other letter -> 255."""
    lut = np.full(256, 255, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    return lut[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]


def decode(codes):
    return BASES[codes].tobytes().decode("ascii")


def pack_kmers(codes, k):
    """All k-mers of a code array as uint64, first base most significant."""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    win = np.lib.stride_tricks.sliding_window_view(codes.astype(np.uint64), k)
    shifts = (np.uint64(2) * np.arange(k - 1, -1, -1, dtype=np.uint64))
    return np.bitwise_or.reduce(win << shifts, axis=1)


def revcomp(v, k):
    v = ~np.asarray(v, dtype=np.uint64)
    m = np.uint64
    v = ((v >> m(2)) & m(0x3333333333333333)) | ((v & m(0x3333333333333333)) << m(2))
    v = ((v >> m(4)) & m(0x0F0F0F0F0F0F0F0F)) | ((v & m(0x0F0F0F0F0F0F0F0F)) << m(4))
    v = ((v >> m(8)) & m(0x00FF00FF00FF00FF)) | ((v & m(0x00FF00FF00FF00FF)) << m(8))
    v = ((v >> m(16)) & m(0x0000FFFF0000FFFF)) | ((v & m(0x0000FFFF0000FFFF)) << m(16))
    v = (v >> m(32)) | (v << m(32))
    return v >> m(64 - 2 * k)


def canonical(v, k):
    v = np.asarray(v, dtype=np.uint64)
    return np.minimum(v, revcomp(v, k))


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def background_keys(seed, start, n, k=31):
    """key_i = canonical(mix64(seed + (i+1)*GOLDEN) & mask) for i in [start, start+n)."""
    i = np.arange(start + 1, start + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        raw = mix64(np.uint64(seed) + i * np.uint64(GOLDEN))
    mask = MASK64 if k == 32 else np.uint64((1 << (2 * k)) - 1)
    return canonical(raw & mask, k)


def background_count(keys):
    """Zipf-like count in [2, 2^20): base = 2 << min(clz(h), 18), plus low bits of h."""
    h = mix64(np.asarray(keys, dtype=np.uint64) ^ np.uint64(0xD6E8FEB86659FD93))
    # clz via float-free bit scan
    lz = np.zeros(h.shape, dtype=np.int64)
    x = h.copy()
    for s in (32, 16, 8, 4, 2, 1):
        top = x >> np.uint64(64 - s)
        z = top == 0
        lz += np.where(z, s, 0)
        x = np.where(z, x << np.uint64(s), x)
    lz = np.where(h == 0, 64, lz)
    lz = np.minimum(lz, 18)
    base = (np.uint64(2) << lz.astype(np.uint64))
    return (base + (h & np.uint64(0xFFFFFFFF) & (base - np.uint64(1)))).astype(np.uint32)


class Panel:
    """targets[i] (str), names[i], truth[i] (dict), and the planted sample content as
    canonical keys + counts (what `jellyfish count -C` would have stored)."""

    def __init__(self, k):
        self.k = k
        self.targets = []
        self.names = []
        self.truth = []
        self.alleles = []          # per target: [(codes uint8[], depth), ...], the reference allele first
        self.keys = np.zeros(0, dtype=np.uint64)
        self.counts = np.zeros(0, dtype=np.uint32)

    def n_ref_kmers(self):
        return sum(len(t) - self.k + 1 for t in self.targets)


KINDS = ("snv", "ins", "del", "dup", "none")
KIND_P = (0.40, 0.20, 0.20, 0.10, 0.10)


def _unique_kmers(codes, k):
    km = pack_kmers(codes, k)
    return len(np.unique(km)) == len(km)


def make_panel(n_targets, seed=PANEL_SEED, k=31, len_lo=62, len_hi=400, two_variant_frac=0.0):
    """SURVEY.md 8d "Config 4 -- panel".  Counts follow a two-allele read model: every
    occurrence of a k-mer in the reference allele adds round(E*(1-VAF)), every occurrence
    in the variant allele adds round(E*VAF); k-mers shared by both alleles therefore sit
    near E, reference k-mers spanning the site near E*(1-VAF), variant-only k-mers near
    E*VAF.  ``two_variant_frac`` adds a second, independent SNV allele to that fraction
    of targets (exercises multi-path clusters)."""
    rng = np.random.default_rng(seed)
    p = Panel(k)
    all_keys, all_counts = [], []
    for t in range(n_targets):
        kind = KINDS[rng.choice(len(KINDS), p=KIND_P)]
        while True:
            L = int(rng.integers(len_lo, len_hi + 1))
            if kind == "snv":
                span, ok = 1, L >= 2 * k + 1
            elif kind == "ins":
                span = int(rng.integers(1, 31)); ok = L >= 2 * k
            elif kind == "del":
                span = int(rng.integers(1, 31)); ok = L >= 2 * k + span
            elif kind == "dup":
                span = int(rng.integers(15, 151)); ok = L >= 2 * k + span
            else:
                span, ok = 0, True
            if not ok:
                continue
            ref = rng.integers(0, 4, size=L, dtype=np.uint8)
            if not _unique_kmers(ref, k):
                continue
            break
        E = float(np.exp(rng.uniform(np.log(50.0), np.log(5000.0))))
        vaf = float(rng.uniform(0.05, 0.5))
        info = {"kind": kind, "E": E, "vaf": vaf, "len": L}
        alleles = []
        if kind == "snv":
            pos = int(rng.integers(k, L - k))
            alt = ref.copy()
            alt[pos] = (alt[pos] + rng.integers(1, 4)) % 4
            info.update(pos=pos)
            alleles.append(alt)
        elif kind == "ins":
            pos = int(rng.integers(k, L - k + 1))
            ins = rng.integers(0, 4, size=span, dtype=np.uint8)
            alleles.append(np.concatenate([ref[:pos], ins, ref[pos:]]))
            info.update(pos=pos, size=span)
        elif kind == "del":
            pos = int(rng.integers(k, L - k - span + 1))
            alleles.append(np.concatenate([ref[:pos], ref[pos + span:]]))
            info.update(pos=pos, size=span)
        elif kind == "dup":
            # tandem duplication of ref[pos-span:pos] inserted at pos
            pos = int(rng.integers(max(k, span), L - k + 1))
            alleles.append(np.concatenate([ref[:pos], ref[pos - span:pos], ref[pos:]]))
            info.update(pos=pos, size=span)
        if alleles and rng.random() < two_variant_frac and L >= 2 * k + 1:
            pos2 = int(rng.integers(k, L - k))
            alt2 = ref.copy()
            alt2[pos2] = (alt2[pos2] + rng.integers(1, 4)) % 4
            alleles.append(alt2)
            info.update(pos2=pos2)
        if alleles:
            c_alt = int(E * vaf / len(alleles) + 0.5)
            c_ref = int(E * (1.0 - vaf) + 0.5)
        else:
            c_alt, c_ref = 0, int(E + 0.5)
        km = [pack_kmers(ref, k)]
        cn = [np.full(len(km[0]), c_ref, dtype=np.int64)]
        for a in alleles:
            ka = pack_kmers(a, k)
            km.append(ka)
            cn.append(np.full(len(ka), c_alt, dtype=np.int64))
        all_keys.append(canonical(np.concatenate(km), k))
        all_counts.append(np.concatenate(cn))
        p.alleles.append([(ref, c_ref)] + [(a, c_alt) for a in alleles])
        p.targets.append(decode(ref))
        p.names.append("synth_%05d" % t)
        p.truth.append(info)
    keys = np.concatenate(all_keys) if all_keys else np.zeros(0, np.uint64)
    cnts = np.concatenate(all_counts) if all_counts else np.zeros(0, np.int64)
    uk, inv = np.unique(keys, return_inverse=True)
    uc = np.zeros(len(uk), dtype=np.int64)
    np.add.at(uc, inv, cnts)
    keep = uc >= 2                                   # jellyfish count -L 2 (run_leucegene.sh:22)
    p.keys = uk[keep]
    p.counts = np.minimum(uc[keep], (1 << 24) - 1).astype(np.uint32)
    return p


def lookup_queries(n, table_seed, table_n, seed=QUERY_SEED, k=31):
    """SURVEY.md 8d lookup microbenchmark: 50 % sampled background keys, 50 % random, all
    submitted as NON-canonical forward k-mers (the kernel pays for reverse-complement+min)."""
    rng = np.random.default_rng(seed)
    mask = MASK64 if k == 32 else np.uint64((1 << (2 * k)) - 1)
    half = n // 2
    idx = rng.integers(0, max(1, table_n), size=half, dtype=np.uint64)
    with np.errstate(over="ignore"):
        hit = mix64(np.uint64(table_seed) + (idx + np.uint64(1)) * np.uint64(GOLDEN)) & mask
    flip = rng.random(half) < 0.5
    hit = np.where(flip, revcomp(hit, k), hit)
    miss = rng.integers(0, 1 << 62, size=n - half, dtype=np.uint64) & mask
    q = np.concatenate([hit, miss])
    rng.shuffle(q)
    return q


# ---- config 5: the reads a sample carrying the panel's alleles would produce ------------------------------------
READ_LEN = 100


def allele_reads(codes, depth, k=31, read_len=READ_LEN):
    """The reads of one allele sequenced to `depth`: `depth` tiling passes, pass j cut at phase (37 j) mod S with
    S = read_len - k + 1, so that every pass covers every k-mer of the allele exactly once (reads of one pass
    overlap by k - 1 bases) and a k-mer's count is the sum of the depths of the alleles that contain it -- the
    two-allele model of make_panel, now as actual reads.  Returns (blocks, multiplicities): blocks[phi] = the
    newline-terminated reads of a pass with phase phi (bytes), multiplicities[phi] = how many passes have it."""
    S = read_len - k + 1
    seq = decode(codes).encode("ascii")
    n_k = len(seq) - k + 1
    if n_k <= 0 or depth <= 0:
        return [], []
    blocks, mult = [], []
    full, rest = divmod(depth, S)
    inv = pow(37, -1, S)                       # pass j has phase (37 j) mod S  <=>  j = inv * phase mod S
    for phi in range(S):
        m = full + (1 if (inv * phi) % S < rest else 0)
        if m == 0:
            continue
        cuts = ([0] if phi > 0 else []) + list(range(phi, n_k, S)) + [n_k]
        reads = [seq[a:b + k - 1] for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        blocks.append(b"\n".join(reads) + b"\n")
        mult.append(m)
    return blocks, mult


def sample_reads(panel, targets=None, read_len=READ_LEN):
    """Byte stream (reads separated by newlines) of the sample that carries `targets` (indices; default all) of the
    panel.  Counting its canonical k-mers and dropping counts < 2 (`jellyfish count -C -L 2`) gives exactly
    panel.keys / panel.counts restricted to those targets."""
    idx = range(len(panel.targets)) if targets is None else targets
    parts = []
    for t in idx:
        for codes, depth in panel.alleles[t]:
            blocks, mult = allele_reads(codes, int(depth), panel.k, read_len)
            for b, m in zip(blocks, mult):
                parts.append(b * m)
    return b"".join(parts)
