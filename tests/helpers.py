"""Shared by the emulation (CPU) and GPU parity tests: turn a BatchResult-shaped object into the
comparable record the goldens use (rows, raw floats, node set, alt sequences)."""
from km_b200 import engine
from km_b200.utils import PathQuant as upq
from km_b200.utils import common as uc


def record_of(res, t, db_name, query_name):
    fields = res.row_fields(t, db_name, query_name)
    paths = [upq.Path(*f) for f in fields]
    paths = sorted(paths, key=lambda x: uc.natsortkey(*x[11].split(" "), x[1], x[3], x[2], x[6], rev_ix=[0]))
    kmers, counts = res.kmers(t), res.counts(t)
    alt = []
    for p in res.paths(t):
        alt.append(kmers[p[0]] + "".join(kmers[i][-1] for i in p[1:]))
    return {"rows": [str(p) for p in paths],
            "raw": [[float(p.rVAF), float(p.expression), float(p.ref_expression)] for p in paths],
            "nodes": sorted([k, int(c)] for k, c in zip(kmers[:-2], counts[:-2])),
            "alt_sequences": sorted(alt)}


def _kmers_of(seq, k=31):
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    out = []
    for i in range(len(seq) - k + 1):
        v = 0
        for ch in seq[i:i + k]:
            v = (v << 2) | code[ch]
        out.append(v)
    return out


def _canonical(v, k=31):
    rc = 0
    x = v
    for _ in range(k):
        rc = (rc << 2) | (3 - (x & 3))
        x >>= 2
    return min(v, rc)


def wide_cluster_case(n_alleles, seed=100):
    """A 140-base reference and the k-mer counts of a sample carrying n substitutions 4 bases apart, each on its own
    haplotype: one cluster of n variants, an (n + 1)-column least-squares problem.  Returns (reference, keys, counts)."""
    import numpy as np
    rng = np.random.default_rng(seed + n_alleles)
    while True:
        ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=140))
        if len(set(_kmers_of(ref))) == len(ref) - 30:
            break
    sites = [60 + 4 * j for j in range(n_alleles)]
    counts = {}

    def add(seq, c):
        for v in _kmers_of(seq):
            counts[_canonical(v)] = counts.get(_canonical(v), 0) + c
    add(ref, 400)
    for j, pos in enumerate(sites):
        alt = "ACGT"[("ACGT".index(ref[pos]) + 1 + j % 3) % 4]
        add(ref[:pos] + alt + ref[pos + 1:], 60 + 25 * j)
    return ref, np.array(list(counts.keys()), dtype=np.uint64), np.array(list(counts.values()), dtype=np.uint32)
