"""Kernel LOGIC vs the reference, on the CPU: the device stage functions of
km_b200/csrc/{table,walk,graph,quant}.h compiled single-lane by g++ (tests/emu/km_emu.cpp) are
run on the bundled samples and the synthetic panel and compared with the golden records of the
UNMODIFIED reference.  What this cannot show -- races, launch geometry, memory placement -- is
covered by the -m gpu tests, which call the CUDA build through the C ABI."""
import os

import numpy as np
import pytest

from oracle import jf_format, km_oracle as ko
from oracle.compare import compare_rows
from oracle.store import KmerStore, lib as olib

from emu_harness import EmuTable, lib as elib
from helpers import record_of, wide_cluster_case


def _check(rec, got, tag):
    assert got["nodes"] == rec["nodes"], tag
    assert got["alt_sequences"] == rec["alt_sequences"], tag
    errs, flips = compare_rows(rec["rows"], got["rows"], rec["raw"], got["raw"])
    assert not errs, (tag, errs)
    return flips


def test_kmer_primitives_match_oracle():
    rng = np.random.default_rng(3)
    E, O = elib(), olib()
    for v in rng.integers(0, 1 << 62, size=500, dtype=np.uint64).tolist():
        assert E.emu_revcomp(v, 31) == O.ks_revcomp(v, 31)
    for k in (5, 17, 31):
        for v in rng.integers(0, 1 << (2 * k), size=50, dtype=np.uint64).tolist():
            assert E.emu_revcomp(v, k) == jf_format.revcomp_packed(v, k)
    for i in range(0, 5000, 37):
        key = E.emu_synth_key(20240001, i, 31)
        assert key == O.ks_synth_key(20240001, i, 31)
        assert E.emu_synth_count(key) == O.ks_synth_count(key)


def test_table_lookup_matches_store(bundled):
    for name in ("02H025_NPM1", "03H116_ITD"):
        _, keys, counts = jf_format.read_jf(os.path.join(bundled, "data/jf/%s.jf" % name))
        t = EmuTable.from_keys(keys, counts.astype(np.uint32))
        s = KmerStore.from_jf(os.path.join(bundled, "data/jf/%s.jf" % name))
        rng = np.random.default_rng(1)
        probes = np.concatenate([keys[:300], rng.integers(0, 1 << 62, size=300, dtype=np.uint64)])
        for v in probes.tolist():
            assert t.query_packed(v) == s.query_packed(v)
            rc = jf_format.revcomp_packed(v, 31)
            assert t.query_packed(rc) == s.query_packed(rc)


def test_crowded_table_still_exact():
    # load factor ~0.97 of the slots: long bucket chains, every key must still be found
    rng = np.random.default_rng(9)
    keys = np.unique(rng.integers(0, 1 << 62, size=3900, dtype=np.uint64))
    keys = np.minimum(keys, np.array([jf_format.revcomp_packed(int(v), 31) for v in keys], dtype=np.uint64))
    keys = np.unique(keys)
    counts = rng.integers(1, 1 << 20, size=len(keys)).astype(np.uint32)
    lines = os.environ.get("KM_TABLE_LINES", "0") not in ("", "0")
    # sector buckets: 2000 buckets = 4000 slots; family lines: 3.2 slots per unit of capacity, two copies per key
    t = EmuTable(31, True, capacity=int(2 * len(keys) / 0.97 / 3.2) if lines else 2000)
    t.insert(keys, counts)
    for kk, c in zip(keys.tolist(), counts.tolist()):
        assert t.query_packed(kk) == c
    for v in rng.integers(0, 1 << 62, size=500, dtype=np.uint64).tolist():
        if v not in set(keys.tolist()) and jf_format.revcomp_packed(v, 31) not in set(keys.tolist()):
            assert t.query_packed(v) == 0
            break


def test_bundled_pairs_match_reference(bundled, bundled_golden):
    tables = {}
    flips = 0
    for rec in bundled_golden:
        if rec["catalog"] != "GRCh38":
            continue
        s = rec["sample"]
        if s not in tables:
            _, keys, counts = jf_format.read_jf(os.path.join(bundled, "data/jf/%s.jf" % s))
            tables[s] = EmuTable.from_keys(keys, counts.astype(np.uint32))
        seqs, _ = ko.read_fasta_records(os.path.join(bundled, "data/catalog", rec["catalog"], rec["target"] + ".fa"))
        res = tables[s].find_batch(["".join(seqs)])
        assert int(res.status[0]) & ~16 == 0
        got = record_of(res, 0, "./data/jf/%s.jf" % s, rec["target"])
        flips += _check(rec, got, (rec["target"], s))
    assert flips <= 2


def test_synthetic_panel_matches_reference(synth_small):
    t = EmuTable.from_keys(synth_small["keys"], synth_small["counts"])
    res = t.find_batch(synth_small["targets"])
    flips = 0
    for i, rec in enumerate(synth_small["records"]):
        assert int(res.status[i]) == 0
        got = record_of(res, i, "synth_small.jf", rec["target"])
        flips += _check(rec, got, rec["target"])
    assert flips <= 4


def test_lookup_accounting_matches_survey_formula(synth_small):
    # algorithmic lookups per target = n_ref_kmers + 4 * n_nodes.  The walk may issue
    # more (dead-end tips); the shared-memory walk issues n_ref - 1 fewer because the successor of a
    # reference k-mer along the reference IS the next reference k-mer, whose count it already holds.
    t = EmuTable.from_keys(synth_small["keys"], synth_small["counts"])
    res = t.find_batch(synth_small["targets"][:20])
    for i in range(20):
        n_ref = len(synth_small["targets"][i]) - 30
        algorithmic = n_ref + 4 * (int(res.n_nodes[i]) - 2)
        assert int(res.lookups[i]) >= algorithmic - (n_ref - 1)
        assert int(res.lookups[i]) <= algorithmic + 4 * 64


def test_degenerate_targets():
    t = EmuTable(31, True, 1024)
    rng = np.random.default_rng(2)
    seq = "".join("ACGT"[i] for i in rng.integers(0, 4, size=80))
    res = t.find_batch([seq, "A" * 40, seq[:31], seq[:20], seq[:40] + "N" + seq[41:]])
    # empty table: one Reference row with nan / nan / 0
    assert int(res.status[0]) == 0
    rec = record_of(res, 0, "empty.jf", "x")
    assert len(rec["rows"]) == 1
    cells = rec["rows"][0].split("\t")
    assert cells[2] == "Reference" and cells[4:7] == ["nan", "nan", "0"] and cells[9] == "nan"
    assert int(res.status[1]) & 2          # repeated k-mer -> ValueError on the host
    assert int(res.status[2]) == 0 and int(res.n_nodes[2]) == 3      # a single k-mer target
    assert int(res.status[3]) & 64         # shorter than k
    assert int(res.status[4]) & 1          # non-ACGT letter


@pytest.mark.parametrize("small", [(512, 512, 64, 8), (512, 512, 1, 8), (512, 1, 64, 8), (512, 512, 64, 2), (64, 512, 64, 8)])
def test_both_scratch_passes_agree(synth_small, small):
    """The shared-memory pass defers to the general pass when a capacity is exceeded; whichever
    pass finishes a target, the records are the reference's."""
    t = EmuTable.from_keys(synth_small["keys"], synth_small["counts"])
    res = t.find_batch(synth_small["targets"][:40], small=small)
    if small == (512, 512, 64, 8):
        assert set(t.passes) == {1}
    else:
        assert 2 in t.passes
    for i, rec in enumerate(synth_small["records"][:40]):
        assert int(res.status[i]) == 0
        _check(rec, record_of(res, i, "synth_small.jf", rec["target"]), rec["target"])


def test_long_refinement_equals_literal_iteration():
    """A tandem duplication of k-2..k-1 bases makes a cluster of two ITD paths of which lstsq gives one a
    negative coefficient; the reference's projected-gradient loop (PathQuant.py:120-142) then runs for
    thousands of iterations.  quant.h crosses the linear stretch in closed form (refine_jump): rows, raw
    floats AND the iteration count must equal the literal loop's."""
    from km_b200 import synth
    panel = synth.make_panel(2000, seed=synth.PANEL_SEED + 5)
    picks = [i for i, tr in enumerate(panel.truth) if tr["kind"] == "dup" and tr.get("size") in (28, 29, 30)][:5]
    assert len(picks) >= 3
    t = EmuTable.from_keys(panel.keys, panel.counts)
    store = KmerStore(31, True, len(panel.keys))
    store.insert(panel.keys, panel.counts)
    jf = ko.OracleJellyfish(store, "p.jf", 0.05, 5)
    seen = []
    orig = ko.Quant.solve

    def counting(self):
        r = orig(self)
        seen.append(self.n_iter)
        return r
    ko.Quant.solve = counting
    try:
        res = t.find_batch([panel.targets[i] for i in picks])
        longest = 0
        for j, i in enumerate(picks):
            del seen[:]
            f = ko.OracleFinder(ko.Target(panel.targets[i], panel.names[i], 31), jf).run()
            want = f.get_paths()
            got = record_of(res, j, "p.jf", panel.names[i])
            errs, _ = compare_rows([str(r) for r in want], got["rows"],
                                   [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in want], got["raw"])
            assert not errs, (panel.names[i], errs)
            rows = res.rows[int(res.row_first[j]):int(res.row_first[j]) + int(res.row_count[j])]
            assert max(seen) == int(rows["n_iter"].max()), (panel.names[i], max(seen), rows["n_iter"].tolist())
            longest = max(longest, max(seen))
        assert longest > 500           # the case this test is about did occur
    finally:
        ko.Quant.solve = orig


def test_exact_three_column_solve_matches_lstsq_at_every_rank():
    """quant.h solve3_exact (clusters of two variants): the minimum-norm solution np.linalg.lstsq returns
    (PathQuant.py:116), from exact integer sums, for contribution matrices of rank 3, 2 (two tandem copies next to
    one: twice = 2 * once - reference), 1 and 0."""
    import ctypes
    L = elib()
    L.emu_solve3.restype = ctypes.c_int
    L.emu_solve3.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(7)
    cases = []
    for _ in range(200):
        n = int(rng.integers(20, 400))
        ref = rng.integers(0, 2, n)
        once = np.clip(ref + rng.integers(0, 2, n) * rng.integers(-1, 2, n), 0, 2)
        third = rng.integers(0, 3, n)
        cases.append(np.stack([ref, once, third], 1))                      # rank 3 (almost surely)
        cases.append(np.stack([ref, once, 2 * once - ref + 0 * third], 1).clip(min=-5))   # rank 2, signed entries excluded below
        cases.append(np.stack([ref, 2 * ref, ref], 1))                      # rank 1
    cases.append(np.zeros((50, 3), dtype=np.int64))                         # rank 0
    ranks = set()
    for A in cases:
        if (A < 0).any():
            continue
        b = rng.integers(0, 5000, A.shape[0]).astype(np.float64)
        G = A.T @ A
        h = (A.T @ b).astype(np.int64)
        acc = np.zeros(12, dtype=np.uint64)
        for i in range(3):
            for j in range(i + 1):
                acc[i * 3 + j] = G[i, j]
        acc[9:12] = h
        x = np.zeros(3)
        assert L.emu_solve3(acc.ctypes.data, x.ctypes.data) == 1
        want = np.linalg.lstsq(A.astype(np.float64), b, rcond=None)[0]
        ranks.add(int(np.linalg.matrix_rank(A)))
        assert np.allclose(x, want, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(want).max()))), (A.shape, x, want)
    assert ranks == {0, 1, 2, 3}


@pytest.mark.parametrize("n_alleles", [2, 3, 4, 5])
def test_wide_clusters_match_the_oracle(n_alleles):
    """Several substitutions within one k-mer length of each other, each on its own haplotype: one cluster of n variants
    (MutationFinder.py:651-723, 749-811), an (n + 1)-column least-squares problem.  quant.h solves three columns exactly
    from integer cofactors, three and four with the state in registers, five and more in memory with the Jacobi
    eigen-decomposition; all must give the oracle's rows and raw floats."""
    ref, keys, vals = wide_cluster_case(n_alleles)
    t = EmuTable.from_keys(keys, vals)
    store = KmerStore(31, True, len(keys))
    store.insert(keys, vals)
    jf = ko.OracleJellyfish(store, "w.jf", 0.05, 5)
    f = ko.OracleFinder(ko.Target(ref, "wide", 31), jf).run()
    want = f.get_paths()
    res = t.find_batch([ref])
    assert int(res.status[0]) == 0
    got = record_of(res, 0, "w.jf", "wide")
    errs, _ = compare_rows([str(r) for r in want], got["rows"],
                           [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in want], got["raw"])
    assert not errs, errs
    sizes = [r for r in got["rows"] if "cluster" in r.split("\t")[11]]
    assert any("n=%d" % n_alleles in r.split("\t")[11] for r in sizes), [r.split("\t")[11] for r in got["rows"]]


def _wide_golden():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wide_clusters.json")) as f:
        return json.load(f)


def test_wide_clusters_and_long_refinements_match_the_reference_goldens():
    """The same cases against records of the UNMODIFIED reference (tests/golden/make_golden_wide.py): clusters of 2..5
    variants and the tandem duplications whose refinement runs for hundreds of steps -- node sets, alternative paths, rows,
    raw floats at 1e-6."""
    from km_b200 import synth
    g = _wide_golden()
    for case in g["wide"]:
        ref, keys, vals = wide_cluster_case(case["n"], seed=case["seed"])
        res = EmuTable.from_keys(keys, vals).find_batch([ref])
        assert int(res.status[0]) == 0
        _check(case["record"], record_of(res, 0, "w.jf", case["name"]), case["name"])
    lg = g["long"]
    panel = synth.make_panel(lg["n_targets"], seed=synth.PANEL_SEED + lg["panel_seed_offset"])
    res = EmuTable.from_keys(panel.keys, panel.counts).find_batch([panel.targets[i] for i in lg["picks"]])
    for j, (i, rec) in enumerate(zip(lg["picks"], lg["records"])):
        assert int(res.status[j]) == 0
        _check(rec, record_of(res, j, "p.jf", panel.names[i]), panel.names[i])
