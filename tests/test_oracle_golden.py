"""Pins the oracle (oracle/) against the reference: golden records produced by the
UNMODIFIED reference (tests/golden/make_golden.py) and the known answers asserted by the
reference's own tests (km/tests/test_main.py, cited per check).  CPU only."""
import os

import numpy as np
import pytest

from oracle import jf_format, km_oracle as ko
from oracle.compare import compare_rows
from oracle.store import KmerStore, lib


def _finder(root, target_rel, sample, walk="dfs"):
    jf_path = "./data/jf/%s.jf" % sample
    store = KmerStore.from_jf(os.path.join(root, jf_path))
    jf = ko.OracleJellyfish(store, jf_path, 0.05, 5)
    tg = ko.Target.from_fasta(os.path.join(root, "data/catalog", target_rel), store.k)
    return ko.OracleFinder(tg, jf, walk=walk).run()


def test_jf_reader_counts_match_reference_min_cov(bundled):
    # km/tests/test_main.py:581-652 (test_min_cov): FLT3 target x two samples
    seq = "".join(ko.read_fasta_records(os.path.join(bundled, "data/catalog/GRCh38/FLT3-ITD_exons_13-15.fa"))[0])
    s = KmerStore.from_jf(os.path.join(bundled, "data/jf/03H112_IandI.jf"))
    c = [s.query(seq[i:i + 31]) for i in range(len(seq) - 30)]
    assert (sum(c), min(c), max(c), len(c), c.count(0)) == (275596, 618, 1368, 315, 0)
    assert "%.2f" % (sum(c) / len(c)) == "874.91"
    s = KmerStore.from_jf(os.path.join(bundled, "data/jf/02H025_NPM1.jf"))
    c = [s.query(seq[i:i + 31]) for i in range(len(seq) - 30)]
    assert sum(c) == 0 and c.count(0) == 315


def test_jf_record_counts(bundled):
    # SURVEY.md Appendix A
    want = {"02H025_NPM1": 1938, "02H033_DNMT3A_sub": 209, "03H112_IandI": 1604,
            "03H116_ITD": 2560, "05H094_FLT3-TKD_del": 274}
    for name, n in want.items():
        h, keys, counts = jf_format.read_jf(os.path.join(bundled, "data/jf/%s.jf" % name))
        assert len(keys) == n and h["canonical"] is True and h["key_len"] == 62
        assert counts.min() >= 2
        assert len(np.unique(keys)) == n


def test_jf_writer_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 1 << 62, size=1000, dtype=np.uint64)
    counts = rng.integers(1, 1 << 32, size=1000, dtype=np.uint64)
    p = str(tmp_path / "x.jf")
    jf_format.write_jf(p, keys, counts)
    h, k2, c2 = jf_format.read_jf(p)
    assert h["key_len"] == 62 and (k2 == keys).all() and (c2 == counts).all()


def test_revcomp_and_canonical_agree_with_strings():
    rng = np.random.default_rng(1)
    L = lib()
    for _ in range(200):
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, size=31))
        v = jf_format.pack(s)
        assert jf_format.unpack(v, 31) == s
        assert jf_format.revcomp_packed(v, 31) == jf_format.pack(jf_format.revcomp(s))
        assert L.ks_revcomp(v, 31) == jf_format.pack(jf_format.revcomp(s))
        # integer min == lexicographic min
        assert min(v, jf_format.revcomp_packed(v, 31)) == jf_format.pack(jf_format.canonical_str(s))


def test_known_answers_of_reference_tests(bundled):
    # test_NPM1 (km/tests/test_main.py:36-139): cluster row
    f = _finder(bundled, "GRCh38/NPM1_4ins_exons_10-11utr.fa", "02H025_NPM1")
    rows = [r.cells() for r in f.get_paths()]
    clus = [c for c in rows if c[11].startswith("cluster")]
    assert clus[0][2] == "Insertion" and clus[0][3] == "45:/TCTG:45"
    assert clus[0][8] == "CGGATGACTGACCAAGAGGCTATTCAAGATCTCTGTCTGGCAGTGGAGGAAGTCTCTTTAAGAAAATAG"
    vs = [c for c in rows if c[11] == "vs_ref" and c[2] != "Reference"][0]
    assert (vs[4], vs[5], vs[9], vs[6]) == ("0.484", "2870.6", "3055.2", "2428")
    # SURVEY.md 8c raw float goldens
    q = [q for kind, alt, q in f.quants if kind == "vs_ref" and tuple(alt) != f.ref_index][0]
    assert abs(q.coef[0] - 2870.598870056498) < 1e-6 and abs(q.coef[1] - 3055.1525423728817) < 1e-6
    # test_FLT3_ITD (:249-361), test_FLT3_IandI (:141-247)
    f = _finder(bundled, "GRCh38/FLT3-ITD_exons_13-15.fa", "03H116_ITD")
    vs = [r.cells() for r in f.get_paths() if r.note == "vs_ref" and not r.name.startswith("Reference")][0]
    assert vs[2] == "ITD" and vs[3].startswith("204:/") and vs[3].endswith(":204")
    assert len(vs[3].split("/")[1].split(":")[0]) == 75
    assert (vs[4], vs[5], vs[9], vs[6]) == ("0.276", "417.6", "1096.7", "443")
    f = _finder(bundled, "GRCh38/FLT3-ITD_exons_13-15.fa", "03H112_IandI")
    vs = [r.cells() for r in f.get_paths() if r.note == "vs_ref" and not r.name.startswith("Reference")][0]
    assert vs[2] == "ITD" and vs[3].startswith("152:/") and len(vs[3].split("/")[1].split(":")[0]) == 93
    assert (vs[4], vs[5], vs[9], vs[6]) == ("0.500", "399.1", "398.5", "285")
    # test_FLT3_TKD (:363-448), test_DNMT3A (:450-553)
    f = _finder(bundled, "GRCh38/FLT3-TKD_exon_20.fa", "05H094_FLT3-TKD_del")
    vs = [r.cells() for r in f.get_paths() if r.note == "vs_ref" and not r.name.startswith("Reference")][0]
    assert (vs[2], vs[3]) == ("Deletion", "32:gat/:35")
    f = _finder(bundled, "GRCh38/DNMT3A_R882_exon_23.fa", "02H033_DNMT3A_sub")
    vs = [r.cells() for r in f.get_paths() if r.note == "vs_ref" and not r.name.startswith("Reference")][0]
    assert (vs[2], vs[3], vs[4], vs[5], vs[9], vs[6]) == ("Substitution", "33:c/T:34", "0.409", "33.7", "48.7", "32")


def test_non_linear_target_raises():
    # test_not_linear (km/tests/test_main.py:555-561)
    with pytest.raises(ValueError):
        ko.ref_kmers("A" * 32, "polyA", 31)


@pytest.mark.parametrize("walk", ["dfs", "closure"])
def test_bundled_records_match_reference(bundled, bundled_golden, walk):
    flips = 0
    stores = {}
    for rec in bundled_golden:
        key = rec["sample"]
        if key not in stores:
            stores[key] = KmerStore.from_jf(os.path.join(bundled, "data/jf/%s.jf" % key))
        jf_path = "./data/jf/%s.jf" % key
        jf = ko.OracleJellyfish(stores[key], jf_path, 0.05, 5)
        tg = ko.Target.from_fasta(os.path.join(bundled, "data/catalog", rec["catalog"], rec["target"] + ".fa"), 31)
        f = ko.OracleFinder(tg, jf, walk=walk).run()
        rows = f.get_paths()
        assert sorted([k, int(v)] for k, v in f.node_data.items()) == rec["nodes"], rec["target"]
        assert sorted(ko.spell(f.kmer, a, True) for a in f.alt_paths) == rec["alt_sequences"]
        errs, fl = compare_rows(rec["rows"], [str(r) for r in rows], rec["raw"],
                                [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in rows])
        assert not errs, (rec["catalog"], rec["target"], rec["sample"], errs)
        flips += fl
    assert flips <= 2


def test_all_zero_target_prints_nan_reference_row(bundled):
    # SURVEY.md D6 / 3.4 aliasing quirk: MYC x 03H116 -> Expression nan (not -1.0)
    f = _finder(bundled, "GRCh38/MYC_T58A_P59R_exon2.fa", "03H116_ITD")
    rows = [r.cells() for r in f.get_paths()]
    assert len(rows) == 1
    assert rows[0][2] == "Reference" and rows[0][4:7] == ["nan", "nan", "0"] and rows[0][9] == "nan"


@pytest.mark.parametrize("walk", ["dfs", "closure"])
def test_synthetic_panel_matches_reference(synth_small, walk):
    store = KmerStore.from_jf(synth_small["jf"])
    jf = ko.OracleJellyfish(store, "synth_small.jf", 0.05, 5)
    flips = 0
    kinds = set()
    for fn, rec in zip(synth_small["files"], synth_small["records"]):
        tg = ko.Target.from_fasta(fn, 31)
        f = ko.OracleFinder(tg, jf, walk=walk).run()
        rows = f.get_paths()
        assert sorted([k, int(v)] for k, v in f.node_data.items()) == rec["nodes"], rec["target"]
        errs, fl = compare_rows(rec["rows"], [str(r) for r in rows], rec["raw"],
                                [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in rows])
        assert not errs, (rec["target"], errs)
        flips += fl
        kinds.update(r.cells()[2] for r in rows)
    assert flips <= 4
    assert {"Reference", "Substitution", "Insertion", "Deletion", "ITD"} <= kinds


def test_analytic_background_equals_enumerated_background():
    """The synthetic table of BASELINE config 4 is decided analytically on the CPU; check
    the inversion against explicit enumeration and against the numpy generator."""
    from km_b200 import synth
    n, seed = 50000, synth.TABLE_SEED
    keys = synth.background_keys(seed, 0, n)
    cnt = synth.background_count(keys)
    L = lib()
    for i in range(0, n, 997):
        assert L.ks_synth_key(seed, i, 31) == int(keys[i])
        assert L.ks_synth_count(int(keys[i])) == int(cnt[i])
    assert cnt.min() >= 2 and cnt.max() < (1 << 20)
    analytic = KmerStore(31, True)
    analytic.set_background(seed, n)
    explicit = KmerStore(31, True, n)
    explicit.insert(keys, cnt)
    rng = np.random.default_rng(5)
    probe = np.concatenate([keys[rng.integers(0, n, 2000)],
                            synth.revcomp(keys[rng.integers(0, n, 2000)], 31),
                            rng.integers(0, 1 << 62, size=4000, dtype=np.uint64),
                            synth.background_keys(seed, n, 2000)])   # just past the end: absent
    a = analytic.query_batch(probe)
    b = explicit.query_batch(probe)
    assert (a == b).all()
    assert (a[:4000] > 0).all() and (a[-2000:] == 0).all()
