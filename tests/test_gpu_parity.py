"""Parity tests proper: the CUDA build, called through the C ABI (ctypes -> libkm_b200.so), against
the oracle and the golden records of the unmodified reference.  Run with -m gpu on a B200."""
import os

import numpy as np
import pytest

from oracle import jf_format, km_oracle as ko
from oracle.compare import compare_rows
from oracle.store import KmerStore

from helpers import record_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import __graft_entry__ as ge
    ge.build()
    from km_b200 import engine as e
    from km_b200._lib import lib
    assert lib().km_device_count() > 0, "these tests need a CUDA device"
    return e


def _check(rec, got, tag):
    assert got["nodes"] == rec["nodes"], tag
    assert got["alt_sequences"] == rec["alt_sequences"], tag
    errs, flips = compare_rows(rec["rows"], got["rows"], rec["raw"], got["raw"])
    assert not errs, (tag, errs)
    return flips


def test_jf_load_and_query_match_store(engine, bundled):
    rng = np.random.default_rng(1)
    for name in ("02H025_NPM1", "02H033_DNMT3A_sub", "03H112_IandI", "03H116_ITD", "05H094_FLT3-TKD_del"):
        path = os.path.join(bundled, "data/jf/%s.jf" % name)
        t = engine.Table.open_jf(path)
        s = KmerStore.from_jf(path)
        _, keys, counts = jf_format.read_jf(path)
        assert t.info()["n_keys"] == len(keys) and t.k == 31 and t.canonical
        probes = np.concatenate([keys, np.array([jf_format.revcomp_packed(int(v), 31) for v in keys[:500]], dtype=np.uint64),
                                 rng.integers(0, 1 << 62, size=5000, dtype=np.uint64)])
        got = t.query_packed(probes)
        assert (got == s.query_batch(probes)).all()
        assert (got[:len(keys)] == counts.astype(np.uint32)).all()
        t.close()


def test_min_cov_known_answers(engine, bundled):
    # km/tests/test_main.py:581-652
    from km_b200.utils import common as uc
    seq = "".join(uc.file_2_seq(os.path.join(bundled, "data/catalog/GRCh38/FLT3-ITD_exons_13-15.fa"))[0])
    res = uc.get_cov(os.path.join(bundled, "data/jf/03H112_IandI.jf"), seq)
    assert res[0] == 275596 and res[2] == 618 and res[3] == 1368 and "%.2f" % res[4] == "874.91"
    assert res[5] == 315 and res[6] == 0
    res = uc.get_cov(os.path.join(bundled, "data/jf/02H025_NPM1.jf"), seq)
    assert res[0] == 0 and res[6] == 315


def test_get_child_matches_oracle(engine, bundled):
    from km_b200.utils.Jellyfish import Jellyfish
    path = os.path.join(bundled, "data/jf/03H116_ITD.jf")
    jf = Jellyfish(path, cutoff=0.05, n_cutoff=5)
    oj = ko.OracleJellyfish(KmerStore.from_jf(path), path, 0.05, 5)
    seq = "".join(ko.read_fasta_records(os.path.join(bundled, "data/catalog/GRCh38/FLT3-ITD_exons_13-15.fa"))[0])
    kmers = [seq[i:i + 31] for i in range(len(seq) - 30)]
    for fwd in (True, False):
        got = jf.get_child_many(kmers, forward=fwd)
        for km, g in zip(kmers, got):
            assert g == oj.get_child(km, forward=fwd)
    assert jf.query(kmers[0]) == oj.query(kmers[0])
    assert jf.get_child(kmers[3]) == oj.get_child(kmers[3])
    # class defaults are 0.30 / 500 (Jellyfish.py:23), not the CLI's
    assert Jellyfish(path).cutoff == 0.30 and Jellyfish(path).n_cutoff == 500


def test_bundled_catalog_matches_reference_records(engine, bundled, bundled_golden):
    """configs 1-3: every bundled target x sample; the whole catalog goes down in ONE batch."""
    flips = 0
    for cat in ("GRCh38", "GRCh37"):
        for sample in sorted({r["sample"] for r in bundled_golden}):
            recs = [r for r in bundled_golden if r["catalog"] == cat and r["sample"] == sample]
            t = engine.Table.open_jf(os.path.join(bundled, "data/jf/%s.jf" % sample))
            seqs = ["".join(ko.read_fasta_records(os.path.join(bundled, "data/catalog", cat, r["target"] + ".fa"))[0])
                    for r in recs]
            res = t.find_batch(seqs)
            assert res.timing["launches"] in (6, 8)   # reference probe, walk, scheduling, (two bubble passes,) three graph passes for the whole catalog
            for i, rec in enumerate(recs):
                assert int(res.status[i]) & ~16 == 0
                db = "./data/jf/%s.jf" % sample
                got = record_of(res, i, db, rec["target"])
                flips += _check(rec, got, (cat, rec["target"], sample))
                # the library's own formatter prints the same sorted rows
                assert res.format_target(i, db, rec["target"]) == "".join(r + "\n" for r in got["rows"])
            t.close()
    assert flips <= 4


def test_synthetic_panel_matches_reference_records(engine, synth_small):
    t = engine.Table.create(capacity=len(synth_small["keys"]))
    t.insert(synth_small["keys"], synth_small["counts"])
    res = t.find_batch(synth_small["targets"])
    flips = 0
    for i, rec in enumerate(synth_small["records"]):
        assert int(res.status[i]) == 0
        got = record_of(res, i, "synth_small.jf", rec["target"])
        flips += _check(rec, got, rec["target"])
        assert res.format_target(i, "synth_small.jf", rec["target"]) == "".join(r + "\n" for r in got["rows"])
    assert flips <= 4
    # repeatability: a second run gives byte-identical text
    res2 = t.find_batch(synth_small["targets"])
    for i, rec in enumerate(synth_small["records"]):
        assert res.format_target(i, "x", rec["target"]) == res2.format_target(i, "x", rec["target"])


def test_synthetic_background_equals_analytic_oracle(engine):
    from km_b200 import synth
    n = 3_000_000
    t = engine.Table.create(capacity=n)
    t.build_synthetic(synth.TABLE_SEED, n)
    s = KmerStore(31, True)
    s.set_background(synth.TABLE_SEED, n)
    assert abs(t.info()["n_keys"] - n) < 10          # collisions among 3e6 draws from 2^61 are ~0
    q = synth.lookup_queries(1 << 20, synth.TABLE_SEED, n)
    got = t.query_packed(q)
    assert (got == s.query_batch(q)).all()
    assert 0.45 < (got > 0).mean() < 0.55
    # a crowded table (load ~0.9 of the slots) must still answer exactly
    # (family lines store every key twice in 3.2 slots per unit of capacity)
    t2 = engine.Table.create(capacity=int(2 * n / 0.9 / 3.2) if t.info()["layout"] else n // 2 + n // 20)
    t2.build_synthetic(synth.TABLE_SEED, n)
    assert (t2.query_packed(q) == got).all()


def test_panel_against_oracle_with_full_background(engine):
    """config 4 in miniature: planted panel + pseudo-random background; the oracle decides
    background membership analytically, so the same comparison works at 2e9 keys (bench.py)."""
    from km_b200 import synth
    n_bg = 2_000_000
    panel = synth.make_panel(300, seed=77, two_variant_frac=0.2)
    t = engine.Table.create(capacity=n_bg + len(panel.keys))
    t.build_synthetic(synth.TABLE_SEED, n_bg)
    t.insert(panel.keys, panel.counts, mode="overwrite")
    res = t.find_batch(panel.targets)
    store = KmerStore(31, True, len(panel.keys))
    store.set_background(synth.TABLE_SEED, n_bg)
    store.insert(panel.keys, panel.counts)
    jf = ko.OracleJellyfish(store, "panel.jf", 0.05, 5)
    flips = 0
    for i in range(0, 300, 3):
        f = ko.OracleFinder(ko.Target(panel.targets[i], panel.names[i], 31), jf).run()
        want = f.get_paths()
        got = record_of(res, i, "panel.jf", panel.names[i])
        errs, fl = compare_rows([str(r) for r in want], got["rows"],
                                [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in want], got["raw"])
        assert not errs, (panel.names[i], errs)
        assert got["nodes"] == sorted([k, int(v)] for k, v in f.node_data.items())
        # table reads actually issued: at least one per reference k-mer; the neighbour masks (km_table_link) answer
        # "absent" for most successors without a read, so the reference's 5 per k-mer is an upper bound only
        assert int(res.lookups[i]) >= len(panel.targets[i]) - 30
        flips += fl
    assert flips <= 4


def test_error_statuses_and_limits(engine):
    from km_b200.utils.Jellyfish import Jellyfish
    from km_b200.utils import MutationFinder as umf
    from km_b200.utils import Sequence as us
    rng = np.random.default_rng(4)
    seq = "".join("ACGT"[i] for i in rng.integers(0, 4, size=120))
    t = engine.Table.create(capacity=4096)
    res = t.find_batch([seq, "A" * 40, seq[:20], seq[:40] + "N" + seq[41:], seq[:31]])
    assert [int(s) for s in res.status] == [0, 2, 64, 1, 0]
    rec = record_of(res, 0, "e.jf", "x")
    cells = rec["rows"][0].split("\t")
    assert len(rec["rows"]) == 1 and cells[2] == "Reference" and cells[4:7] == ["nan", "nan", "0"]
    # non-linear target raises ValueError on the host, like the reference (test_not_linear)
    with pytest.raises(ValueError):
        us.RefSeq("A" * 32, "polyA", 31)
    # node limit -> sys.exit with the reference's message (MutationFinder.py:143-148)
    keys = np.array([min(jf_format.pack(seq[i:i + 31]), jf_format.revcomp_packed(jf_format.pack(seq[i:i + 31]), 31))
                     for i in range(90)], dtype=np.uint64)
    t.insert(keys, np.full(90, 100, np.uint32))
    jf = Jellyfish("e.jf", cutoff=0.05, n_cutoff=5, table=t)
    with pytest.raises(SystemExit) as ex:
        umf.MutationFinder(us.RefSeq(seq[:60], "lim", 31), jf, 500, 10, 20)
    assert "Node query count limit exceeded: max=20" in str(ex.value)
    # the walk beyond the target's end is explored and dropped (dead-end tip)
    f = umf.MutationFinder(us.RefSeq(seq[:60], "tip", 31), jf)
    assert f.num_k == 30 + 2


def test_mutation_finder_api_reads_like_the_reference(engine, bundled):
    """Mirrors km/tests/test_main.py:36-71 (test_NPM1) on the drop-in classes."""
    from km_b200.utils.Jellyfish import Jellyfish
    from km_b200.utils import MutationFinder as umf
    from km_b200.utils import Sequence as us
    from km_b200.utils import common as uc
    cwd = os.getcwd()
    os.chdir(bundled)
    try:
        jf = Jellyfish("./data/jf/02H025_NPM1.jf", cutoff=0.05, n_cutoff=5)
        seqs, _ = uc.file_2_seq("./data/catalog/GRCh38/NPM1_4ins_exons_10-11utr.fa")
        refpath = us.RefSeq("".join(seqs), "NPM1_4ins_exons_10-11utr", jf.k)
        finder = umf.MutationFinder(refpath, jf, 500, 10, 10000)
        finder.graph_analysis()
        finder.quantify_paths(False)
        finder.quantify_clusters(False)
        rows = [str(p).split("\t") for p in finder.get_paths(sort=True)]
    finally:
        os.chdir(cwd)
    assert finder.num_k == 82 and len(finder.alt_paths) == 2
    clus = rows[-1]
    assert clus[2] == "Insertion" and clus[3] == "45:/TCTG:45"
    assert clus[8] == "CGGATGACTGACCAAGAGGCTATTCAAGATCTCTGTCTGGCAGTGGAGGAAGTCTCTTTAAGAAAATAG"
    vs = [r for r in rows if r[11] == "vs_ref" and r[2] != "Reference"][0]
    assert (vs[4], vs[5], vs[9], vs[6]) == ("0.484", "2870.6", "3055.2", "2428")
    ref = [r for r in rows if r[2] == "Reference"][0]
    assert ref[3] == "" and ref[4] == "nan" and ref[5] == "-1.0" and ref[6] == "2379"


def test_cli_find_mutation_and_report_equal_reference_text(engine, bundled, bundled_cli_golden):
    """configs 1-2 through the CLI entry points: `km find_mutation ... | km find_report ...` must
    print what the reference prints (volatile '#func:' / '#Elapsed time:' lines excluded, as in
    SURVEY.md section 0), modulo printed-digit boundary flips."""
    import io
    import sys
    from argparse import Namespace
    from km_b200.tools import find_mutation as fm
    from km_b200.tools import find_report as fr
    cwd = os.getcwd()
    os.chdir(bundled)
    try:
        for case in bundled_cli_golden:
            target = "./data/catalog/GRCh38/%s.fa" % case["target"]
            args = Namespace(count=5, graphical=False, jellyfish_fn="./data/jf/%s.jf" % case["sample"], ratio=0.05,
                             steps=500, branchs=10, nodes=10000, target_fn=[target], verbose=False, debug=False)
            old = sys.stdout
            sys.stdout = buf = io.StringIO()
            try:
                fm.main_find_mut(args, None)
            finally:
                sys.stdout = old
            mine = [l for l in buf.getvalue().split("\n") if not l.startswith("#Elapsed time:")]
            want = case["find_mutation"].split("\n")
            assert [l for l in mine if l.startswith("#")] == [l for l in want if l.startswith("#")]
            body_m = [l for l in mine if l and not l.startswith("#")]
            body_w = [l for l in want if l and not l.startswith("#")]
            assert body_m[0] == body_w[0]
            errs, flips = compare_rows(body_w[1:], body_m[1:])
            assert not errs and flips == 0, (case["target"], errs)
            assert body_m == body_w          # order and every printed digit
            a = Namespace(target=target, infile=io.StringIO(buf.getvalue()), info="vs_ref", min_cov=1, exclu="", format=None)
            sys.stdout = rep = io.StringIO()
            try:
                fr.main_find_report(a, None)
            finally:
                sys.stdout = old
            assert rep.getvalue() == case["report"]["stdout"]
    finally:
        os.chdir(cwd)


def test_pipelined_text_call_equals_two_step_path(engine, synth_small):
    """km_find_text (sub-batches in flight, rows formatted while later sub-batches run) must print
    exactly what km_find_batch + km_result_format_all print, for every split."""
    from km_b200 import synth
    panel = synth.make_panel(700, seed=21, two_variant_frac=0.2)
    t = engine.Table.create(capacity=len(panel.keys) + 200000)
    t.build_synthetic(synth.TABLE_SEED, 200000)
    t.insert(panel.keys, panel.counts, mode="overwrite")
    packed = engine.PackedTargets(panel.targets, panel.names)
    res = t.find_batch(packed, want_graph=False)
    want = res.format_all("panel.jf", packed)
    assert want.count("\n") >= 700
    for n_sub in (1, 2, 3, 7):
        got, status = t.find_text(packed, "panel.jf", n_sub=n_sub)
        assert got == want, n_sub
        assert (status == res.status).all()
    got, _ = t.find_text(engine.PackedTargets(panel.targets[:5], panel.names[:5]), "panel.jf")
    assert got == t.find_batch(panel.targets[:5], want_graph=False).format_all("panel.jf", panel.names[:5])


def test_long_targets_take_the_general_kernels(engine):
    """Targets beyond the shared-memory capacities (more than 448 reference k-mers for the walk, more than
    510 graph nodes for the graph pass) go through the general walk / graph kernels: same answers."""
    from km_b200 import synth
    panel = synth.make_panel(60, seed=31, len_lo=430, len_hi=900, two_variant_frac=0.3)
    t = engine.Table.create(capacity=len(panel.keys) + 500000)
    t.build_synthetic(synth.TABLE_SEED, 500000)
    t.insert(panel.keys, panel.counts, mode="overwrite")
    res = t.find_batch(panel.targets)
    assert (np.array([len(s) for s in panel.targets]) - 30 > 448).sum() >= 20       # the general walk is exercised
    assert (res.n_nodes > 512).sum() >= 10                                           # and the general graph pass
    store = KmerStore(31, True, len(panel.keys))
    store.set_background(synth.TABLE_SEED, 500000)
    store.insert(panel.keys, panel.counts)
    jf = ko.OracleJellyfish(store, "panel.jf", 0.05, 5)
    flips = 0
    for i in range(0, 60, 2):
        f = ko.OracleFinder(ko.Target(panel.targets[i], panel.names[i], 31), jf).run()
        want = f.get_paths()
        got = record_of(res, i, "panel.jf", panel.names[i])
        errs, fl = compare_rows([str(r) for r in want], got["rows"],
                                [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in want], got["raw"])
        assert not errs, (panel.names[i], errs)
        assert got["nodes"] == sorted([k, int(v)] for k, v in f.node_data.items())
        flips += fl
    assert flips <= 4


def test_full_size_panel_properties(engine):
    """BASELINE.json config 4 at full size (10,000 targets x 2e9-key table, if the GPU has the memory):
    properties that need no oracle -- the planted variant of every target is reported with its type, the
    text does not depend on the order or the grouping of the targets, lookups match the SURVEY formula."""
    import torch
    from km_b200 import synth
    free, _ = torch.cuda.mem_get_info(0)
    n_bg = 2_000_000_000 if free > (90 << 30) else 50_000_000
    n_t = 10000
    panel = synth.make_panel(n_t, seed=synth.PANEL_SEED)
    t = engine.Table.create(capacity=n_bg + len(panel.keys))
    t.build_synthetic(synth.TABLE_SEED, n_bg)
    t.insert(panel.keys, panel.counts, mode="overwrite")
    packed = engine.PackedTargets(panel.targets, panel.names)
    text, status = t.find_text(packed, "panel.jf")
    assert (status & ~np.uint32(16) == 0).all()
    blocks = {}
    for ln in text.split("\n"):
        if ln:
            blocks.setdefault(ln.split("\t")[1], []).append(ln)
    assert len(blocks) == n_t
    want_type = {"snv": "Substitution", "ins": ("Insertion", "ITD"), "del": "Deletion", "dup": ("ITD", "Insertion"), "none": "Reference"}
    missing = 0
    for name, truth in zip(panel.names, panel.truth):
        types = {ln.split("\t")[2] for ln in blocks[name]}
        w = want_type[truth["kind"]]
        if not (types & set([w] if isinstance(w, str) else w)):
            missing += 1
        assert "Reference" in types                          # every target prints its reference row
    assert missing <= n_t // 200, missing                    # a stray background hit can hide a variant; almost never
    # order / grouping invariance: a shuffled batch cut differently prints the same block per target
    rng = np.random.default_rng(0)
    perm = rng.permutation(n_t)[:3000]
    sub = engine.PackedTargets([panel.targets[i] for i in perm], [panel.names[i] for i in perm])
    text2, _ = t.find_text(sub, "panel.jf", n_sub=5)
    blocks2 = {}
    for ln in text2.split("\n"):
        if ln:
            blocks2.setdefault(ln.split("\t")[1], []).append(ln)
    assert all(blocks2[panel.names[i]] == blocks[panel.names[i]] for i in perm)
    # lookups: the walk issues the algorithmic count minus the reused reference successors, plus dead-end tips
    res = t.find_batch(engine.PackedTargets(panel.targets[:2000]), want_graph=False)
    n_ref = np.array([len(s) - 30 for s in panel.targets[:2000]])
    algorithmic = n_ref + 4 * (res.n_nodes.astype(np.int64) - 2)
    assert (res.lookups.astype(np.int64) >= n_ref).all()          # (the neighbour masks answer most absent successors without a read)
    # (a chain level asks for the sixteen grandchildren along with the four children: up to 20 per novel node)
    n_novel = res.n_nodes.astype(np.int64) - 2 - n_ref
    assert (res.lookups.astype(np.int64) <= algorithmic + 16 * n_novel + 4 * 64).all()


def test_table_counted_from_reads_equals_host_count(engine):
    """config 5's table build (C19: jellyfish count -C ... -L 2): canonical 31-mers of synthetic reads counted
    on the device, count < 2 dropped, against a host count of the same reads."""
    from collections import Counter
    rng = np.random.default_rng(12)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, size=5000))
    reads = []
    for _ in range(3000):
        s = int(rng.integers(0, 5000 - 100))
        r = genome[s:s + 100]
        if rng.random() < 0.5:
            r = r.translate(str.maketrans("ACGT", "TGCA"))[::-1]
        if rng.random() < 0.05:
            p = int(rng.integers(0, 100))
            r = r[:p] + "N" + r[p + 1:]
        reads.append(r)
    host = Counter()
    for r in reads:
        for i in range(len(r) - 30):
            km = r[i:i + 31]
            if "N" in km:
                continue
            v = jf_format.pack(km)
            host[min(v, jf_format.revcomp_packed(v, 31))] += 1
    t = engine.Table.create(capacity=4 * len(host) + 1024)
    t.count_reads(reads)
    assert t.info()["n_keys"] == len(host)
    left = t.drop_below(2)
    kept = {k: c for k, c in host.items() if c >= 2}
    assert left == len(kept)
    keys = np.array(list(host.keys()), dtype=np.uint64)
    got = t.query_packed(keys)
    want = np.array([host[int(k)] if host[int(k)] >= 2 else 0 for k in keys], dtype=np.uint32)
    assert (got == want).all()


def test_repeated_runs_print_the_same_text(engine):
    """The kernels race by design (atomics hand out node numbers, path slots and rows); the canonical
    numbering and the sorted output must hide it: the same batch, run again and again, prints the same text."""
    from km_b200 import synth
    panel = synth.make_panel(900, seed=41, two_variant_frac=0.3)
    t = engine.Table.create(capacity=len(panel.keys) + 300000)
    t.build_synthetic(synth.TABLE_SEED, 300000)
    t.insert(panel.keys, panel.counts, mode="overwrite")
    packed = engine.PackedTargets(panel.targets, panel.names)
    first, status = t.find_text(packed, "panel.jf")
    assert (status & ~np.uint32(16) == 0).all()
    for n_sub in (1, 4, 6, 2, 5, 3, 6, 1):
        again, _ = t.find_text(packed, "panel.jf", n_sub=n_sub)
        assert again == first
    plan = t.plan(panel.targets)
    for _ in range(5):
        plan.launch()
        assert plan.fetch(want_graph=False).format_all("panel.jf", packed) == first


@pytest.mark.gpu
def test_long_insertions_overflow_the_shared_memory_walk(engine):
    """An insertion of 120-300 bases is a chain of 150-330 novel k-mers: more than the shared-memory walk holds
    per target (KM_WS_NOVEL = 128), so its chain level raises the overflow flag half-way and the general walk
    kernel redoes the target.  Same records as the oracle."""
    from km_b200 import synth
    rng = np.random.default_rng(77)
    k = 31
    targets, names, keys, counts = [], [], [], []
    for i in range(12):
        L = int(rng.integers(150, 260))
        while True:
            ref = rng.integers(0, 4, size=L, dtype=np.uint8)
            ins = rng.integers(0, 4, size=int(rng.integers(120, 301)), dtype=np.uint8)
            pos = int(rng.integers(k, L - k + 1))
            alt = np.concatenate([ref[:pos], ins, ref[pos:]])
            ka, kr = synth.pack_kmers(alt, k), synth.pack_kmers(ref, k)
            if len(np.unique(ka)) == len(ka) and len(np.unique(kr)) == len(kr):
                break
        keys += [synth.canonical(kr, k), synth.canonical(ka, k)]
        counts += [np.full(len(kr), 300, dtype=np.int64), np.full(len(ka), 120, dtype=np.int64)]
        targets.append(synth.decode(ref))
        names.append("longins_%02d" % i)
    uk, inv = np.unique(np.concatenate(keys), return_inverse=True)
    uc = np.zeros(len(uk), dtype=np.int64)
    np.add.at(uc, inv, np.concatenate(counts))
    t = engine.Table.create(capacity=len(uk) + 100000)
    t.build_synthetic(synth.TABLE_SEED, 100000)
    t.insert(uk, uc.astype(np.uint32), mode="overwrite")
    res = t.find_batch(targets)
    assert (res.status & ~np.uint32(16) == 0).all()
    store = KmerStore(31, True, len(uk))
    store.set_background(synth.TABLE_SEED, 100000)
    store.insert(uk, uc.astype(np.uint32))
    jf = ko.OracleJellyfish(store, "long.jf", 0.05, 5)
    novel_max = 0
    for i, (name, seq) in enumerate(zip(names, targets)):
        f = ko.OracleFinder(ko.Target(seq, name, 31), jf).run()
        want = f.get_paths()
        got = record_of(res, i, "long.jf", name)
        errs, _ = compare_rows([str(r) for r in want], got["rows"],
                               [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in want], got["raw"])
        assert not errs, (name, errs)
        assert got["nodes"] == sorted([kk, int(v)] for kk, v in f.node_data.items())
        novel_max = max(novel_max, f.num_k - 2 - (len(seq) - 30))
        assert any(r.split("\t")[2] in ("Insertion", "ITD") for r in got["rows"]), name
    assert novel_max > 128
    # the one-call path prints the same rows
    text, _ = t.find_text(engine.PackedTargets(targets, names), "long.jf")
    assert text == "".join(res.format_target(i, "long.jf", names[i]) for i in range(len(targets)))


@pytest.mark.gpu
def test_text_call_survives_capacity_retries(engine, synth_small):
    """km_find_text with node capacities far too small at first: the fetch loop grows them and runs the
    sub-batch again (layout, upload, kernels AND the device formatter); the text is the same."""
    t = engine.Table.create(capacity=len(synth_small["keys"]) + 1024)
    t.insert(synth_small["keys"], synth_small["counts"].astype(np.uint32))
    packed = engine.PackedTargets(synth_small["targets"], synth_small["names"])
    want, _ = t.find_text(packed, "synth_small.jf")
    got, status = t.find_text(packed, "synth_small.jf", extra_nodes=2, n_sub=3)
    assert got == want and (status & ~np.uint32(16) == 0).all()
    assert t.last_timing["retries"] > 0


@pytest.mark.gpu
def test_find_cohort_equals_find_mutation_per_sample(engine, bundled):
    """`km find_cohort -t <catalog dir> <all .jf>`: the rows printed for every sample are the rows
    `km find_mutation <catalog dir> <that .jf>` prints, in the order of the databases on the command line."""
    import io
    import sys
    from argparse import Namespace
    from km_b200.tools import find_cohort as fc
    from km_b200.tools import find_mutation as fm
    cwd = os.getcwd()
    os.chdir(bundled)
    try:
        samples = sorted(f for f in os.listdir("./data/jf") if f.endswith(".jf"))
        assert len(samples) == 5
        catalog = "./data/catalog/GRCh38"

        def run(main, args):
            old = sys.stdout
            sys.stdout = buf = io.StringIO()
            try:
                main(args, None)
            finally:
                sys.stdout = old
            return [l for l in buf.getvalue().split("\n") if l and not l.startswith("#")]
        base = dict(count=5, ratio=0.05, steps=500, branchs=10, nodes=10000)
        cohort = run(fc.main_find_cohort, Namespace(target_fn=[catalog], jellyfish_fn=["./data/jf"], device=0, **base))
        want = [cohort[0]]
        for s in samples:
            rows = run(fm.main_find_mut, Namespace(target_fn=[catalog], jellyfish_fn="./data/jf/" + s, graphical=False,
                                                   verbose=False, debug=False, **base))
            assert rows[0] == cohort[0]
            want += rows[1:]
        assert cohort == want
        assert len(cohort) > 5 * 9
        # every table resident in HBM at once: the same rows
        resident = run(fc.main_find_cohort, Namespace(target_fn=[catalog], jellyfish_fn=["./data/jf"], device=0, resident=True, **base))
        assert resident == cohort
        # -f table: per target, what `km find_report -t <target> -f table` makes of ALL samples' rows (find_report.py:290-327)
        from km_b200.tools import find_report as fr
        table = run(fc.main_find_cohort, Namespace(target_fn=[catalog], jellyfish_fn=["./data/jf"], device=0, resident=True,
                                                   format="table", min_cov=1, **base))
        npm1 = "NPM1_4ins_exons_10-11utr"
        at = table.index("Target\t" + npm1)
        block = []
        for ln in table[at + 1:]:
            if ln.startswith("Target\t"):
                break
            block.append(ln)
        rows = [r for r in cohort[1:] if r.split("\t")[1] == npm1]
        want_tab = run(lambda a, _p: fr.create_report(a),
                       Namespace(target="./data/catalog/GRCh38/%s.fa" % npm1, infile=io.StringIO("".join(r + "\n" for r in rows)),
                                 info="vs_ref", min_cov=1, exclu="", format="table"))
        assert block == want_tab and block[0].startswith("Sample\t")
        # the NPM1 insertion is seen in the NPM1 sample only (the other samples hold none of this target's k-mers: their
        # Reference rows have Min_coverage 0 and find_report drops them under -m 1, find_report.py:141-142)
        assert len(block) == 2
        cells = {ln.split("\t")[0]: ln.split("\t")[1:] for ln in block[1:]}
        hdr = block[0].split("\t")[1:]
        ins_col = [i for i, h in enumerate(hdr) if "171410544" in h]
        assert ins_col and cells["./data/jf/02H025_NPM1.jf"][ins_col[0]] == "0.484"
        # and FLT3 shows its three samples side by side
        flt3 = table.index("Target\tFLT3-ITD_exons_13-15")
        assert sum(1 for ln in table[flt3 + 2:flt3 + 8] if ln.startswith("./data/jf/")) >= 2
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
def test_empty_and_single_target_batches(engine, synth_small):
    """Edge sizes of the one-call path: no target at all, one target, and more sub-batches asked for than
    there are targets."""
    t = engine.Table.create(capacity=len(synth_small["keys"]) + 1024)
    t.insert(synth_small["keys"], synth_small["counts"].astype(np.uint32))
    text, status = t.find_text(engine.PackedTargets([], []), "x.jf")
    assert text == "" and len(status) == 0
    res = t.find_batch([])
    assert len(res.status) == 0
    one = engine.PackedTargets(synth_small["targets"][:1], synth_small["names"][:1])
    want = t.find_batch(one.sequences).format_target(0, "x.jf", synth_small["names"][0])
    for n_sub in (0, 1, 6):
        got, status = t.find_text(one, "x.jf", n_sub=n_sub)
        assert got == want and len(status) == 1
    three = engine.PackedTargets(synth_small["targets"][:3], synth_small["names"][:3])
    a, _ = t.find_text(three, "x.jf", n_sub=1)
    b, _ = t.find_text(three, "x.jf", n_sub=8)
    assert a == b and a.count("\n") >= 3
