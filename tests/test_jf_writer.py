"""The binary/sorted WRITER: what order are the records of a real Jellyfish file in, and does
km_table_write_jf produce a file of that kind?  The CPU test pins the order on the five files bundled with
the reference (they were written by jellyfish 2.2.3); the GPU tests round-trip tables through the writer."""
import json
import os

import numpy as np
import pytest

from oracle import jf_format


def matrix_positions(header, keys):
    """pos = M * key over GF(2), bit b of the key selecting column c-1-b of matrix1, masked to `size`."""
    m = header["matrix1"]
    cols = [int(c) for c in m["columns"]]
    c = int(m["c"])
    pos = np.zeros(len(keys), dtype=np.uint64)
    for b in range(c):
        bit = (keys >> np.uint64(b)) & np.uint64(1)
        pos ^= bit * np.uint64(cols[c - 1 - b])
    return pos & np.uint64(int(header["size"]) - 1)


def test_bundled_files_are_sorted_by_their_matrix(bundled):
    d = os.path.join(bundled, "data", "jf")
    names = sorted(f for f in os.listdir(d) if f.endswith(".jf"))
    assert len(names) == 5
    for name in names:
        header, keys, counts = jf_format.read_jf(os.path.join(d, name))
        pos = matrix_positions(header, keys).astype(np.int64)
        assert (np.diff(pos) >= 0).all(), name
        # the other bit order does NOT sort them: the finding is not vacuous
        cols = [int(c) for c in header["matrix1"]["columns"]]
        alt = np.zeros(len(keys), dtype=np.uint64)
        for b in range(62):
            alt ^= ((keys >> np.uint64(b)) & np.uint64(1)) * np.uint64(cols[b])
        assert (np.diff((alt & np.uint64(int(header["size"]) - 1)).astype(np.int64)) < 0).any(), name
        assert header["reprobes"][:5] == [1, 1, 3, 6, 10] and len(header["reprobes"]) == 127


@pytest.fixture(scope="module")
def engine():
    from km_b200 import engine as e
    from km_b200._lib import lib
    if lib().km_device_count() <= 0:
        pytest.skip("no CUDA device")
    return e


@pytest.mark.gpu
def test_bundled_table_round_trips_through_the_writer(engine, bundled, tmp_path):
    from km_b200.utils.Jellyfish import Jellyfish
    for name in ("02H025_NPM1", "03H116_ITD"):
        src = os.path.join(bundled, "data", "jf", name + ".jf")
        header, keys, counts = jf_format.read_jf(src)
        t = engine.Table.open_jf(src)
        ek, ec = t.export()
        order = np.argsort(ek)
        ref = np.argsort(keys)
        assert (ek[order] == keys[ref]).all() and (ec[order] == counts[ref].astype(np.uint32)).all()
        out = str(tmp_path / (name + "_copy.jf"))
        t.write_jf(out)
        h2, k2, c2 = jf_format.read_jf(out)
        assert h2["format"] == "binary/sorted" and h2["key_len"] == 62 and h2["counter_len"] == 4 and h2["canonical"] is True
        assert (9 + int(open(out, "rb").read(9))) % 8 == 0                     # records start 8-byte aligned
        assert (np.diff(matrix_positions(h2, k2).astype(np.int64)) >= 0).all()
        assert (np.sort(k2) == np.sort(keys)).all()
        assert dict(zip(k2.tolist(), c2.tolist())) == dict(zip(keys.tolist(), counts.tolist()))
        # the product's own reader and km's wrapper take the copy for the original
        jf = Jellyfish(out)
        assert jf.k == 31 and jf.canonical
        some = [jf_format.unpack(int(v), 31) for v in keys[:50]]
        assert [jf.query(s) for s in some] == [int(c) for c in counts[:50]]
        assert jf.query(jf_format.revcomp(some[0])) == int(counts[0])


@pytest.mark.gpu
def test_counted_reads_written_and_read_back(engine, tmp_path):
    """reads -> device count (-C) -> drop below 2 (-L 2) -> .jf -> open: the pipeline that replaces
    `jellyfish count` in km's workflow (run_leucegene.sh:22)."""
    from collections import Counter
    rng = np.random.default_rng(5)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, size=3000))
    reads = [genome[s:s + 80] for s in rng.integers(0, 3000 - 80, size=2000).tolist()]
    host = Counter()
    for r in reads:
        for i in range(len(r) - 30):
            v = jf_format.pack(r[i:i + 31])
            host[min(v, jf_format.revcomp_packed(v, 31))] += 1
    t = engine.Table.create(capacity=4 * len(host) + 1024)
    t.count_reads(reads)
    t.drop_below(2)
    out = str(tmp_path / "reads.jf")
    t.write_jf(out, counter_len=2)
    h, k2, c2 = jf_format.read_jf(out)
    want = {k: min(v, 65535) for k, v in host.items() if v >= 2}
    assert h["counter_len"] == 2 and dict(zip(k2.tolist(), c2.tolist())) == want
    t2 = engine.Table.open_jf(out)
    q = np.array(list(want.keys())[:500], dtype=np.uint64)
    assert (t2.query_packed(q) == np.array([want[int(v)] for v in q], dtype=np.uint32)).all()


def _write_fastq(path, reads, quals, gz=False):
    import gzip
    text = "".join("@r%d\n%s\n+\n%s\n" % (i, r, q) for i, (r, q) in enumerate(zip(reads, quals)))
    with (gzip.open(path, "wt") if gz else open(path, "w")) as f:
        f.write(text)


def test_read_sequences_fasta_fastq_quality_mask(tmp_path):
    from km_b200.tools.count import parse_size, read_sequences
    fa = tmp_path / "x.fa"
    fa.write_text(">a desc\nACGT\nACGT\n>b\nTTTT\n")
    assert list(read_sequences(str(fa))) == [b"ACGTACGT", b"TTTT"]
    fq = str(tmp_path / "x.fq.gz")
    _write_fastq(fq, ["ACGTAC", "GGGG"], ["II*III", "++++"], gz=True)
    assert list(read_sequences(fq)) == [b"ACGTAC", b"GGGG"]
    assert list(read_sequences(fq, min_qual=ord("+"))) == [b"ACNTAC", b"GGGG"]       # '*' < '+' is masked, '+' is kept
    assert parse_size("3G") == 3_000_000_000 and parse_size("100M") == 100_000_000 and parse_size("1234") == 1234


@pytest.mark.gpu
def test_km_count_cli_equals_host_count(engine, tmp_path):
    """`km count -m 31 -C -L 2 -Q +` on a FASTQ with low-quality bases == a host count of the masked reads."""
    import subprocess
    import sys
    from collections import Counter
    rng = np.random.default_rng(11)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, size=2000))
    reads, quals = [], []
    for s in rng.integers(0, 2000 - 90, size=1500).tolist():
        r = genome[s:s + 90]
        if rng.random() < 0.5:
            r = jf_format.revcomp(r)
        q = ["I"] * 90
        for p in rng.integers(0, 90, size=int(rng.integers(0, 3))).tolist():
            q[p] = "#"
        reads.append(r)
        quals.append("".join(q))
    fq = str(tmp_path / "reads.fastq")
    _write_fastq(fq, reads, quals)
    host = Counter()
    for r, q in zip(reads, quals):
        masked = "".join(b if c >= "+" else "N" for b, c in zip(r, q))
        for i in range(len(masked) - 30):
            w = masked[i:i + 31]
            if "N" not in w:
                v = jf_format.pack(w)
                host[min(v, jf_format.revcomp_packed(v, 31))] += 1
    out = str(tmp_path / "reads.jf")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, "-m", "km_b200", "count", "-m", "31", "-s", "1M", "-C", "-L", "2", "-Q", "+",
                           "-t", "8", "-o", out, fq], cwd=root)
    h, k2, c2 = jf_format.read_jf(out)
    assert dict(zip(k2.tolist(), c2.tolist())) == {k: v for k, v in host.items() if v >= 2}
    assert h["canonical"] is True and h["key_len"] == 62


@pytest.mark.gpu
def test_native_reader_equals_python_reader(engine, tmp_path):
    """km_table_count_file (FASTA with wrapped lines and CRLF, FASTQ, .gz, quality mask) == the same sequences
    parsed by tools/count.read_sequences and counted through count_reads."""
    from km_b200.tools.count import count_into
    rng = np.random.default_rng(3)
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=int(n))) for n in rng.integers(40, 400, size=60)]
    fa = tmp_path / "wrapped.fa"
    with open(fa, "w", newline="") as f:
        for i, s in enumerate(seqs):
            f.write(">s%d some text\r\n" % i)
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + "\r\n")
            f.write("\r\n" if i % 7 == 0 else "")
    quals = ["".join(rng.choice(list("#+5I"), size=len(s), p=[0.01, 0.01, 0.01, 0.97])) for s in seqs]
    fq = str(tmp_path / "r.fq.gz")
    _write_fastq(fq, seqs, quals, gz=True)
    for files, q in (([str(fa)], None), ([fq], None), ([fq], ord("+")), ([str(fa), fq], ord("5"))):
        a = engine.Table.create(capacity=1 << 16)
        b = engine.Table.create(capacity=1 << 16)
        ra = count_into(a, files, q, native=True)
        rb = count_into(b, files, q, native=False)
        assert ra == rb and ra[0] == 60 * len(files)
        ka, ca = a.export()
        kb, cb = b.export()
        oa, ob = np.argsort(ka), np.argsort(kb)
        assert (ka[oa] == kb[ob]).all() and (ca[oa] == cb[ob]).all() and len(ka) > 500
