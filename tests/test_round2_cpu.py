"""CPU-side checks of the round-2 test infrastructure and host logic (no GPU): the read generator of the cohort
config against the counting oracle, the reference runner against the oracle port, the quality mask of the counting
oracle."""
import os
import tempfile

import numpy as np
import pytest

from km_b200 import synth
from oracle import count_oracle, km_oracle as ko
from oracle.compare import compare_rows
from oracle.store import KmerStore


def test_sample_reads_count_to_the_panel_model():
    """config 5: counting the canonical 31-mers of the generated reads and dropping counts < 2 (`jellyfish count -C -L 2`,
    example/run_leucegene.sh:22) gives exactly the two-allele model of make_panel -- the reads ARE the sample."""
    panel = synth.make_panel(30, seed=11)
    stream = synth.sample_reads(panel)
    reads = stream.split(b"\n")[:-1]
    assert max(map(len, reads)) == synth.READ_LEN and min(map(len, reads)) >= 31
    keys, counts = count_oracle.count_stream(stream)
    keep = counts >= 2
    order = np.argsort(panel.keys)
    assert (keys[keep] == panel.keys[order]).all() and (counts[keep] == panel.counts[order]).all()
    # a subset of targets gives the subset's model
    sub = synth.make_panel(10, seed=11)
    assert sub.targets == panel.targets[:10]
    k2, c2 = count_oracle.count_stream(synth.sample_reads(panel, range(10)))
    o2 = np.argsort(sub.keys)
    assert (k2[c2 >= 2] == sub.keys[o2]).all() and (c2[c2 >= 2] == sub.counts[o2]).all()


def test_count_oracle_masks_low_quality_and_breaks_at_separators():
    seq = b"ACGTACGTACGTACGTACGTACGTACGTACGTACGT\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"
    k, c = count_oracle.count_stream(seq, k=31)
    assert int(c.sum()) == (36 - 30) + (35 - 30)                       # no k-mer spans the newline
    qual = bytearray(b"I" * len(seq))
    qual[10] = ord("#")                                                # one bad base in the first read
    k, c = count_oracle.count_stream(seq, k=31, qual=bytes(qual), min_qual=ord("+"))
    assert int(c.sum()) == 5                                           # every window of read 1 holds base 10


def test_reference_runner_equals_the_oracle_port():
    """The unmodified reference (baseline/_ref or /root/reference) driven the way bench.py drives it, against the
    oracle port on the same store: identical rows (skipped where no reference tree is present)."""
    from oracle import reference_runner as rr
    if rr.locate_reference() is None:
        pytest.skip("no reference tree (baseline/_ref is staged by tools/stage_reference.py)")
    panel = synth.make_panel(12, seed=3)
    store = KmerStore(31, True, len(panel.keys))
    store.set_background(synth.TABLE_SEED, 1_000_000)
    store.insert(panel.keys, panel.counts)
    with tempfile.TemporaryDirectory() as d:
        s = rr.ReferenceSession(store, d)
        rows, issued = s.find_mutation(s.write_targets(panel.names, panel.targets))
    jf = ko.OracleJellyfish(store, "panel.jf", 0.05, 5)
    port = []
    for n, seq in zip(panel.names, panel.targets):
        port += [str(r) for r in ko.OracleFinder(ko.Target(seq, n, 31), jf).run().get_paths()]
    errs, _ = compare_rows(rows, port)
    assert not errs and len(rows) >= 12 and issued > 0
    assert os.getcwd() != d
