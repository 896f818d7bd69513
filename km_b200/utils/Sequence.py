"""Reference / alternative sequence holders -- host mirror of km/utils/Sequence.py."""
import sys

from . import common as uc


class RefSeq:
    """Target sequence, its name and its k-mers (Sequence.py:10-60).  A repeated k-mer
    raises ValueError from get_ref_kmer, exactly where the reference raises it."""

    def __init__(self, seq, name, k):
        self.seq = seq
        self.name = name
        self.k = k
        self.first_kmer = seq[:k]
        self.last_kmer = seq[-k:]
        self.ref_mer = uc.get_ref_kmer(seq, name, k)
        assert len(self.ref_mer)

    def set_index(self, kmer):
        """Node index of every reference k-mer in the finder's node list (Sequence.py:48-53)."""
        where = {}
        for ix, km in enumerate(kmer):
            where.setdefault(km, ix)
        self.seq_index = tuple(where[km] for km in self.ref_mer)
        self.first_ix = self.seq_index[0]
        self.last_ix = self.seq_index[-1]

    def __getitem__(self, item):
        if not hasattr(self, "seq_index"):
            sys.stderr.write("Attribute `seq_index` is not set yet\n")
            return None
        return self.seq_index[item]


class AltSeq:
    """One alternative path of a finder (Sequence.py:63-82)."""

    def __init__(self, alt_index, finder):
        self.finder = finder
        self.seq_index = alt_index
        self.seq = finder.get_seq(alt_index, skip_prefix=False)
        self.first_ix = alt_index[0]
        self.last_ix = alt_index[-1]
        self.seq_len = len(alt_index)
        self.ref_index = finder.refpath.seq_index
        self.ref_name = finder.refpath.name

    def __getitem__(self, item):
        return self.seq_index[item]
