"""Drop-in for km/utils/Jellyfish.py: the same front end over a GPU-resident table.

``Jellyfish(filename, cutoff, n_cutoff)`` loads the .jf database into HBM
(km_table_open_jf); ``query`` and ``get_child`` keep the reference's per-k-mer signature and
run through the batched probe kernels (one-element batches), ``query_many`` /
``get_child_many`` expose the batched form.  No Jellyfish library, no CPU fallback.
"""
import numpy as np

from .. import engine


class Jellyfish:
    """A python interface for querying a k-mer count table held on the GPU.

    Methods
    -------
    query
    get_child
    """

    def __init__(self, filename, cutoff=0.30, n_cutoff=500, device=0, table=None):
        # Jellyfish.py:23-45 -- open the DB, learn k and `canonical` from its header
        self.jf = table if table is not None else engine.Table.open_jf(filename, device=device)
        self.k = self.jf.k
        self.filename = filename
        self.cutoff = cutoff
        self.n_cutoff = n_cutoff
        self.canonical = self.jf.canonical

    def query(self, seq):
        """Count of one k-mer, 0 when absent (Jellyfish.py:47-53)."""
        return int(self.jf.query_ascii([seq])[0])

    def query_many(self, seqs):
        """Counts of many k-mers in one launch."""
        return self.jf.query_ascii(list(seqs))

    def get_child(self, seq, forward=True):
        """Neighbouring k-mers whose count reaches max(sum * cutoff, n_cutoff), in A,C,G,T
        order (Jellyfish.py:55-72)."""
        return self.get_child_many([seq], forward)[0]

    def get_child_many(self, seqs, forward=True):
        seqs = list(seqs)
        packed = np.array([engine.pack_kmer(s) for s in seqs], dtype=np.uint64)
        _, mask = self.jf.get_child_packed(packed, self.cutoff, self.n_cutoff, forward)
        out = []
        for s, m in zip(seqs, mask.tolist()):
            if forward:
                out.append([s[1:] + b for i, b in enumerate("ACGT") if m >> i & 1])
            else:
                out.append([b + s[:-1] for i, b in enumerate("ACGT") if m >> i & 1])
        return out
