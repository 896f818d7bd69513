"""Drop-in for km/utils/MutationFinder.py.

The reference walks, builds the graph and quantifies one target at a time in Python; here
all three stages run on the GPU for a whole batch of targets (km_find_batch) and this class
is a per-target VIEW of that result with the reference's attribute and method names.
``MutationFinder(refpath, jf, ...)`` on its own runs a batch of one; ``find_batch`` is what
the find_mutation tool uses so that every target of a run shares one launch.
"""
import logging as log
import sys
from collections import namedtuple

from .. import engine
from . import PathQuant as upq
from . import Sequence as us
from . import common as uc

PathDiff = namedtuple("PathDiff", ["start", "end_ref", "end_var", "kmers_ref", "kmers_var", "end_ref_overlap"])


def find_batch(refpaths, jf, max_stack=500, max_break=10, max_node=10000):
    """One km_find_batch call for all targets -> list of MutationFinder, in order.  A finder
    whose target failed raises when it is materialised (`finder.check()` / any method), so
    rows of earlier targets can still be printed first."""
    result = jf.jf.find_batch([r.seq for r in refpaths], count=jf.n_cutoff, ratio=jf.cutoff,
                              steps=max_stack, branchs=max_break, nodes=max_node)
    return [MutationFinder(r, jf, max_stack, max_break, max_node, _batch=(result, i))
            for i, r in enumerate(refpaths)]


class MutationFinder:
    """Per-target results of the GPU find_mutation pipeline.

    Attributes (as in the reference): refpath, jf, max_stack, max_break, max_node, first_seq,
    last_seq, ref_set, node_data, kmer, counts, num_k, first_seq_ix, last_seq_ix, start_kmers,
    end_kmers, start_kmers_ix, end_kmers_ix, alt_paths, alt_groups, paths.

    Node numbering is canonical: reference k-mers in sequence order, novel k-mers by
    ascending 2-bit value, then BigBang and BigCrunch (the reference's numbering follows
    Python set iteration and is not reproducible).
    """

    def __init__(self, refpath, jf, max_stack=500, max_break=10, max_node=10000, _batch=None):
        self.refpath = refpath
        self.first_seq = "BigBang"
        self.last_seq = "BigCrunch"
        self.ref_set = set(refpath.ref_mer)
        self.jf = jf
        self.max_stack = max_stack
        self.max_break = max_break
        self.max_node = max_node
        self.paths = []
        if _batch is None:
            self._res = jf.jf.find_batch([refpath.seq], count=jf.n_cutoff, ratio=jf.cutoff, steps=max_stack,
                                         branchs=max_break, nodes=max_node)
            self._t = 0
            self.check()
            self._load_nodes()
        else:
            self._res, self._t = _batch
            self._loaded = False

    # ---- materialisation ---------------------------------------------------------------
    def check(self):
        """Raise what the reference would raise for this target (node limit -> sys.exit)."""
        engine.raise_for_status(self._res.status[self._t], self.refpath.name, self.max_node)

    @property
    def touched_limit(self):
        """True when some walk hit max_stack/max_break: the reference's own answer then
        depends on Python's set iteration order."""
        return bool(int(self._res.status[self._t]) & engine.ST_TOUCHED_LIMIT)

    def _load_nodes(self):
        res, t = self._res, self._t
        self.kmer = res.kmers(t)
        self.counts = res.counts(t)
        self.num_k = len(self.kmer)
        self.node_data = dict(zip(self.kmer[:-2], self.counts[:-2]))
        log.info("Ref. set contains %d kmers.", len(self.ref_set))
        log.info("k-mer graph contains %d nodes.", self.num_k)
        self.refpath.set_index(self.kmer)
        self.first_seq_ix = self.num_k - 2
        self.last_seq_ix = self.num_k - 1
        self.start_kmers = {self.refpath.first_kmer}
        self.end_kmers = {self.refpath.last_kmer}
        self.start_kmers_ix = {self.kmer.index(k) for k in self.start_kmers}
        self.end_kmers_ix = {self.kmer.index(k) for k in self.end_kmers}
        log.info("BigBang=%d, BigCrunch=%d" % (self.first_seq_ix, self.last_seq_ix))
        self._loaded = True

    def _ensure(self):
        if not getattr(self, "_loaded", False):
            self.check()
            self._load_nodes()

    # ---- the four stages of tools/find_mutation.py:49-57 ---------------------------------
    def graph_analysis(self):
        """Alternative paths found by the GPU graph stage (MutationFinder.py:496-572)."""
        self._ensure()
        self.paths = []
        self.alt_paths = [us.AltSeq(p, self) for p in self._res.paths(self._t)]
        groups = {}
        for path in self.alt_paths:
            groups.setdefault(path.ref_name, []).append(path)
        self.alt_groups = groups

    def _rows(self, kind):
        fields = self._res.row_fields(self._t, self.jf.filename, self.refpath.name)
        return [upq.Path(*f) for f in fields if (f[11] == "vs_ref") == (kind == "vs_ref")]

    def quantify_paths(self, graphical=False):
        """vs_ref rows (MutationFinder.py:575-648).  `graphical` needs matplotlib and the
        reference's interactive plots; not supported here."""
        self._ensure()
        if graphical:
            raise NotImplementedError("-g/--graphical plots are outside the GPU path")
        self.paths.extend(self._rows("vs_ref"))

    def quantify_clusters(self, graphical=False):
        """cluster rows (MutationFinder.py:726-811)."""
        self._ensure()
        if graphical:
            raise NotImplementedError("-g/--graphical plots are outside the GPU path")
        self.paths.extend(self._rows("cluster"))

    def get_paths(self, sort=True):
        """Rows, sorted like the reference (MutationFinder.py:813-833)."""
        if not sort:
            return self.paths
        return sorted(self.paths,
                      key=lambda x: uc.natsortkey(*x[11].split(" "), x[1], x[3], x[2], x[6], rev_ix=[0]))

    def format_rows(self):
        """Sorted rows as the TSV text find_mutation prints, formatted by the library."""
        self.check()
        return self._res.format_target(self._t, self.jf.filename, self.refpath.name)

    # ---- helpers kept for API compatibility (host-side, cheap) -----------------------------
    @staticmethod
    def diff_path_without_overlap(ref, seq, k):
        """Same contract as MutationFinder.py:190-373 (the GPU computes this per row; this
        host version serves callers that pass their own index lists)."""
        n_ref, n_seq = len(ref), len(seq)
        i = 0
        while i < n_ref and i < n_seq and ref[i] == seq[i]:
            i += 1
        j_ref, j_seq = n_ref, n_seq
        while j_ref >= i + k and j_seq >= i + k and ref[j_ref - 1] == seq[j_seq - 1]:
            j_ref -= 1
            j_seq -= 1
        k_ref, k_seq = j_ref, j_seq
        while k_ref > i and ref[k_ref - 1] == seq[k_seq - 1]:
            k_ref -= 1
            k_seq -= 1
        return PathDiff(i, j_ref, j_seq, ref[i:j_ref], seq[i:j_seq], k_ref)

    def get_seq(self, path, skip_prefix=True):
        """Spell a list of node indices (MutationFinder.py:375-403)."""
        self._ensure()
        if not path:
            return ""
        head = self.kmer[path[0]]
        return (head[-1] if skip_prefix else head) + "".join(self.kmer[i][-1] for i in path[1:])

    def get_name(self, ref_ix, path_ix, offset=0):
        """Type and position string of one path against a reference path
        (MutationFinder.py:405-488)."""
        self._ensure()
        k = self.jf.k
        diff = self.diff_path_without_overlap(ref_ix, path_ix, k)
        if len(ref_ix) - len(diff.kmers_ref) + len(diff.kmers_var) != len(path_ix):
            sys.stderr.write("ERROR: %s %d != %d" % ("mutation identification could be incorrect",
                                                     len(ref_ix) - len(diff.kmers_ref) + len(diff.kmers_var),
                                                     len(path_ix)))
            raise Exception()
        del_seq = self.get_seq(diff.kmers_ref, skip_prefix=True)
        ins_seq = self.get_seq(diff.kmers_var, skip_prefix=True)
        shared = 0
        if del_seq:
            assert del_seq != ins_seq
            while shared < len(del_seq) and shared < len(ins_seq) and del_seq[-1 - shared] == ins_seq[-1 - shared]:
                shared += 1
        if shared:
            del_seq, ins_seq = del_seq[:-shared], ins_seq[:-shared]
        if diff.end_ref == diff.end_var:
            variant = "Reference" if diff.start == diff.end_ref else "Substitution"
        elif diff.start == diff.end_ref_overlap:
            variant = "ITD"
        elif diff.end_ref < diff.end_var:
            variant = "Insertion" if not del_seq else "Indel"
        else:
            variant = "Deletion" if not ins_seq else "Indel"
        if variant == "Reference":
            return variant + "\t"
        return "{}\t{}:{}:{}".format(variant, diff.start + k + offset, del_seq.lower() + "/" + ins_seq,
                                     diff.end_ref + 1 + offset)

    def get_counts(self, path):
        self._ensure()
        return [self.node_data[self.kmer[i]] for i in path]

    @staticmethod
    def output_header():
        upq.Path.output_header()
