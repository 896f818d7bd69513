"""ORACLE / TEST INFRASTRUCTURE -- stand-in for the ``jellyfish`` Python binding.

Exposes exactly the four calls km makes (km/utils/Jellyfish.py:24-25,50-53):
``QueryMerFile(fn)``, ``qf[mer]``, ``MerDNA(str)``, ``MerDNA.k()``, ``.canonicalize()``.
Putting this directory and /root/reference on PYTHONPATH lets the reference's
UNMODIFIED Python run in this container (oracle/validate_vs_reference.py,
tests/golden/make_golden.py).  Backed by oracle.store.KmerStore.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import jf_format  # noqa: E402
from oracle.store import KmerStore  # noqa: E402

_K = [0]
# tests can register in-memory stores under a pseudo file name
REGISTRY = {}
QUERY_COUNT = [0]


class MerDNA:
    def __init__(self, s):
        self.s = s

    @staticmethod
    def k(*a):
        if a:
            _K[0] = int(a[0])
        return _K[0]

    def canonicalize(self):
        self.s = jf_format.canonical_str(self.s)

    def __str__(self):
        return self.s


class QueryMerFile:
    def __init__(self, fn):
        if fn in REGISTRY:
            self.store = REGISTRY[fn]
        else:
            self.store = KmerStore.from_jf(fn)
        _K[0] = self.store.k
        # MerDNA already canonicalised by km when the DB is canonical; do not do it twice
        self._raw = KmerStore.__new__(KmerStore)

    def __getitem__(self, mer):
        QUERY_COUNT[0] += 1
        s = mer.s
        # km canonicalises before the lookup iff header says canonical; the store would
        # canonicalise again, which is idempotent, so a plain query is exact either way.
        return self.store.query(s)
