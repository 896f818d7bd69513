"""ORACLE / TEST INFRASTRUCTURE -- Python face of oracle/kmer_store.c.

``KmerStore`` plays the role of Jellyfish's ``QueryMerFile`` for the oracle
(km/utils/Jellyfish.py:24,53): canonical k-mer -> count, 0 when absent.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference)
may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import jf_format

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libkmer_store.so")
    src = os.path.join(_HERE, "kmer_store.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libkmer_store.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u64, u32, vp, ci = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int
        L.ks_create.restype = vp
        L.ks_create.argtypes = [ci, ci, u64]
        L.ks_destroy.argtypes = [vp]
        L.ks_insert.restype = ci
        L.ks_insert.argtypes = [vp, vp, vp, u64, ci]
        L.ks_set_background.argtypes = [vp, u64, u64]
        L.ks_size.restype = u64
        L.ks_size.argtypes = [vp]
        L.ks_query_packed.restype = u32
        L.ks_query_packed.argtypes = [vp, u64]
        L.ks_query_ascii.restype = u32
        L.ks_query_ascii.argtypes = [vp, ctypes.c_char_p]
        L.ks_query_batch.argtypes = [vp, vp, u64, vp]
        L.ks_revcomp.restype = u64
        L.ks_revcomp.argtypes = [u64, ci]
        L.ks_synth_key.restype = u64
        L.ks_synth_key.argtypes = [u64, u64, ci]
        L.ks_synth_raw.restype = u64
        L.ks_synth_raw.argtypes = [u64, u64]
        L.ks_synth_count.restype = u32
        L.ks_synth_count.argtypes = [u64]
        _LIB = L
    return _LIB


class KmerStore:
    """CPU k-mer -> count map with Jellyfish's missing -> 0 convention."""

    def __init__(self, k=31, canonical=True, capacity_hint=0):
        self.k = int(k)
        self.canonical = bool(canonical)
        self._h = lib().ks_create(self.k, int(self.canonical), int(capacity_hint))
        if not self._h:
            raise MemoryError("ks_create failed")

    @classmethod
    def from_jf(cls, path):
        header, keys, counts = jf_format.read_jf(path)
        s = cls(k=int(header["key_len"]) // 2, canonical=bool(header["canonical"]),
                capacity_hint=len(keys))
        s.header = header
        s.insert(keys, counts)
        return s

    def insert(self, keys, counts, overwrite=True):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        assert keys.shape == counts.shape
        rc = lib().ks_insert(self._h, keys.ctypes.data, counts.ctypes.data, keys.size, int(overwrite))
        if rc:
            raise MemoryError("ks_insert failed")

    def set_background(self, seed, n):
        lib().ks_set_background(self._h, int(seed), int(n))

    def __len__(self):
        return int(lib().ks_size(self._h))

    def query(self, seq):
        """ASCII k-mer -> count (canonicalised when the store is canonical)."""
        if len(seq) != self.k:
            raise ValueError("k-mer length %d != k=%d" % (len(seq), self.k))
        c = lib().ks_query_ascii(self._h, seq.encode("ascii"))
        if c == 0xFFFFFFFF:
            raise ValueError("non-ACGT k-mer %r (parity unpinned, SURVEY.md 8c)" % seq)
        return int(c)

    def query_packed(self, v):
        return int(lib().ks_query_packed(self._h, int(v)))

    def query_batch(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(kmers.shape, dtype=np.uint32)
        lib().ks_query_batch(self._h, kmers.ctypes.data, kmers.size, out.ctypes.data)
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().ks_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
