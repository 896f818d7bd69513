"""TEST INFRASTRUCTURE ONLY -- drives tests/emu/km_emu.cpp (the single-lane host build of the
device stage functions) and returns results in the shape of km_b200.engine.BatchResult so the
product's own host code (row spelling, Path objects, sorting) is exercised on top of it."""
import ctypes
import os
import subprocess

import numpy as np

from km_b200 import engine
from km_b200._lib import ROW_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "km_emu.cpp")
SO = os.environ.get("KM_EMU_SO") or os.path.join(HERE, "emu", "libkm_emu_test_asan.so" if os.environ.get("KM_EMU_SANITIZE") else "libkm_emu_test.so")
CSRC = os.path.join(os.path.dirname(HERE), "km_b200", "csrc")
_L = None


def lib():
    global _L
    if _L is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            # KM_EMU_SANITIZE=1: the same stage functions under AddressSanitizer + UBSan (run pytest with
            # LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0); compute-sanitizer is closed on
            # this GPU pool, so this is where out-of-bounds scratch / shared-memory indexing gets caught
            extra = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g", "-O1"] if os.environ.get("KM_EMU_SANITIZE") else ["-O2"]
            subprocess.check_call(["g++", *extra, "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-o", SO, SRC])
        L = ctypes.CDLL(SO)
        vp, u64, u32, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
        L.emu_table_create.restype = vp
        L.emu_table_create.argtypes = [ci, ci, u64]
        L.emu_table_free.argtypes = [vp]
        L.emu_table_insert.argtypes = [vp, vp, vp, u64, ci]
        L.emu_table_synthetic.argtypes = [vp, u64, u64]
        L.emu_query.restype = u32
        L.emu_query.argtypes = [vp, u64]
        L.emu_revcomp.restype = u64
        L.emu_revcomp.argtypes = [u64, ci]
        L.emu_synth_key.restype = u64
        L.emu_synth_key.argtypes = [u64, u64, ci]
        L.emu_synth_count.restype = u32
        L.emu_synth_count.argtypes = [u64]
        L.emu_find_target.restype = ci
        L.emu_find_target.argtypes = [vp, vp, ci, ctypes.c_double, ctypes.c_int64, ci, ci, ci, ci,
                                      vp, vp, vp, ci, vp, vp, vp, ci, ci, vp, vp, ci, vp, ci, ci, ci, ci, vp]
        assert L.emu_row_size() == ROW_DTYPE.itemsize
        _L = L
    return _L


class EmuTable:
    def __init__(self, k=31, canonical=True, capacity=4096):
        self.k, self.canonical = k, canonical
        self._h = lib().emu_table_create(k, int(canonical), int(capacity))

    @classmethod
    def from_keys(cls, keys, counts, k=31, canonical=True):
        t = cls(k, canonical, max(1024, len(keys)))
        t.insert(keys, counts)
        return t

    def insert(self, keys, counts, mode=1):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        assert lib().emu_table_insert(self._h, keys.ctypes.data, counts.ctypes.data, keys.size, mode) == 0

    def build_synthetic(self, seed, n):
        assert lib().emu_table_synthetic(self._h, int(seed), int(n)) == 0

    def query_packed(self, v):
        return int(lib().emu_query(self._h, int(v)))

    def find_batch(self, sequences, count=5, ratio=0.05, steps=500, branchs=10, nodes=10000, extra_nodes=256,
                   small=(512, 512, 64, 8)):
        """small = (nodes, candidate edges, paths, columns) capacities of the shared-memory pass."""
        n = len(sequences)
        res = engine.BatchResult()
        res.k = self.k
        res.sequences = list(sequences)
        status, n_nodes, lookups = [], [], []
        self.passes = []
        node_off = [0]
        kmers, cnts, pfirst, pcount, plen, poff, pool, rfirst, rcount, rows = [], [], [], [], [], [], [], [], [], []
        for t, seq in enumerate(sequences):
            codes = engine._CODE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)].copy()
            extra = extra_nodes
            while True:
                cap = max(1, len(seq) - self.k + 1) + extra
                out_n = ctypes.c_int32()
                o_k = np.zeros(cap, np.uint64)
                o_c = np.zeros(cap, np.uint32)
                npaths, nrows = ctypes.c_int32(), ctypes.c_int32()
                pl = np.zeros(1024, np.int32)
                po = np.zeros(1 << 20, np.int32)
                rw = np.zeros(4096, dtype=ROW_DTYPE)
                lk = ctypes.c_uint64()
                which = ctypes.c_int32()
                st = lib().emu_find_target(self._h, codes.ctypes.data, len(seq), float(ratio), int(count), int(steps),
                                           int(branchs), int(nodes), extra, ctypes.byref(out_n), o_k.ctypes.data,
                                           o_c.ctypes.data, cap, ctypes.byref(npaths), pl.ctypes.data, po.ctypes.data,
                                           1024, 1 << 20, ctypes.byref(nrows), rw.ctypes.data, 4096, ctypes.byref(lk),
                                           small[0], small[1], small[2], small[3], ctypes.byref(which))
                assert st >= 0
                if st & engine.ST_NODE_OVERFLOW:
                    extra *= 8
                    continue
                break
            status.append(st)
            self.passes.append(which.value)
            n_nodes.append(out_n.value)
            lookups.append(lk.value)
            nn = max(0, out_n.value - 2)
            kmers.append(o_k[:nn])
            cnts.append(o_c[:nn])
            node_off.append(node_off[-1] + nn)
            pfirst.append(len(plen))
            pcount.append(npaths.value)
            base = sum(len(x) for x in pool)
            at = 0
            for p in range(npaths.value):
                poff.append(base + at)
                plen.append(int(pl[p]))
                at += int(pl[p])
            pool.append(po[:at].copy())
            rfirst.append(sum(len(x) for x in rows))
            rcount.append(nrows.value)
            r = rw[:nrows.value].copy()
            r["path_id"] += pfirst[-1]
            r["target"] = t
            rows.append(r)
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        res.status = np.array(status, np.uint32)
        res.n_nodes = np.array(n_nodes, np.int32)
        res.node_off = np.array(node_off, np.int64)
        res.node_kmer = cat(kmers, np.uint64)
        res.node_count = cat(cnts, np.uint32)
        res.path_first = np.array(pfirst, np.int32)
        res.path_count = np.array(pcount, np.int32)
        res.path_off = np.array(poff, np.int64)
        res.path_len = np.array(plen, np.int32)
        res.path_pool = cat(pool, np.int32)
        res.row_first = np.array(rfirst, np.int32)
        res.row_count = np.array(rcount, np.int32)
        res.rows = np.concatenate(rows) if rows else np.zeros(0, ROW_DTYPE)
        res.lookups = np.array(lookups, np.uint64)
        res.timing = {}
        return res

    def close(self):
        if self._h:
            lib().emu_table_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
