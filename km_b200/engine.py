"""Host layer over the C ABI: the GPU k-mer table and the batched find_mutation engine.

Everything numeric happens in libkm_b200.so; this module moves buffers and turns the flat
result arrays into the objects km's Python API exposes (km/utils/MutationFinder.py,
km/utils/PathQuant.py).
"""
import ctypes
import os
import sys

import numpy as np

from . import _lib
from ._lib import FindParams, ResultView, ROW_DTYPE, TableInfo, check, lib

BASES = "ACGT"
TYPE_NAMES = ("Reference", "Substitution", "ITD", "Indel", "Insertion", "Deletion")
CAP_SRC, CAP_SNK = "BigBang", "BigCrunch"     # MutationFinder.py:97-98

ST_BAD_BASE, ST_DUP_KMER, ST_NODE_OVERFLOW, ST_NODE_LIMIT, ST_TOUCHED_LIMIT = 1, 2, 4, 8, 16
ST_PATH_OVERFLOW, ST_TOO_SHORT, ST_TOO_MANY_COLS, ST_SOLVER_WATCHDOG, ST_NAME_MISMATCH = 32, 64, 128, 256, 512

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


def pack_kmer(seq):
    """ASCII k-mer -> 2-bit packed int, first base most significant."""
    codes = _CODE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]
    if (codes > 3).any():
        raise ValueError("k-mer %r holds a letter outside ACGT" % seq)
    v = 0
    for c in codes.tolist():
        v = (v << 2) | c
    return v


def unpack_kmer(v, k):
    v = int(v)
    return "".join(BASES[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


class PackedTargets:
    """A batch of targets as the C ABI takes it: the sequences concatenated in one host buffer plus
    offsets (and, optionally, the query names the same way).  Build it once when the same batch is
    submitted repeatedly; `Table.find_batch` and `BatchResult.format_all` accept it in place of lists."""

    def __init__(self, sequences, names=None):
        self.sequences = list(sequences)
        self.blob, self.off = Table._pack_targets(self.sequences)
        self.names = list(names) if names is not None else None
        if self.names is not None:
            self.name_blob, self.name_off = _pack_names(self.names)

    def __len__(self):
        return len(self.sequences)

    def slice(self, lo, hi):
        """Targets [lo, hi) as a PackedTargets of their own (buffers cut, not re-encoded)."""
        sub = PackedTargets.__new__(PackedTargets)
        sub.sequences = self.sequences[lo:hi]
        sub.blob = self.blob[int(self.off[lo]):int(self.off[hi])]
        sub.off = (self.off[lo:hi + 1] - self.off[lo]).copy()
        sub.names = self.names[lo:hi] if self.names is not None else None
        if self.names is not None:
            sub.name_blob = self.name_blob[int(self.name_off[lo]):int(self.name_off[hi])]
            sub.name_off = (self.name_off[lo:hi + 1] - self.name_off[lo]).copy()
        return sub


def _pack_names(names):
    blob = "".join(names).encode()
    off = np.zeros(len(names) + 1, dtype=np.int64)
    if names:
        np.cumsum(np.fromiter(map(len, names), dtype=np.int64, count=len(names)), out=off[1:])
        if off[-1] != len(blob):          # non-ASCII names: fall back to byte lengths
            np.cumsum([len(x.encode()) for x in names], out=off[1:])
    return blob, off


class Table:
    """Device-resident k-mer count table (replaces jellyfish.QueryMerFile)."""

    def __init__(self, handle):
        self._h = handle
        info = TableInfo()
        check(lib().km_table_get_info(self._h, ctypes.byref(info)))
        self.k = int(info.k)
        self.canonical = bool(info.canonical)
        self.device = int(info.device)

    @classmethod
    def open_jf(cls, path, device=0):
        h = ctypes.c_void_p()
        check(lib().km_table_open_jf(str(path).encode(), int(device), ctypes.byref(h)))
        return cls(h)

    @classmethod
    def create(cls, k=31, canonical=True, capacity=1 << 20, device=0, layout=None):
        """layout: None = library default (sector buckets, or family lines when KM_TABLE_LINES is set),
        0 = sector buckets, 1 = family lines (include/km_b200.h)."""
        h = ctypes.c_void_p()
        if layout is None:
            check(lib().km_table_create(int(device), int(k), int(bool(canonical)), int(capacity), ctypes.byref(h)))
        else:
            check(lib().km_table_create_layout(int(device), int(k), int(bool(canonical)), int(capacity), int(layout),
                                               ctypes.byref(h)))
        return cls(h)

    def info(self):
        info = TableInfo()
        check(lib().km_table_get_info(self._h, ctypes.byref(info)))
        return {"k": info.k, "canonical": bool(info.canonical), "device": info.device, "layout": int(info.reserved),
                "n_keys": int(info.n_keys), "n_buckets": int(info.n_buckets), "bytes": int(info.bytes)}

    def insert(self, keys, counts, mode="overwrite"):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        if keys.shape != counts.shape:
            raise ValueError("keys and counts differ in shape")
        m = {"keep": 0, "overwrite": 1, "add": 2}[mode]
        check(lib().km_table_insert(self._h, keys.ctypes.data, counts.ctypes.data, keys.size, m))

    def build_synthetic(self, seed, n_keys):
        check(lib().km_table_build_synthetic(self._h, int(seed), int(n_keys)))

    def count_reads(self, reads):
        blob = "".join(reads).encode("ascii")
        off = np.zeros(len(reads) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in reads], out=off[1:])
        check(lib().km_table_count_reads(self._h, blob, off.ctypes.data, len(reads)))

    def count_text(self, text, qual=None, min_qual=None):
        """Counts the k-mers of a byte stream of sequences separated by any non-ACGT byte (bytes / bytearray / uint8
        array); `qual` = the FASTQ quality character of every byte (same length) and `min_qual` the `-Q` threshold,
        applied on the device."""
        buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, dtype=np.uint8)
        q = 0 if min_qual is None else (ord(min_qual[0]) if isinstance(min_qual, str) else int(min_qual))
        qp = None
        if qual is not None and q > 0:
            qbuf = np.frombuffer(qual, dtype=np.uint8) if not isinstance(qual, np.ndarray) else np.ascontiguousarray(qual, dtype=np.uint8)
            if qbuf.size != buf.size:
                raise ValueError("qual and text differ in length")
            qp = qbuf.ctypes.data
        check(lib().km_table_count_text(self._h, buf.ctypes.data, qp, buf.size, q))

    def link(self):
        """Write the neighbour masks now (km_table_link) instead of at the next find call."""
        check(lib().km_table_link(self._h))

    def recount(self):
        """Number of records in this table (this shard), counted on the device."""
        n = ctypes.c_uint64()
        check(lib().km_table_recount(self._h, ctypes.byref(n)))
        return int(n.value)

    def count_file(self, path, min_qual=None):
        """Counts the k-mers of a FASTA / FASTQ file (plain or .gz) read by the library itself; min_qual = the
        `-Q` quality character (or its byte value).  Returns (reads, bases)."""
        q = 0 if min_qual is None else (ord(min_qual[0]) if isinstance(min_qual, str) else int(min_qual))
        nr, nb = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib().km_table_count_file(self._h, os.fsencode(path), q, ctypes.byref(nr), ctypes.byref(nb)))
        return int(nr.value), int(nb.value)

    def count_reads_blob(self, blob, offsets):
        """The same from one bytes object of concatenated reads and its int64 offsets [n_reads + 1]."""
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        check(lib().km_table_count_reads(self._h, blob, off.ctypes.data, len(off) - 1))

    def drop_below(self, min_count):
        left = ctypes.c_uint64()
        check(lib().km_table_drop_below(self._h, int(min_count), ctypes.byref(left)))
        return int(left.value)

    def export(self):
        """`jellyfish dump`: (canonical keys uint64[n], counts uint32[n]) of every record, in no particular order."""
        n = self.info()["n_keys"]
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint32)
        got = ctypes.c_uint64()
        check(lib().km_table_export(self._h, keys.ctypes.data, counts.ctypes.data, n, ctypes.byref(got)))
        m = min(n, int(got.value))
        return keys[:m], counts[:m]

    def write_jf(self, path, counter_len=4):
        """The table as a Jellyfish binary/sorted file (what Jellyfish(filename) and Table.open_jf read)."""
        check(lib().km_table_write_jf(self._h, os.fsencode(path), int(counter_len)))

    def query_packed(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(kmers.shape, dtype=np.uint32)
        check(lib().km_query_batch(self._h, kmers.ctypes.data, kmers.size, out.ctypes.data))
        return out

    def query_ascii(self, kmers):
        """list of k-letter strings -> uint32 counts."""
        blob = "".join(kmers).encode("ascii")
        if len(blob) != self.k * len(kmers):
            raise ValueError("every k-mer must be %d letters" % self.k)
        out = np.empty(len(kmers), dtype=np.uint32)
        check(lib().km_query_ascii(self._h, blob, len(kmers), out.ctypes.data))
        return out

    def get_child_packed(self, kmers, ratio, count, forward=True):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        counts = np.empty((kmers.size, 4), dtype=np.uint32)
        mask = np.empty(kmers.size, dtype=np.uint8)
        check(lib().km_get_child_batch(self._h, kmers.ctypes.data, kmers.size, int(bool(forward)), float(ratio),
                                       int(count), counts.ctypes.data, mask.ctypes.data))
        return counts, mask

    @staticmethod
    def _pack_targets(sequences):
        blob = "".join(sequences).encode("ascii")
        off = np.zeros(len(sequences) + 1, dtype=np.int64)
        if sequences:
            np.cumsum(np.fromiter(map(len, sequences), dtype=np.int64, count=len(sequences)), out=off[1:])
        return blob, off

    def find_batch(self, sequences, count=5, ratio=0.05, steps=500, branchs=10, nodes=10000, extra_nodes=0,
                   want_graph=True, no_refine_jump=False):
        """Run the whole find_mutation path for a list of target sequences in one call.
        want_graph=False skips copying node arrays / index paths back (rows and text only)."""
        if isinstance(sequences, PackedTargets):
            blob, off, seqs = sequences.blob, sequences.off, sequences.sequences
        else:
            seqs = list(sequences)
            blob, off = self._pack_targets(seqs)
        prm = FindParams(float(ratio), int(count), int(steps), int(branchs), int(nodes), int(extra_nodes),
                         (0 if want_graph else 1) | (2 if no_refine_jump else 0), 0)
        h = ctypes.c_void_p()
        check(lib().km_find_batch(self._h, blob, off.ctypes.data, len(seqs), ctypes.byref(prm), ctypes.byref(h)))
        return BatchResult.from_handle(h, seqs, self.k)

    def find_text(self, targets, db_name, count=5, ratio=0.05, steps=500, branchs=10, nodes=10000, extra_nodes=0,
                  n_sub=0, as_bytes=False):
        """The text `km find_mutation` prints for a batch (rows of every target, sorted, in target
        order) in one library call; `targets` is a PackedTargets with names.  Sub-batches are in
        flight at once, so copies, kernels and text building overlap.  Returns (text, status array)."""
        if not isinstance(targets, PackedTargets) or targets.names is None:
            raise TypeError("find_text takes PackedTargets(sequences, names)")
        prm = FindParams(float(ratio), int(count), int(steps), int(branchs), int(nodes), int(extra_nodes), 1, 0)
        h = ctypes.c_void_p()
        db = db_name.encode()
        check(lib().km_find_text(self._h, targets.blob, targets.off.ctypes.data, len(targets), ctypes.byref(prm), db,
                                 targets.name_blob, targets.name_off.ctypes.data, int(n_sub), ctypes.byref(h)))
        out = TextResult(h, db, targets)
        self.last_timing = out.timing
        view = out.text_view()
        return ((view if as_bytes else view.tobytes().decode()), out.status)

    def plan(self, sequences, count=5, ratio=0.05, steps=500, branchs=10, nodes=10000, extra_nodes=0):
        """Upload a batch once; launch it any number of times (FindPlan)."""
        return FindPlan(self, sequences, count, ratio, steps, branchs, nodes, extra_nodes)

    def close(self):
        if getattr(self, "_h", None):
            lib().km_table_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FindPlan:
    """km_find_plan_*: a batch resident in HBM.  launch() enqueues the two kernels on a CUDA stream
    (asynchronously), fetch() synchronises and returns a BatchResult."""

    def __init__(self, table, sequences, count, ratio, steps, branchs, nodes, extra_nodes):
        self.table = table
        self.sequences = list(sequences)
        blob, off = Table._pack_targets(self.sequences)
        prm = FindParams(float(ratio), int(count), int(steps), int(branchs), int(nodes), int(extra_nodes), 0, 0)
        self._h = ctypes.c_void_p()
        check(lib().km_find_plan_create(table._h, blob, off.ctypes.data, len(self.sequences), ctypes.byref(prm),
                                        ctypes.byref(self._h)))

    def launch(self, stream=None):
        check(lib().km_find_plan_launch(self._h, ctypes.c_void_p(stream) if stream else None))

    def last_ms(self):
        """(walk_ms, graph_ms) of the most recent launch, from CUDA events on its stream."""
        w, g = ctypes.c_float(), ctypes.c_float()
        check(lib().km_find_plan_last_ms(self._h, ctypes.byref(w), ctypes.byref(g)))
        return w.value, g.value

    def kernel_ms(self):
        """(probe_ms, walk_ms, graph_ms) of the most recent launch: reference-probe kernel (with the
        memsets), the two walk kernels, the two graph kernels."""
        out = (ctypes.c_float * 3)()
        check(lib().km_find_plan_kernel_ms(self._h, out))
        return out[0], out[1], out[2]

    def fetch(self, want_graph=True):
        h = ctypes.c_void_p()
        check(lib().km_find_plan_fetch(self._h, int(bool(want_graph)), ctypes.byref(h)))
        return BatchResult.from_handle(h, self.sequences, self.table.k)

    def close(self):
        if getattr(self, "_h", None):
            lib().km_find_plan_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TextResult:
    """Owner of a km_find_text result.  `text_view()` returns a uint8 array over the library's text buffer
    (no copy); the array keeps this object -- and with it the buffer -- alive, and this object does not
    refer back to the array, so the result is released as soon as the last view goes away."""

    def __init__(self, handle, db, targets):
        self._h = handle
        v = ResultView()
        check(lib().km_result_get(handle, ctypes.byref(v)))
        self.status = np.array(_view(v.status, np.uint32, v.n_targets))
        self.timing = {"h2d_ms": v.ms_h2d, "walk_ms": v.ms_walk, "graph_ms": v.ms_graph, "d2h_ms": v.ms_d2h,
                       "launches": v.n_launches, "retries": v.n_retries, "h2d_bytes": int(v.bytes_h2d),
                       "d2h_bytes": int(v.bytes_d2h)}
        ptr = ctypes.c_void_p()
        need = lib().km_result_text(handle, db, targets.name_blob, targets.name_off.ctypes.data, 0, ctypes.byref(ptr))
        if need < 0:
            check(int(need))
        self._ptr, self._len = ptr.value, int(need)

    def text_view(self):
        if not self._len:
            return np.zeros(0, dtype=np.uint8)
        raw = (ctypes.c_char * self._len).from_address(self._ptr)
        raw._owner = self
        return np.frombuffer(raw, dtype=np.uint8, count=self._len)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().km_result_free(self._h)
                self._h = None
        except Exception:
            pass


def _view(ptr, dtype, n):
    if not ptr or n <= 0:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


class BatchResult:
    """Flat result of km_find_batch (see km_result_view in include/km_b200.h)."""

    def __init__(self):
        self._h = None

    @classmethod
    def from_handle(cls, h, sequences, k):
        r = cls()
        r._h = h
        v = ResultView()
        check(lib().km_result_get(h, ctypes.byref(v)))
        n = v.n_targets
        r.k = int(v.k)
        r.sequences = sequences
        r.status = _view(v.status, np.uint32, n)
        r.n_nodes = _view(v.n_nodes, np.int32, n)
        r.node_off = _view(v.node_off, np.int64, n + 1)
        r.has_graph = bool(v.has_graph)
        total_nodes = int(r.node_off[-1]) if (n and r.has_graph) else 0
        r.node_kmer = _view(v.node_kmer, np.uint64, total_nodes)
        r.node_count = _view(v.node_count, np.uint32, total_nodes)
        r.path_first = _view(v.path_first, np.int32, n)
        r.path_count = _view(v.path_count, np.int32, n)
        r.path_off = _view(v.path_off, np.int64, v.n_paths)
        r.path_len = _view(v.path_len, np.int32, v.n_paths)
        pool_n = int((r.path_off + r.path_len).max()) if (v.n_paths and r.has_graph) else 0
        r.path_pool = _view(v.path_pool, np.int32, pool_n)
        r.row_first = _view(v.row_first, np.int32, n)
        r.row_count = _view(v.row_count, np.int32, n)
        r.rows = _view(v.rows, ROW_DTYPE, v.n_rows)
        r.lookups = _view(v.lookups, np.uint64, n)
        r.timing = {"h2d_ms": v.ms_h2d, "walk_ms": v.ms_walk, "graph_ms": v.ms_graph, "d2h_ms": v.ms_d2h,
                    "total_ms": v.ms_total, "launches": v.n_launches, "retries": v.n_retries,
                    "h2d_bytes": int(v.bytes_h2d), "d2h_bytes": int(v.bytes_d2h), "simple_graphs": int(v.reserved)}
        return r

    # ---- per-target accessors ------------------------------------------------------------
    def n_targets(self):
        return len(self.status)

    def kmers(self, t):
        """MutationFinder.kmer: node k-mers in canonical numbering + the two caps."""
        n = int(self.n_nodes[t]) - 2
        o = int(self.node_off[t])
        return [unpack_kmer(v, self.k) for v in self.node_kmer[o:o + n]] + [CAP_SRC, CAP_SNK]

    def counts(self, t):
        """MutationFinder.counts (caps carry -1, MutationFinder.py:123)."""
        n = int(self.n_nodes[t]) - 2
        o = int(self.node_off[t])
        return [int(c) for c in self.node_count[o:o + n]] + [-1, -1]

    def paths(self, t):
        """Unique alternative paths as index tuples, lexicographic order."""
        out = []
        for p in range(int(self.path_first[t]), int(self.path_first[t]) + int(self.path_count[t])):
            o = int(self.path_off[p])
            out.append(tuple(int(x) for x in self.path_pool[o:o + int(self.path_len[p])]))
        return out

    def _spell(self, t, path, begin, end):
        if end <= begin:
            return ""
        o = int(self.node_off[t])
        o_p = int(self.path_off[path])
        idx = self.path_pool[o_p + begin:o_p + end]
        first = unpack_kmer(self.node_kmer[o + int(idx[0])], self.k)
        tail = (self.node_kmer[o + idx[1:]] & np.uint64(3)).astype(np.int64)
        return first + "".join(BASES[c] for c in tail.tolist())

    def row_fields(self, t, db_name, query_name):
        """Rows of target t in the reference's emission order (quantify_paths rows, then
        quantify_clusters rows) as tuples ready for PathQuant.Path."""
        k = self.k
        seq = self.sequences[t]
        out = []
        for i in range(int(self.row_first[t]), int(self.row_first[t]) + int(self.row_count[t])):
            w = self.rows[i]
            typ = TYPE_NAMES[int(w["type"])]
            p = int(w["path_id"])
            o_p = int(self.path_off[p])
            o_n = int(self.node_off[t])
            if typ == "Reference":
                name = "Reference\t"                                   # MutationFinder.py:480-481
            else:
                db, dl = int(w["del_begin"]), int(w["del_len"])
                ib, il = int(w["ins_begin"]), int(w["ins_len"])
                gone = seq[db + k - 1:db + k - 1 + dl].lower()
                idx = self.path_pool[o_p + ib:o_p + ib + il]
                new = "".join(BASES[int(c)] for c in (self.node_kmer[o_n + idx] & np.uint64(3)).tolist())
                name = "%s\t%d:%s/%s:%d" % (typ, int(w["name_start"]), gone, new, int(w["name_end"]))
            rb, re_ = int(w["ref_begin"]), int(w["ref_end"])
            ref_seq = seq[rb:re_ + k - 1] if re_ > rb else ""
            note = "vs_ref" if int(w["kind"]) == 0 else "cluster %d n=%d" % (int(w["cluster_id"]), int(w["cluster_n"]))
            out.append((db_name, query_name, name, float(w["rvaf"]), float(w["expr"]), int(w["min_cov"]),
                        int(w["start_off"]), self._spell(t, p, int(w["var_begin"]), int(w["var_end"])),
                        float(w["ref_rvaf"]), float(w["ref_expr"]), ref_seq, note))
        return out

    def format_target(self, t, db_name, query_name):
        """Sorted TSV text of target t, formatted by the library (the CLI's fast path)."""
        need = lib().km_result_format_target(self._h, int(t), db_name.encode(), query_name.encode(), None, 0)
        if need < 0:
            check(int(need))
        buf = ctypes.create_string_buffer(int(need) + 1)
        lib().km_result_format_target(self._h, int(t), db_name.encode(), query_name.encode(), buf, int(need) + 1)
        return buf.raw[:int(need)].decode()

    def format_all(self, db_name, names, threads=0, as_bytes=False):
        """Sorted TSV text of every target, in target order, formatted on host threads.
        as_bytes=True returns a uint8 numpy array (no extra copies of a multi-megabyte text)."""
        if isinstance(names, PackedTargets):
            blob, off = names.name_blob, names.name_off
        else:
            blob, off = _pack_names(list(names))
        ptr = ctypes.c_void_p()
        need = lib().km_result_text(self._h, db_name.encode(), blob, off.ctypes.data, int(threads), ctypes.byref(ptr))
        if need < 0:
            check(int(need))
        view = _view(ptr.value, np.uint8, int(need))      # the library's own buffer, valid while the result lives
        if as_bytes:
            return view
        return view.tobytes().decode()

    def close(self):
        if getattr(self, "_h", None):
            # drop numpy views before the C++ vectors go away
            for name in ("status", "n_nodes", "node_off", "node_kmer", "node_count", "path_first", "path_count",
                         "path_off", "path_len", "path_pool", "row_first", "row_count", "rows", "lookups"):
                setattr(self, name, np.array(getattr(self, name)))
            lib().km_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().km_result_free(self._h)
                self._h = None
        except Exception:
            pass


def raise_for_status(status, name, max_node):
    """Turn a per-target status into the reference's error behaviour."""
    status = int(status)
    if status & ST_NODE_LIMIT:
        # MutationFinder.py:143-148
        sys.exit("ERROR: Node query count limit exceeded: max={}".format(max_node))
    if status & ST_BAD_BASE:
        raise ValueError("target %s holds a letter outside ACGT (behaviour of the reference is unpinned)" % name)
    if status & ST_DUP_KMER:
        raise ValueError("a k-mer occurs multiple times in reference %s" % name)
    if status & ST_TOO_SHORT:
        raise AssertionError("target %s is shorter than k" % name)      # Sequence.py:45 `assert len(self.ref_mer)`
    if status & ST_NODE_OVERFLOW:
        raise RuntimeError("km_b200: target %s visits more nodes off the reference than max_node + 4 * max_stack + 4096 "
                           "allows (the reference bounds only the nodes it keeps; its own walk would not end here)" % name)
    if status & (ST_TOO_MANY_COLS | ST_SOLVER_WATCHDOG | ST_NAME_MISMATCH | ST_PATH_OVERFLOW):
        raise RuntimeError("km_b200: target %s could not be processed (status 0x%x)" % (name, status))
