"""Either side of the path at real-sample scale (example/run_leucegene.sh:22-27), on one GPU:
  * km_count_text_kernel alone on a device-resident byte stream (CUDA events): k-mers/s, GB/s of reads
  * km_table_count_file on a plain FASTQ file written to /tmp (host reader + H2D + kernel, -Q on the device): GB/s
  * km_table_write_jf of a 1e8-record table, then km_table_open_jf of that file: records/s, GB/s
Prints one JSON line.   python tools/io_bench.py [--records 100000000] [--reads 20000000]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from km_b200 import engine, synth          # noqa: E402
from km_b200._lib import check, lib        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--fastq-reads", type=int, default=4_000_000)
    ap.add_argument("--tmp", default="/tmp")
    a = ap.parse_args()
    out = {}
    # ---- count kernel alone ---------------------------------------------------------------------------------------
    t = engine.Table.create(capacity=2 * a.genome + (1 << 20))
    ms = (ctypes.c_float * 2)()
    nk = ctypes.c_uint64()
    check(lib().km_bench_count(t._h, a.reads, 100, a.genome, 7, 3, ms, ctypes.byref(nk)))
    n_bytes = a.reads * 101
    out["count_kernel"] = {"reads": a.reads, "read_len": 100, "genome_bases": a.genome, "coverage": a.reads * 100 / a.genome,
                           "kmers": int(nk.value), "ms_existing_keys": ms[0], "ms_first_pass_new_keys": ms[1],
                           "kmers_per_s": nk.value / (ms[0] / 1e3), "read_GBps": n_bytes / (ms[0] / 1e3) / 1e9,
                           "kmers_per_s_first_pass": nk.value / (ms[1] / 1e3), "distinct_keys": t.info()["n_keys"],
                           "hbm_note": "per k-mer: one 16-byte key read + one atomic add on a random bucket (a 128-byte DRAM line "
                                       "read and, once dirty, written back): ~256 B of DRAM traffic per k-mer when the table exceeds L2"}
    t.close()
    # ---- FASTQ file through the host reader -------------------------------------------------------------------------
    rng = np.random.default_rng(3)
    genome = rng.integers(0, 4, size=5_000_000, dtype=np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[genome]
    starts = rng.integers(0, len(genome) - 100, size=a.fastq_reads)
    fq = os.path.join(a.tmp, "km_io_bench.fastq")
    with open(fq, "wb") as f:
        qual = np.where(rng.random(100) < 0.02, 35, rng.integers(43, 75, size=100)).astype(np.uint8).tobytes()   # ~2 % of positions below '+'
        for lo in range(0, a.fastq_reads, 200_000):
            chunk = []
            for i, s in enumerate(starts[lo:lo + 200_000].tolist()):
                chunk.append(b"@r%d\n" % (lo + i))
                chunk.append(letters[s:s + 100].tobytes())
                chunk.append(b"\n+\n")
                chunk.append(qual)
                chunk.append(b"\n")
            f.write(b"".join(chunk))
    size = os.path.getsize(fq)
    with open(fq, "rb") as f:           # warm the page cache, like `wc -l sample.jf` in run_leucegene.sh:26-27
        while f.read(1 << 26):
            pass
    for label, q in (("no_Q", None), ("Q+", "+")):
        t = engine.Table.create(capacity=2 * len(genome) + (1 << 20))
        t0 = time.perf_counter()
        nr, nb = t.count_file(fq, min_qual=q)
        dt = time.perf_counter() - t0
        out["count_file_" + label] = {"file_bytes": size, "reads": nr, "bases": nb, "seconds": dt, "file_GBps": size / dt / 1e9,
                                      "kmers_per_s": (nb - 30 * nr) / dt, "distinct_keys": t.info()["n_keys"]}
        t.close()
    os.remove(fq)
    # ---- .jf writer and loader ------------------------------------------------------------------------------------
    t = engine.Table.create(capacity=a.records + (1 << 20))
    t.build_synthetic(synth.TABLE_SEED, a.records)
    n = t.info()["n_keys"]
    jf = os.path.join(a.tmp, "km_io_bench.jf")
    t0 = time.perf_counter()
    t.write_jf(jf)
    w = time.perf_counter() - t0
    probe = synth.background_keys(synth.TABLE_SEED, 0, 100000)
    want = t.query_packed(probe)
    t.close()
    size = os.path.getsize(jf)
    t0 = time.perf_counter()
    t2 = engine.Table.open_jf(jf)
    r = time.perf_counter() - t0
    same = bool((t2.query_packed(probe) == want).all()) and t2.info()["n_keys"] == n
    out["jf"] = {"records": n, "file_bytes": size, "write_s": w, "open_s": r, "open_records_per_s": n / r, "open_GBps": size / r / 1e9,
                 "round_trip_equal": same}
    t2.close()
    os.remove(jf)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
