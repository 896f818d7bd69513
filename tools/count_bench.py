"""Throughput of the device k-mer counter (km_table_count_reads): synthetic reads sampled from a random
genome (both strands), one call per batch, host buffers in.
    python tools/count_bench.py [n_reads] [read_len] [genome_len]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                  # noqa: E402
from km_b200 import engine          # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 100
genome_len = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000_000
rng = np.random.default_rng(1)
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=genome_len)]
starts = rng.integers(0, genome_len - read_len, size=n_reads)
idx = starts[:, None] + np.arange(read_len)[None, :]
reads = genome[idx]                                   # (n_reads, read_len) bytes, forward strand only (the counter canonicalises)
blob = reads.tobytes()
off = np.arange(n_reads + 1, dtype=np.int64) * read_len
t = engine.Table.create(k=31, canonical=True, capacity=int(1.3 * genome_len))
out = {"n_reads": n_reads, "read_len": read_len, "kmers_per_call": n_reads * (read_len - 30), "runs": []}
for it in range(3):
    t0 = time.perf_counter()
    t.count_reads_blob(blob, off)
    dt = time.perf_counter() - t0
    out["runs"].append({"s": dt, "kmers_per_s": out["kmers_per_call"] / dt, "bases_per_s": n_reads * read_len / dt,
                        "distinct": t.info()["n_keys"]})
t0 = time.perf_counter()
left = t.drop_below(2)
out["drop_below_2_s"] = time.perf_counter() - t0
out["left"] = left
print(json.dumps(out))
