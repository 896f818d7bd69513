import os, sys
os.environ["KM_TABLE_LINES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from collections import Counter
from km_b200 import engine
from oracle import jf_format
rng = np.random.default_rng(12)
genome = "".join("ACGT"[i] for i in rng.integers(0, 4, size=5000))
reads = []
for _ in range(3000):
    s = int(rng.integers(0, 5000 - 100)); r = genome[s:s + 100]
    if rng.random() < 0.5: r = r.translate(str.maketrans("ACGT", "TGCA"))[::-1]
    reads.append(r)
host = Counter()
for r in reads:
    for i in range(len(r) - 30):
        v = jf_format.pack(r[i:i + 31]); host[min(v, jf_format.revcomp_packed(v, 31))] += 1
t = engine.Table.create(capacity=4 * len(host) + 1024)
t.count_reads(reads)
keys = np.array(list(host.keys()), dtype=np.uint64)
got = t.query_packed(keys)
want = np.array([host[int(k)] for k in keys], dtype=np.uint32)
print("before drop: mismatches", int((got != want).sum()), "of", len(keys))
rc = np.array([jf_format.revcomp_packed(int(k), 31) for k in keys], dtype=np.uint64)
got_rc = t.query_packed(rc)
print("queried as reverse complement: mismatches", int((got_rc != want).sum()))
left = t.drop_below(2)
got2 = t.query_packed(keys); want2 = np.where(want >= 2, want, 0)
bad = np.nonzero(got2 != want2)[0]
print("after drop: mismatches", len(bad), "left", left, "expected", int((want >= 2).sum()))
for i in bad[:10]:
    print(hex(int(keys[i])), "got", got2[i], "want", want2[i], "rc query", t.query_packed(np.array([rc[i]], dtype=np.uint64))[0])
