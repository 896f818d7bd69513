"""A small batch through every kernel (bundled-size table, mixed targets incl. long ones) for
compute-sanitizer:   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from km_b200 import engine, synth      # noqa: E402

panel = synth.make_panel(48, seed=5, two_variant_frac=0.3)
long_panel = synth.make_panel(8, seed=6, len_lo=500, len_hi=800)
targets = panel.targets + long_panel.targets
names = panel.names + ["long_" + n for n in long_panel.names]
t = engine.Table.create(capacity=1 << 16)
t.build_synthetic(synth.TABLE_SEED, 20000)
t.insert(panel.keys, panel.counts, mode="overwrite")
t.insert(long_panel.keys, long_panel.counts, mode="overwrite")
res = t.find_batch(targets)
text = res.format_all("s.jf", names)
packed = engine.PackedTargets(targets, names)
text2, status = t.find_text(packed, "s.jf", n_sub=3)
assert text == text2 and not (status & (0xFFFFFFFF ^ 16)).any()
text3, _ = t.find_text(packed, "s.jf", n_sub=3)          # the same layout again: captured into a CUDA graph ...
text4, _ = t.find_text(packed, "s.jf", n_sub=3)          # ... and replayed
assert text3 == text and text4 == text
q = synth.lookup_queries(4096, synth.TABLE_SEED, 20000)
t.query_packed(q)
# counting from reads (rolling kernel, -Q mask on the device), -L 2, neighbour masks, and the walk over the counted table
reads = synth.sample_reads(panel, range(12))
c = engine.Table.create(capacity=1 << 16)
c.count_text(reads, qual=bytes([60]) * len(reads), min_qual="+")
c.drop_below(2)
c.link()
res2 = c.find_batch(panel.targets[:12])
assert not (res2.status & (0xFFFFFFFF ^ 16)).any() and int(res2.row_count.sum()) >= 12
print("ok", len(text), "bytes of rows")
