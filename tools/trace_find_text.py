"""Host-clock timeline of km_find_text on the bench configuration (KM_TRACE):
    python tools/trace_find_text.py [n_sub] [panel seed offset]          (KM_TRACE=2: device timeline too)"""
import os
import sys
import time

os.environ.setdefault("KM_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from km_b200 import engine, synth   # noqa: E402

n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 0
panel = synth.make_panel(10000, seed=synth.PANEL_SEED + (int(sys.argv[2]) if len(sys.argv) > 2 else 0))
t = engine.Table.create(capacity=2_000_000_000 + len(panel.keys))
t.build_synthetic(synth.TABLE_SEED, 2_000_000_000)
t.insert(panel.keys, panel.counts)
packed = engine.PackedTargets(panel.targets, panel.names)
for i in range(5):
    sys.stderr.write("---- call %d\n" % i)
    t0 = time.perf_counter()
    text, status = t.find_text(packed, "panel.jf", as_bytes=True, n_sub=n_sub)
    sys.stderr.write("python-side %.3f ms, %d bytes\n" % (1e3 * (time.perf_counter() - t0), len(text)))
