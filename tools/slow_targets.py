"""Which targets does the graph pass spend its time on?  (KM_PHASE_TIMERS build, run on a GPU box)
    python tools/slow_targets.py [panel seed offset] [n_targets]"""
import ctypes
import os
import sys

os.environ["KM_PHASE_TIMERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                  # noqa: E402
from km_b200 import build as kb     # noqa: E402
kb.build(force=True)
from km_b200 import engine, synth   # noqa: E402
from km_b200._lib import lib        # noqa: E402

off = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
panel = synth.make_panel(n, seed=synth.PANEL_SEED + off)
t = engine.Table.create(capacity=50_000_000 + len(panel.keys))
t.build_synthetic(synth.TABLE_SEED, 50_000_000)
t.insert(panel.keys, panel.counts)
plan = t.plan(panel.targets)
for _ in range(3):
    plan.launch()
print("kernel ms (probe, walk, graph):", plan.kernel_ms())
res = plan.fetch(want_graph=False)
L = lib()
L.km_debug_target_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
cyc = np.zeros(n, dtype=np.uint32)
assert L.km_debug_target_cycles(cyc.ctypes.data, n) == 0
order = np.argsort(-cyc.astype(np.int64))
print("graph-pass cycles per target: mean %.0f  median %.0f  p99 %.0f  max %d" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 99), cyc.max()))
for i in order[:15]:
    tr = panel.truth[i]
    print("target %5d  cycles %9d  len %3d  nodes %4d  paths %3d  rows %3d  lookups %6d  truth %s" % (
        i, cyc[i], len(panel.targets[i]), res.n_nodes[i], res.path_count[i], res.row_count[i], res.lookups[i], tr))
