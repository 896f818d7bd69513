"""Which targets does the graph pass spend its time on?  (KM_PHASE_TIMERS build, run on a GPU box)
    python tools/slow_targets.py [panel seed offset] [n_targets]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                  # noqa: E402
if not os.environ.get("KM_B200_LIB"):      # else: a variant built with -DKM_PHASE_TIMERS (tools/build_variant.py)
    os.environ["KM_PHASE_TIMERS"] = "1"
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
L = lib()
for _ in range(3):
    plan.launch()
plan.last_ms()
tl = (ctypes.c_ulonglong * 32)()
L.km_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.km_debug_timeline(None, 1)
plan.launch()
plan.last_ms()
L.km_debug_timeline(tl, 0)
t0 = min(tl[2 * i] for i in range(16) if tl[2 * i + 1])
NAMES = ["scheduler", "CTA-per-target 256", "CTA-per-target 512", "general", "bubbles 256", "bubbles 512", "", ""] + ["target #%d of the CTA-per-target passes" % i for i in range(8)]
print("graph phase timeline (us after its first kernel started):")
for i in range(16):
    if tl[2 * i + 1]:
        print("  %-42s %8.1f .. %8.1f" % (NAMES[i], (tl[2 * i] - t0) / 1e3, (tl[2 * i + 1] - t0) / 1e3))
print("kernel ms (probe, walk, graph):", plan.kernel_ms())
res = plan.fetch(want_graph=False)
L.km_debug_target_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
cyc = np.zeros(n, dtype=np.uint32)
if L.km_debug_target_cycles(cyc.ctypes.data, n) != 0:
    sys.exit(0)                       # a -DKM_TIMELINE build: the timeline above is all there is
order = np.argsort(-cyc.astype(np.int64))
print("graph-pass cycles per target: mean %.0f  median %.0f  p99 %.0f  max %d" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 99), cyc.max()))
for i in order[:15]:
    tr = panel.truth[i]
    print("target %5d  cycles %9d  len %3d  nodes %4d  paths %3d  rows %3d  lookups %6d  truth %s" % (
        i, cyc[i], len(panel.targets[i]), res.n_nodes[i], res.path_count[i], res.row_count[i], res.lookups[i], tr))
