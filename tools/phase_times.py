"""Per-phase SM cycles of the graph pass (build with KM_PHASE_TIMERS=1).  Run on a GPU box:
    KM_PHASE_TIMERS=1 python tools/phase_times.py [n_targets]          (KM_ONLY=i,j,...: only those targets of the panel)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if not os.environ.get("KM_B200_LIB"):      # else: a variant built with -DKM_PHASE_TIMERS (tools/build_variant.py)
    os.environ["KM_PHASE_TIMERS"] = "1"
    from km_b200 import build as kb     # noqa: E402
    kb.build(force=True)
from km_b200 import engine, synth   # noqa: E402
from km_b200._lib import lib        # noqa: E402

NAMES = {0: "numbering", 1: "adjacency", 2: "shortest trees", 3: "strip chain", 4: "candidates", 5: "alloc",
         6: "materialise", 7: "sort paths", 8: "spell", 10: "diffs", 11: "clusters", 12: "vs_ref setup / cluster rows",
         13: "cluster rows tail", 14: "solve_columns", 15: "min_count"}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
panel = synth.make_panel(n, seed=synth.PANEL_SEED)
only = [int(x) for x in os.environ.get("KM_ONLY", "").split(",") if x]      # KM_ONLY=794: the phases of that one target
if only:
    panel.targets = [panel.targets[i] for i in only]
    n = len(only)
t = engine.Table.create(capacity=50_000_000 + len(panel.keys))
t.build_synthetic(synth.TABLE_SEED, 50_000_000)
t.insert(panel.keys, panel.counts)
plan = t.plan(panel.targets)
for _ in range(3):
    plan.launch()
plan.last_ms()
L = lib()
L.km_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 64)()
L.km_debug_walk_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
wbuf = (ctypes.c_ulonglong * 64)()
L.km_debug_phase_cycles(buf, 1)
L.km_debug_walk_cycles(wbuf, 1)
plan.launch()
w, g = plan.last_ms()
L.km_debug_phase_cycles(buf, 0)
L.km_debug_walk_cycles(wbuf, 0)
for i in range(32, 40):
    buf[i] = wbuf[i]
tot = sum(buf[:16])
print("walk %.3f ms  graph %.3f ms   targets %d   (probe %.3f, walks %.3f, graph %.3f)" % ((w, g, n) + tuple(plan.kernel_ms())))
WALK = {32: "walk: set-up", 33: "walk: phase 1 (ref k-mers)", 34: "walk: level 0", 35: "walk: later levels", 36: "walk: peel", 37: "walk: results"}
for i in range(32, 40):
    if buf[i]:
        print("%2d %-28s %8.1f cycles/target" % (i, WALK[i], buf[i] / n))
BUB = {40: "bubble: numbering", 41: "bubble: k-mer map", 42: "bubble: edges", 43: "bubble: test + chain", 44: "bubble: alloc",
       45: "bubble: paths + spelling", 46: "bubble: rows", 48: "  cluster row: quant_pair", 49: "  cluster row: diff_paths",
       50: "  cluster row: write_row", 51: "  vs_ref row: quant_pair", 52: "  vs_ref row: write_row"}
BUB.update({53: "  wide cluster: Gram sums", 54: "  wide cluster: eigen solve", 55: "  wide cluster: refine (all)", 56: "  refine_jump calls (every row)"})
for i in range(40, 57):
    if buf[i]:
        print("%2d %-28s %8.1f cycles/target" % (i, BUB[i], buf[i] / n))
print("simple graphs:", plan.fetch(want_graph=False).timing.get("simple_graphs"))
TREE = {16: "fwd tree: all iterations", 17: "fwd run", 18: "fwd simple step", 19: "fwd general", 20: "bwd all", 21: "bwd run", 22: "bwd simple step", 23: "bwd general"}
for i in range(16, 24):
    if buf[i]:
        print("%2d %-28s %8.1f cycles/target  count/target %.1f  cycles/count %.1f" % (i, TREE[i], buf[i] / n, buf[i + 8] / n, buf[i] / max(1, buf[i + 8])))
for i, c in enumerate(buf[:16]):
    if c:
        print("%2d %-28s %8.1f cycles/target  %5.1f%%" % (i, NAMES.get(i, "?"), c / n, 100.0 * c / tot))
print("sum %.0f cycles/target" % (tot / n))
