"""How many targets of the bench panel the bubble pass finishes itself (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from km_b200 import engine, synth
n = 10000
keys = int(float(os.environ.get("KM_PT_KEYS", "2e9")))
panel = synth.make_panel(n, seed=synth.PANEL_SEED)
t = engine.Table.create(capacity=keys + len(panel.keys))
t.build_synthetic(synth.TABLE_SEED, keys)
t.insert(panel.keys, panel.counts)
plan = t.plan(panel.targets)
for _ in range(3):
    plan.launch()
res = plan.fetch(want_graph=False)
print("simple_graphs raw", res.timing.get("simple_graphs"), "kernel ms", plan.kernel_ms())
