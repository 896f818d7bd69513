"""Per-call wall time of km_find_text on the bench panel, 300 calls, without and with an `nvidia-smi -lms 100` loop beside it
(what bench.py's clock sampler runs during its timed regions): where do occasional slow steps come from?"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                  # noqa: E402
from km_b200 import engine, synth   # noqa: E402

panel = synth.make_panel(10000, seed=synth.PANEL_SEED)
t = engine.Table.create(capacity=2_000_000_000 + len(panel.keys))
t.build_synthetic(synth.TABLE_SEED, 2_000_000_000)
t.insert(panel.keys, panel.counts)
packed = engine.PackedTargets(panel.targets, panel.names)
for _ in range(8):
    t.find_text(packed, "panel.jf", as_bytes=True)


def run(tag):
    ms = []
    for _ in range(300):
        t0 = time.perf_counter()
        text, status = t.find_text(packed, "panel.jf", as_bytes=True)
        ms.append(1e3 * (time.perf_counter() - t0))
    a = np.sort(np.array(ms))
    print("%-28s median %.3f  mean %.3f  p90 %.3f  p99 %.3f  max %.3f ms; calls over 1.5 ms: %d at %s" % (
        tag, np.median(a), a.mean(), a[269], a[296], a[-1], int((np.array(ms) > 1.5).sum()),
        [i for i, v in enumerate(ms) if v > 1.5][:12]))


run("alone")
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.5)
run("with nvidia-smi -lms 100")
p.terminate()
run("alone again")
