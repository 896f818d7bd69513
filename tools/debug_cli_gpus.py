"""Debug aid: `km find_mutation` on one GPU and with --gpus 2 over the first 40 targets of the synthetic test panel;
prints the lines that differ.  Run on a box with two GPUs."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                      # noqa: E402
from oracle import jf_format            # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
meta = json.load(open(os.path.join(G, "synth_small.json")))
z = np.load(os.path.join(G, "synth_small.npz"))
d = tempfile.mkdtemp()
jf = os.path.join(d, "synth_small.jf")
jf_format.write_jf(jf, z["keys"], z["counts"])
files = []
for name, seq in list(zip(meta["names"], meta["targets"]))[:40]:
    fn = os.path.join(d, name + ".fa")
    open(fn, "w").write(">chrS:1-%d | name=%s\n%s\n" % (len(seq), name, seq))
    files.append(fn)


def run(extra):
    out = subprocess.run([sys.executable, "-m", "km_b200", "find_mutation", *extra, *files, jf], cwd=ROOT, capture_output=True, text=True)
    print("rc", out.returncode, "stderr tail:", out.stderr[-600:])
    return [l for l in out.stdout.split("\n") if l and not l.startswith("#Elapsed") and not l.startswith("#func:") and not l.startswith("#gpus:")]


one, two = run([]), run(["--gpus", "2"])
print(len(one), len(two))
n = 0
for i, (a, b) in enumerate(zip(one, two)):
    if a != b:
        fa, fb = a.split("\t"), b.split("\t")
        print("line", i, "fields that differ:", [(j, x[:60], y[:60]) for j, (x, y) in enumerate(zip(fa, fb)) if x != y], len(fa), len(fb))
        n += 1
        if n > 12:
            break
