"""TEST INFRASTRUCTURE -- goldens for the wide-cluster solver, from the UNMODIFIED reference (needs /root/reference; run in
the build container, commit the output):

    python tests/golden/make_golden_wide.py      ->  tests/golden/wide_clusters.json

Cases: tests/helpers.wide_cluster_case (clusters of 2..5 substitutions: 3..6-column least-squares problems) and the
long-refinement targets of tests/test_gpu_round2.py (tandem duplications of 27-30 bases, panel seed + 5).  One record per
target, as oracle/run_reference.py --raw prints it: rows, unrounded floats, node set."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import wide_cluster_case          # noqa: E402
from km_b200 import synth                      # noqa: E402
from oracle import jf_format                   # noqa: E402

RUN = os.path.join(ROOT, "oracle", "run_reference.py")
ENV = dict(os.environ, PYTHONHASHSEED="0", PYTHONDONTWRITEBYTECODE="1")
WIDE = [(2, 100), (3, 100), (4, 100), (5, 100), (2, 300), (3, 300)]


def records(targets, names, keys, counts, db_label):
    with tempfile.TemporaryDirectory() as d:
        jf = os.path.join(d, db_label)
        jf_format.write_jf(jf, keys, counts)
        files = []
        for name, seq in zip(names, targets):
            fn = os.path.join(d, name + ".fa")
            with open(fn, "w") as f:
                f.write(">chrS:1-%d | name=%s\n%s\n" % (len(seq), name, seq))
            files.append(fn)
        out = subprocess.run([sys.executable, RUN, "--raw", *files, jf], check=True, cwd=d, capture_output=True, text=True,
                             env=ENV).stdout
        recs = [json.loads(l) for l in out.split("\n") if l]
    for r in recs:
        r["rows"] = [x.replace(jf, db_label) for x in r["rows"]]
    return recs


def main():
    out = {"wide": [], "long": None}
    for n, seed in WIDE:
        ref, keys, counts = wide_cluster_case(n, seed=seed)
        name = "wide_n%d_s%d" % (n, seed)
        out["wide"].append({"n": n, "seed": seed, "name": name, "record": records([ref], [name], keys, counts, "w.jf")[0]})
    panel = synth.make_panel(2000, seed=synth.PANEL_SEED + 5)
    picks = [i for i, tr in enumerate(panel.truth) if tr["kind"] == "dup" and tr.get("size") in (27, 28, 29, 30)][:12]
    recs = records([panel.targets[i] for i in picks], [panel.names[i] for i in picks], panel.keys, panel.counts, "p.jf")
    out["long"] = {"panel_seed_offset": 5, "n_targets": 2000, "picks": picks, "records": recs}
    with open(os.path.join(HERE, "wide_clusters.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "wide_clusters.json"), "-", len(out["wide"]), "wide cases,", len(picks), "long-refinement targets")


if __name__ == "__main__":
    main()
