"""ORACLE / TEST INFRASTRUCTURE -- pin the oracle against the real reference.

Runs the UNMODIFIED reference (oracle/run_reference.py, several PYTHONHASHSEEDs) and
oracle/km_oracle.py on (a) every bundled target x sample pair and (b) a synthetic panel
with planted variants, and compares the TSV rows.  Needs /root/reference: build
container only.  Exit code 0 = identical (modulo ``cluster <i>`` renumbering, which the
reference itself does not fix -- see km_oracle's module docstring).

usage: python -m oracle.validate_vs_reference [--synthetic N] [--seeds 0,1,2]
"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import jf_format, km_oracle as ko   # noqa: E402
from oracle.store import KmerStore              # noqa: E402
from oracle.compare import compare_rows         # noqa: E402

REF = os.environ.get("KM_REFERENCE", "/root/reference")


def reference_records(targets, jf, hashseed, extra=()):
    env = dict(os.environ, PYTHONHASHSEED=str(hashseed), PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, os.path.join(HERE, "run_reference.py"), "--raw", *extra,
                          *targets, jf], check=True, capture_output=True, text=True, env=env).stdout
    return [json.loads(l) for l in out.split("\n") if l]


def oracle_records(targets, jf_path, walk="dfs", **kw):
    store = KmerStore.from_jf(jf_path)
    jf = ko.OracleJellyfish(store, jf_path, 0.05, 5)
    out = []
    for t in targets:
        tg = ko.Target.from_fasta(t, store.k)
        f = ko.OracleFinder(tg, jf, walk=walk, **kw).run()
        rows = f.get_paths(sort=True)
        out.append({"target": tg.name, "rows": [str(r) for r in rows],
                    "raw": [[float(r.rvaf), float(r.expr), float(r.ref_expr)] for r in rows],
                    "nodes": sorted((k, int(v)) for k, v in f.node_data.items()),
                    "alt_sequences": sorted(ko.spell(f.kmer, a, True) for a in f.alt_paths)})
    return out


FLIPS = [0]


def diff_records(tag, want, got):
    bad = 0
    for w, g in zip(want, got):
        errs = []
        if [list(x) for x in w["nodes"]] != [list(x) for x in g["nodes"]]:
            errs.append("node set / counts differ")
        if w["alt_sequences"] != g["alt_sequences"]:
            errs.append("alt path sequences differ")
        e, flips = compare_rows(w["rows"], g["rows"], w["raw"], g["raw"])
        FLIPS[0] += flips
        errs += e
        if errs:
            bad += 1
            print("MISMATCH", tag, w["target"])
            for x in errs[:8]:
                print("   ", x)
    return bad


def compare(tag, targets, jf, seeds):
    bad = 0
    want = reference_records(targets, jf, seeds[0])
    for hs in seeds[1:]:
        bad += diff_records(tag + " [reference seed %d vs %d]" % (seeds[0], hs), want,
                            reference_records(targets, jf, hs))
    for walk in ("dfs", "closure"):
        bad += diff_records(tag + " [oracle %s]" % walk, want, oracle_records(targets, jf, walk))
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", type=int, default=200)
    ap.add_argument("--seeds", default="0,1,7")
    ap.add_argument("--panel-seed", type=int, default=11)
    a = ap.parse_args()
    seeds = [int(s) for s in a.seeds.split(",")]
    bad = 0
    n = 0
    os.chdir(REF)   # the reference's tests use ./data/... paths
    for jf in sorted(glob.glob("./data/jf/*.jf")):
        for cat in ("GRCh38", "GRCh37"):
            tg = sorted(glob.glob("./data/catalog/%s/*.fa" % cat))
            bad += compare("%s x %s (all targets, one call)" % (cat, jf), tg, jf, seeds[:2])
            n += 1
            for t in tg:
                bad += compare("%s x %s" % (t, jf), [t], jf, seeds[:1])
                n += 1
    if a.synthetic:
        from km_b200 import synth
        panel = synth.make_panel(a.synthetic, seed=a.panel_seed, two_variant_frac=0.3)
        with tempfile.TemporaryDirectory() as d:
            jf = os.path.join(d, "synth.jf")
            jf_format.write_jf(jf, panel.keys, panel.counts)
            files = []
            for name, seq in zip(panel.names, panel.targets):
                fn = os.path.join(d, name + ".fa")
                with open(fn, "w") as f:
                    f.write(">chrS:1-%d | name=%s\n%s\n" % (len(seq), name, seq))
                files.append(fn)
            step = 25
            for i in range(0, len(files), step):
                bad += compare("synthetic[%d:%d]" % (i, i + step), files[i:i + step], jf, seeds)
                n += 1
    print("validated %d comparisons, %d mismatching, %d printed-digit boundary flips" % (n, bad, FLIPS[0]))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
