"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, through oracle/run_reference.py and the jellyfish stand-in) in the build
container.  The GPU box has no /root/reference; tests only read the files written here.

    python tests/golden/make_golden.py

Outputs
  bundled_inputs.npz   the five bundled .jf samples (header bytes + record bytes) and the
                       GRCh38/GRCh37 catalog FASTA text -- inputs of the reference's own tests
  bundled.json         per (catalog, target, sample): rows, unrounded floats, node set, alt
                       path sequences, as the reference computed them (PYTHONHASHSEED=0)
  bundled_cli.json     full stdout of `km find_mutation` (volatile lines dropped) and of
                       `km find_report` (default / -f vcf / -f table) for the five pairs the
                       reference tests use (km/tests/test_main.py)
  synth_small.npz/json a 96-target synthetic panel (km_b200.synth, seed 11) + reference records
"""
import glob
import io
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import jf_format   # noqa: E402
from km_b200 import synth      # noqa: E402

RUN = os.path.join(ROOT, "oracle", "run_reference.py")
ENV = dict(os.environ, PYTHONHASHSEED="0", PYTHONDONTWRITEBYTECODE="1")

PAIRS = [("NPM1_4ins_exons_10-11utr", "02H025_NPM1"), ("FLT3-ITD_exons_13-15", "03H112_IandI"),
         ("FLT3-ITD_exons_13-15", "03H116_ITD"), ("FLT3-TKD_exon_20", "05H094_FLT3-TKD_del"),
         ("DNMT3A_R882_exon_23", "02H033_DNMT3A_sub")]


def raw_records(targets, jf, cwd):
    out = subprocess.run([sys.executable, RUN, "--raw", *targets, jf], check=True, cwd=cwd,
                         capture_output=True, text=True, env=ENV).stdout
    return [json.loads(l) for l in out.split("\n") if l]


def find_report(fm_text, target, fmt=None, info="vs_ref"):
    code = (
        "import sys, io, argparse\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from km.tools import find_report as fr\n"
        "a = argparse.Namespace(target=%r, infile=io.StringIO(sys.stdin.read()), info=%r, min_cov=1, exclu='', format=%r)\n"
        "fr.main_find_report(a, None)\n" % (os.path.join(ROOT, "oracle", "jellyfish_standin"), REF, target, info, fmt))
    r = subprocess.run([sys.executable, "-c", code], input=fm_text, cwd=REF, capture_output=True,
                       text=True, env=ENV, check=True)
    return {"stdout": r.stdout, "stderr": r.stderr}


def main():
    inputs = {}
    for jf in sorted(glob.glob(os.path.join(REF, "data/jf/*.jf"))):
        name = os.path.basename(jf)[:-3]
        _, off = jf_format.read_header(jf)
        blob = np.fromfile(jf, dtype=np.uint8)
        inputs["jf_header__" + name] = blob[:off]
        inputs["jf_records__" + name] = blob[off:]
    fasta = {}
    for cat in ("GRCh38", "GRCh37"):
        for fa in sorted(glob.glob(os.path.join(REF, "data/catalog", cat, "*.fa"))):
            fasta["%s/%s" % (cat, os.path.basename(fa))] = open(fa).read()
    inputs["fasta_json"] = np.frombuffer(json.dumps(fasta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "bundled_inputs.npz"), **inputs)

    records = []
    for cat in ("GRCh38", "GRCh37"):
        targets = sorted(glob.glob(os.path.join(REF, "data/catalog", cat, "*.fa")))
        rel = ["./" + os.path.relpath(t, REF) for t in targets]
        for jf in sorted(glob.glob(os.path.join(REF, "data/jf/*.jf"))):
            jrel = "./" + os.path.relpath(jf, REF)
            for rec in raw_records(rel, jrel, REF):
                rec.update(catalog=cat, sample=os.path.basename(jf)[:-3])
                records.append(rec)
    with open(os.path.join(HERE, "bundled.json"), "w") as f:
        json.dump(records, f, separators=(",", ":"))

    cli = []
    for tname, sname in PAIRS:
        target = "./data/catalog/GRCh38/%s.fa" % tname
        jf = "./data/jf/%s.jf" % sname
        fm = subprocess.run([sys.executable, RUN, target, jf], check=True, cwd=REF,
                            capture_output=True, text=True, env=ENV).stdout
        cli.append({"target": tname, "sample": sname, "find_mutation": fm,
                    "report": find_report(fm, target), "report_vcf": find_report(fm, target, "vcf"),
                    "report_table": find_report(fm, target, "table"),
                    "report_cluster": find_report(fm, target, None, "cluster")})
    with open(os.path.join(HERE, "bundled_cli.json"), "w") as f:
        json.dump(cli, f, indent=0)

    panel = synth.make_panel(96, seed=11, two_variant_frac=0.3)
    with tempfile.TemporaryDirectory() as d:
        jf = os.path.join(d, "synth_small.jf")
        jf_format.write_jf(jf, panel.keys, panel.counts)
        files = []
        for name, seq in zip(panel.names, panel.targets):
            fn = os.path.join(d, name + ".fa")
            with open(fn, "w") as f:
                f.write(">chrS:1-%d | name=%s\n%s\n" % (len(seq), name, seq))
            files.append(fn)
        recs = raw_records(files, jf, d)
    for r in recs:
        r["rows"] = [x.replace(jf, "synth_small.jf") for x in r["rows"]]
    np.savez_compressed(os.path.join(HERE, "synth_small.npz"), keys=panel.keys, counts=panel.counts)
    with open(os.path.join(HERE, "synth_small.json"), "w") as f:
        json.dump({"names": panel.names, "targets": panel.targets, "truth": panel.truth,
                   "records": recs}, f, separators=(",", ":"))
    print("wrote goldens:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
