"""ORACLE / TEST INFRASTRUCTURE -- run the UNMODIFIED reference find_mutation in this
container (needs /root/reference, so never used on the GPU box).

usage: python oracle/run_reference.py [--raw] [--count N --ratio R --steps S --branchs B --nodes M] target.fa... db.jf

Default: prints what `km find_mutation` prints (km/tools/find_mutation.py:17-60), minus
the two volatile lines (#func:, #Elapsed time:).
--raw: drives the reference's own classes the way find_mutation.py:47-58 does and prints
one JSON object per target: rows, their unrounded floats, and the node set.
"""
import argparse
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KM_REFERENCE", "/root/reference")


def _paths():
    sys.path.insert(0, os.path.join(HERE, "jellyfish_standin"))
    sys.path.insert(0, REF)


def run(targets, jf, count=5, ratio=0.05, steps=500, branchs=10, nodes=10000):
    _paths()
    from km.tools import find_mutation as fm
    args = argparse.Namespace(count=count, graphical=False, jellyfish_fn=jf, ratio=ratio,
                              steps=steps, branchs=branchs, nodes=nodes, target_fn=list(targets),
                              verbose=False, debug=False)
    old = sys.stdout
    buf = io.StringIO()
    sys.stdout = buf
    try:
        fm.main_find_mut(args, None)
    finally:
        sys.stdout = old
    keep = [l for l in buf.getvalue().split("\n")
            if not l.startswith("#func:") and not l.startswith("#Elapsed time:")]
    return "\n".join(keep)


def run_raw(targets, jf_fn, count=5, ratio=0.05, steps=500, branchs=10, nodes=10000):
    _paths()
    from km.utils import MutationFinder as umf
    from km.utils import common as uc
    from km.utils import Sequence as us
    from km.utils.Jellyfish import Jellyfish
    jf = Jellyfish(jf_fn, cutoff=ratio, n_cutoff=count)
    out = []
    for seq_f in targets:
        name = os.path.splitext(os.path.basename(seq_f))[0]
        seqs, _ = uc.file_2_seq(seq_f)
        refpath = us.RefSeq("".join(seqs), name, jf.k)
        finder = umf.MutationFinder(refpath, jf, steps, branchs, nodes)
        finder.graph_analysis()
        finder.quantify_paths(False)
        finder.quantify_clusters(False)
        rows, raw = [], []
        for p in finder.get_paths(sort=True):
            rows.append(str(p))
            raw.append([float(p.rVAF), float(p.expression), float(p.ref_expression)])
        node = sorted((k, int(v)) for k, v in finder.node_data.items())
        alt = sorted(finder.get_seq(a.seq_index, skip_prefix=False) for a in finder.alt_paths)
        out.append({"target": name, "rows": rows, "raw": raw, "nodes": node, "alt_sequences": alt})
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw", action="store_true")
    ap.add_argument("--count", type=int, default=5)
    ap.add_argument("--ratio", type=float, default=0.05)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--branchs", type=int, default=10)
    ap.add_argument("--nodes", type=int, default=10000)
    ap.add_argument("files", nargs="+")
    a = ap.parse_args()
    if a.raw:
        for rec in run_raw(a.files[:-1], a.files[-1], a.count, a.ratio, a.steps, a.branchs, a.nodes):
            sys.stdout.write(json.dumps(rec) + "\n")
    else:
        sys.stdout.write(run(a.files[:-1], a.files[-1], a.count, a.ratio, a.steps, a.branchs, a.nodes))
