"""`km find_cohort -t <target(s)> <db.jf> [<db.jf> ...]` -- one invocation over MANY samples.

km's users loop `km find_mutation` over their .jf files in the shell (example/run_leucegene.sh:29-35) and
concatenate; here the targets are read and packed once, every database is loaded into HBM in turn, and each
sample is ONE km_find_text call (text formatted on the device).  The output is what that loop prints: one header,
then for every sample the rows of every target -- the Database column tells the samples apart, which is what
`km find_report -f table` aggregates (find_report.py:290-327).  The reference has no such sub-command."""
import os
import sys
import time

from .. import engine
from ..utils import MutationFinder as umf
from ..utils import Sequence as us
from ..utils import common as uc


def _jf_files(args):
    out = []
    for p in args:
        if os.path.isdir(p):
            out.extend(sorted(os.path.join(p, f) for f in os.listdir(p) if f.endswith(".jf")))
        else:
            out.append(p)
    return out


def _cohort_table(seq_files, names, header, rows_by_query, min_cov):
    """Per target: `km find_report -t <target> -f table` (find_report.py:290-327) over the rows of ALL samples -- one
    line per sample, one column per variant, the cells the variants' rVAF."""
    import argparse
    import contextlib
    import io
    from . import find_report as fr
    for seq_f, name in zip(seq_files, names):
        rows = rows_by_query.get(name, [])
        sys.stdout.write("Target\t%s\n" % name)
        if not rows:
            continue
        rep = argparse.Namespace(target=seq_f, infile=io.StringIO(header + "".join(r + "\n" for r in rows)), info="vs_ref",
                                 min_cov=min_cov, exclu="", format="table")
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            fr.create_report(rep)
        sys.stdout.write(buf.getvalue())


def main_find_cohort(args, argparser):
    time_start = time.time()
    as_table = getattr(args, "format", "rows") == "table"
    for name, value in vars(args).items():
        sys.stdout.write("#" + str(name) + ":" + str(value) + "\n")
    if not as_table:
        umf.MutationFinder.output_header()
    seq_files = []
    for tgt in args.target_fn:
        seq_files.extend(uc.target_2_seqfiles([tgt]))
    jf_files = _jf_files(args.jellyfish_fn)
    # --resident: every sample's table in HBM at once (a cohort of real samples is ~7 GB each, example/README.rst:47-48:
    # two dozen fit one B200); otherwise one at a time
    tables = [engine.Table.open_jf(fn, device=args.device) for fn in jf_files] if getattr(args, "resident", False) else None
    packed, k_seen = None, None
    rows_by_query = {}
    for ix, jf_fn in enumerate(jf_files):
        table = tables[ix] if tables is not None else engine.Table.open_jf(jf_fn, device=args.device)
        if packed is None or table.k != k_seen:
            names, seqs = [], []
            for seq_f in seq_files:
                ref_seqs, _ = uc.file_2_seq(seq_f)
                ref = us.RefSeq("".join(ref_seqs), os.path.splitext(os.path.basename(seq_f))[0], table.k)   # raises on a non-linear target
                names.append(ref.name)
                seqs.append(ref.seq)
            packed, k_seen = engine.PackedTargets(seqs, names), table.k
        text, status = table.find_text(packed, jf_fn, count=args.count, ratio=args.ratio, steps=args.steps,
                                       branchs=args.branchs, nodes=args.nodes)
        if as_table:
            for ln in text.split("\n"):
                if ln:
                    rows_by_query.setdefault(ln.split("\t")[1], []).append(ln)
        else:
            sys.stdout.write(text)
        for i, st in enumerate(status.tolist()):
            engine.raise_for_status(st, packed.names[i], args.nodes)
        if tables is None:
            table.close()
    if as_table:
        _cohort_table(seq_files, packed.names if packed else [], "", rows_by_query, getattr(args, "min_cov", 1))
    if tables is not None:
        for t in tables:
            t.close()
    sys.stdout.write("#Elapsed time:" + str(time.time() - time_start) + "\n")
