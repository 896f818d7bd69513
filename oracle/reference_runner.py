"""ORACLE / TEST INFRASTRUCTURE -- drives the UNMODIFIED reference (km 2.2.2) over in-memory inputs.

Only bench.py's `--impl reference` / `cpu_baseline` legs and tests/ import this.  The reference's
Python comes from baseline/_ref/ (staged by tools/stage_reference.py; travels to the GPU box) or, in
the build container, from /root/reference.  Its one native dependency, Jellyfish, is absent from the
reference tree (pyproject.toml:10), so `import jellyfish` resolves to oracle/jellyfish_standin -- the
four calls of km/utils/Jellyfish.py:24-25,50-53 on top of oracle/kmer_store.c (pinned by the raw
counts of km/tests/test_main.py:581-652).  Everything above that boundary is the reference's own
code, running verbatim: `main_find_mut` (km/tools/find_mutation.py:17-60) is called the way the
reference's tests call it (km/tests/test_main.py:38-52), with an argparse.Namespace.
"""
import argparse
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def locate_reference():
    """Directory that holds the reference's `km` package, or None."""
    for top in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("KM_REFERENCE", "/root/reference")):
        if top and os.path.isfile(os.path.join(top, "km", "tools", "find_mutation.py")):
            return top
    return None


class ReferenceSession:
    """One process's handle on the reference: an in-memory k-mer store registered under a header-only
    .jf file in `workdir` (the reference opens the file itself to read `canonical`, Jellyfish.py:29-45),
    target FASTA files written there, and `find_mutation()` = one `main_find_mut` call."""

    def __init__(self, store, workdir, db_name="panel.jf", ref_root=None):
        ref_root = ref_root or locate_reference()
        if ref_root is None:
            raise RuntimeError("reference not staged: run tools/stage_reference.py where /root/reference exists")
        for p in (ref_root, os.path.join(HERE, "jellyfish_standin")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import jellyfish                                   # the stand-in
        from oracle import jf_format
        self.jellyfish = jellyfish
        self.workdir = workdir
        self.db_name = db_name
        os.makedirs(workdir, exist_ok=True)
        jf_format.write_jf(os.path.join(workdir, db_name), [], [], k=store.k, canonical=store.canonical)
        # the reference is handed the RELATIVE name (so the Database column reads like the GPU arm's);
        # find_mutation() runs with workdir as the current directory
        jellyfish.REGISTRY[db_name] = store
        from km.tools import find_mutation as fm           # the reference's own module
        self.fm = fm
        self.ref_root = ref_root

    def write_targets(self, names, seqs):
        files = []
        for name, seq in zip(names, seqs):
            fn = os.path.join(self.workdir, name + ".fa")
            with open(fn, "w") as f:
                f.write(">chrS:1-%d | name=%s\n%s\n" % (len(seq), name, seq))
            files.append(name + ".fa")
        return files

    def find_mutation(self, files, count=5, ratio=0.05, steps=500, branchs=10, nodes=10000):
        """Returns (rows: list of TSV lines without the '#' echo and the header, queries issued)."""
        args = argparse.Namespace(count=count, graphical=False, jellyfish_fn=self.db_name, ratio=ratio, steps=steps,
                                  branchs=branchs, nodes=nodes, target_fn=list(files), verbose=False, debug=False)
        q0 = self.jellyfish.QUERY_COUNT[0]
        old_out, old_cwd = sys.stdout, os.getcwd()
        buf = io.StringIO()
        os.chdir(self.workdir)
        sys.stdout = buf
        try:
            self.fm.main_find_mut(args, None)
        finally:
            sys.stdout = old_out
            os.chdir(old_cwd)
        rows = [ln for ln in buf.getvalue().split("\n") if ln and not ln.startswith("#") and not ln.startswith("Database\t")]
        return rows, self.jellyfish.QUERY_COUNT[0] - q0
