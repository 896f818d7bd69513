"""Stage the UNMODIFIED reference (km 2.2.2, /root/reference) under baseline/_ref/ so that
`bench.py --impl reference` can time the reference's own `main_find_mut` on the GPU box, where
/root/reference does not exist.  baseline/_ref/ is git-ignored (no
reference source enters this repository's history) but travels with the gpurun snapshot.

Two steps, outcome recorded in baseline/_ref/STAGED.json:
  1. the offline pip install the bench contract names (--no-index --no-build-isolation --no-deps,
     from a copy under /tmp because /root/reference is read-only).  It succeeds, but the reference's
     pyproject lists `packages = ["km"]` only, so the wheel holds km/__init__.py, __main__.py and km.py
     and NONE of the sub-packages (km/tools, km/utils, km/argparser): `import km.tools` fails.
  2. therefore the package tree is copied file for file from /root/reference/km (tests excluded).

Run from the repo root: python tools/stage_reference.py   (build() does it when /root/reference exists)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KM_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def _tree_digest(top):
    h = hashlib.sha256()
    for d, _dirs, files in sorted(os.walk(top)):
        if {"__pycache__", "tests"} & set(os.path.relpath(d, top).split(os.sep)):
            continue
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, top).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def stage(force=False):
    """Returns the STAGED.json record, or None when there is no reference to stage from."""
    marker = os.path.join(DST, "STAGED.json")
    if not os.path.isdir(os.path.join(REF, "km")):
        if os.path.exists(marker):
            with open(marker) as f:
                return json.load(f)
        return None
    want = _tree_digest(os.path.join(REF, "km"))
    if not force and os.path.exists(marker):
        with open(marker) as f:
            rec = json.load(f)
        if rec.get("source_sha256") == want and os.path.isdir(os.path.join(DST, "km", "tools")):
            return rec
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    tmp = "/tmp/km_ref_src_%d" % os.getpid()
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns("data", "example", "__pycache__"))
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                          "--find-links", "/opt/wheelhouse", "--target", DST, tmp],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    shutil.rmtree(tmp, ignore_errors=True)
    pip_ok = pip.returncode == 0
    wheel_has_subpackages = os.path.isdir(os.path.join(DST, "km", "tools"))
    # the wheel lacks the sub-packages: complete it with the source tree, byte for byte
    src = os.path.join(REF, "km")
    for d, dirs, files in os.walk(src):
        dirs[:] = [x for x in dirs if x not in ("__pycache__", "tests")]
        rel = os.path.relpath(d, src)
        os.makedirs(os.path.join(DST, "km", rel), exist_ok=True)
        for f in files:
            shutil.copyfile(os.path.join(d, f), os.path.join(DST, "km", rel, f))
    rec = {"reference": "iric-soft/km", "version": "2.2.2", "pip_install_ok": pip_ok,
           "pip_tail": pip.stdout.strip().splitlines()[-1:] if pip.stdout else [],
           "wheel_has_subpackages": wheel_has_subpackages,
           "completed_from_source_tree": True, "source_sha256": want,
           "staged_sha256": _tree_digest(os.path.join(DST, "km"))}
    with open(marker, "w") as f:
        json.dump(rec, f, indent=1)
    return rec


if __name__ == "__main__":
    print(json.dumps(stage(force="--force" in sys.argv), indent=1))
