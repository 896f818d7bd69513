import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def bundled(tmp_path_factory):
    """Re-materialise the reference's bundled inputs (five .jf samples, two catalogs) from
    tests/golden/bundled_inputs.npz with the reference's relative layout: <root>/data/jf,
    <root>/data/catalog/<build>.  Returns the root directory."""
    root = tmp_path_factory.mktemp("bundled")
    z = np.load(os.path.join(GOLDEN, "bundled_inputs.npz"))
    os.makedirs(root / "data" / "jf")
    for key in z.files:
        if key.startswith("jf_header__"):
            name = key[len("jf_header__"):]
            with open(root / "data" / "jf" / (name + ".jf"), "wb") as f:
                f.write(z[key].tobytes())
                f.write(z["jf_records__" + name].tobytes())
    fasta = json.loads(z["fasta_json"].tobytes().decode())
    for rel, text in fasta.items():
        p = root / "data" / "catalog" / rel
        os.makedirs(p.parent, exist_ok=True)
        p.write_text(text)
    return str(root)


@pytest.fixture(scope="session")
def bundled_golden():
    with open(os.path.join(GOLDEN, "bundled.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def bundled_cli_golden():
    with open(os.path.join(GOLDEN, "bundled_cli.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def synth_small(tmp_path_factory):
    """The 96-target synthetic panel + the reference's records for it."""
    from oracle import jf_format
    with open(os.path.join(GOLDEN, "synth_small.json")) as f:
        meta = json.load(f)
    z = np.load(os.path.join(GOLDEN, "synth_small.npz"))
    d = tmp_path_factory.mktemp("synth_small")
    jf = str(d / "synth_small.jf")
    jf_format.write_jf(jf, z["keys"], z["counts"])
    files = []
    for name, seq in zip(meta["names"], meta["targets"]):
        fn = str(d / (name + ".fa"))
        with open(fn, "w") as f:
            f.write(">chrS:1-%d | name=%s\n%s\n" % (len(seq), name, seq))
        files.append(fn)
    meta.update(jf=jf, files=files, keys=z["keys"], counts=z["counts"], dir=str(d))
    return meta
