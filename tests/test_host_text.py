"""Host-side text path of libkm_b200.so (no GPU): number formatting and natural sorting must be
what Python's '%.3f' / '%.1f' and km.utils.common.natsortkey (common.py:95-116) produce."""
import ctypes
import random
import re
import struct

import pytest

from km_b200 import build as kb
from km_b200._lib import lib


@pytest.fixture(scope="module")
def L():
    kb.build()
    return lib()


def fmt(L, v, prec):
    buf = ctypes.create_string_buffer(512)
    n = L.km_debug_format_fixed(float(v), prec, buf)
    assert n >= 0
    return buf.value.decode()


def test_fixed_point_printing_equals_python(L):
    rng = random.Random(11)
    vals = [0.0, -0.0, 0.0005, 0.0015, 0.0025, 0.3625, 0.4845, 2870.598870056498, 3055.1525423728817, 0.05, 0.25, 0.35,
            1e-9, 5e-4, 4.999999999e-4, 0.9995, 0.99949999, 1234567.25, 2.5, 3.5, -1.0, 1e15, 4503599627370495.5, 1e18,
            float("nan"), float("inf"), -float("inf"), 5e-324, 2.2250738585072014e-308]
    for _ in range(3000):
        vals.append(rng.random() * 10 ** rng.randint(-6, 7))
        vals.append(rng.randint(0, 10 ** 7) / 2000.0)             # many exact ties at the third digit
        vals.append(struct.unpack("<d", struct.pack("<Q", rng.getrandbits(62)))[0])
    for v in vals:
        for prec in (1, 3):
            assert fmt(L, v, prec) == "%.*f" % (prec, v), (v, prec)


def natkey(s):
    return [int(c) if c.isdigit() else c.lower() for c in re.split("([0-9]+)", s)]


def test_natural_compare_equals_reference_key(L):
    rng = random.Random(5)
    alphabet = "0123456789acgtACGT:/="
    words = ["", "Reference", "45:/TCTG:45", "45:/tctg:45", "204:/ACG:204", "32:gat/:35", "33:c/T:34", "n=2", "n=10", "007", "7", "70",
             "vs_ref", "cluster", "a1b", "a01b", "a1"]
    for _ in range(3000):
        words.append("".join(rng.choice(alphabet) for _ in range(rng.randint(0, 9))))
    for _ in range(6000):
        a, b = rng.choice(words), rng.choice(words)
        ka, kb_ = natkey(a), natkey(b)
        want = -1 if ka < kb_ else (1 if ka > kb_ else 0)
        assert L.km_debug_nat_cmp(a.encode(), b.encode()) == want, (a, b)
