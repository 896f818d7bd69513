"""ORACLE / TEST INFRASTRUCTURE -- reader for Jellyfish ``binary/sorted`` .jf files.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (km_b200/) has its own loader in C++
(km_b200/csrc/jf_loader.cpp) and never touches oracle/.

The algorithm lives in a third-party dependency that is absent from
/root/reference: Jellyfish >= 2.2 (pyproject.toml:10 ``pyjellyfish>=1.3.0``;
CI built 2.2.6, .travis.yml:20-22).  The on-disk layout restated here is the
published binary/sorted dumper format, pinned on the five bundled files
 and by the raw counts asserted in the reference's
``test_min_cov`` (km/tests/test_main.py:581-652).

Layout:
  bytes 0..8   ASCII decimal H (zero padded) = length of the header blob
  bytes 9..9+H JSON header (+ NUL padding up to the 8-byte aligned record start)
  then         records of ceil(key_len/8) LE key bytes + counter_len LE count bytes
Key encoding: A=0 C=1 G=2 T=3, first base in the most significant bits.
"""
import json

import numpy as np

BASES = "ACGT"
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def read_header(path):
    """Return (header_dict, record_offset).  Brace matching mirrors what km does
    to find ``canonical`` (km/utils/Jellyfish.py:29-45)."""
    with open(path, "rb") as f:
        digits = f.read(9)
        hlen = int(digits.decode("ascii"))
        blob = f.read(hlen)
    text = blob.decode("ascii", errors="ignore")
    start = text.index("{")
    depth = 0
    end = None
    for i in range(start, len(text)):
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                end = i + 1
                break
    if end is None:
        raise ValueError("unterminated JSON header in %s" % path)
    return json.loads(text[start:end]), 9 + hlen


def read_jf(path):
    """Return (header, keys uint64[n], counts uint64[n])."""
    header, off = read_header(path)
    if header.get("format") != "binary/sorted":
        raise ValueError("unsupported .jf format %r" % header.get("format"))
    key_bytes = (int(header["key_len"]) + 7) // 8
    cnt_bytes = int(header["counter_len"])
    if key_bytes > 8 or cnt_bytes > 8:
        raise ValueError("key_len/counter_len too wide for this reader")
    raw = np.fromfile(path, dtype=np.uint8, offset=off)
    rec = key_bytes + cnt_bytes
    if raw.size % rec:
        raise ValueError("truncated .jf payload")
    raw = raw.reshape(-1, rec)
    keys = np.zeros(raw.shape[0], dtype=np.uint64)
    for b in range(key_bytes):
        keys |= raw[:, b].astype(np.uint64) << np.uint64(8 * b)
    counts = np.zeros(raw.shape[0], dtype=np.uint64)
    for b in range(cnt_bytes):
        counts |= raw[:, key_bytes + b].astype(np.uint64) << np.uint64(8 * b)
    return header, keys, counts


def pack(seq):
    """2-bit pack, first base most significant."""
    v = 0
    for c in seq:
        v = (v << 2) | _CODE[c]
    return v


def unpack(v, k):
    out = []
    for i in range(k):
        out.append(BASES[(v >> (2 * (k - 1 - i))) & 3])
    return "".join(out)


_COMP = str.maketrans("ACGT", "TGCA")


def revcomp(seq):
    return seq.translate(_COMP)[::-1]


def canonical_str(seq):
    rc = revcomp(seq)
    return rc if rc < seq else seq


def revcomp_packed(v, k):
    """Bitwise reverse complement of a 2-bit packed k-mer (k <= 32)."""
    v = (~v) & 0xFFFFFFFFFFFFFFFF
    v = ((v >> 2) & 0x3333333333333333) | ((v & 0x3333333333333333) << 2)
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0F) | ((v & 0x0F0F0F0F0F0F0F0F) << 4)
    v = ((v >> 8) & 0x00FF00FF00FF00FF) | ((v & 0x00FF00FF00FF00FF) << 8)
    v = ((v >> 16) & 0x0000FFFF0000FFFF) | ((v & 0x0000FFFF0000FFFF) << 16)
    v = ((v >> 32) | (v << 32)) & 0xFFFFFFFFFFFFFFFF
    return v >> (64 - 2 * k)


def write_jf(path, keys, counts, k=31, canonical=True, counter_len=4):
    """Write a ``binary/sorted``-layout file this repo's readers (and km's header
    parse, km/utils/Jellyfish.py:29-45) accept.  Records are NOT in Jellyfish's
    matrix1 hash order, so real Jellyfish would not binary-search it correctly --
    test fixture use only."""
    keys = np.asarray(keys, dtype=np.uint64)
    counts = np.asarray(counts, dtype=np.uint64)
    header = {"alignment": 8, "canonical": bool(canonical), "cmdline": ["km_b200-oracle"],
              "counter_len": int(counter_len), "format": "binary/sorted", "key_len": 2 * int(k),
              "max_reprobe": 126, "size": int(max(1, len(keys))), "val_len": 12}
    blob = json.dumps(header, separators=(",", ":")).encode("ascii")
    pad = (-(9 + len(blob))) % 8
    blob += b"\0" * pad
    key_bytes = (2 * int(k) + 7) // 8
    rec = np.zeros((len(keys), key_bytes + counter_len), dtype=np.uint8)
    for b in range(key_bytes):
        rec[:, b] = ((keys >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8)
    for b in range(counter_len):
        rec[:, key_bytes + b] = ((counts >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(("%09d" % len(blob)).encode("ascii"))
        f.write(blob)
        f.write(rec.tobytes())
