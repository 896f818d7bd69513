"""ORACLE / TEST INFRASTRUCTURE -- row comparison rules shared by the parity tests.

BASELINE.json north_star: Type / Variant_name / Min_coverage / offsets / sequences are
bit-exact; rVAF, Expression and Reference_expression agree within 1e-6 relative on the
UNROUNDED floats.  The printed %.3f / %.1f cells (PathQuant.py:37-49) may then differ by
one unit in the last place only where the true value sits on a rounding boundary (the
reference itself flips such digits between PYTHONHASHSEEDs); those are counted, not
failed.  ``cluster <i>`` ids are not fixed by the reference when a target has several
clusters (set iteration order), so they are compared as '*'.
"""
import math
import re

EXACT_CELLS = (0, 1, 2, 3, 6, 7, 8, 10)   # db, query, type, name, min_cov, start_off, seq, ref_seq
FLOAT_CELLS = ((4, 0.001), (5, 0.1), (9, 0.1))
RTOL = 1e-6


def split_row(row):
    c = row.split("\t")
    if len(c) != 12:
        raise ValueError("row has %d cells: %r" % (len(c), row[:120]))
    return c


def key_of(cells):
    info = re.sub(r"cluster \d+ ", "cluster * ", cells[11])
    return tuple(cells[i] for i in EXACT_CELLS) + (info,)


def _close_print(a, b, ulp):
    if a == b:
        return True
    fa, fb = float(a), float(b)
    if math.isnan(fa) or math.isnan(fb):
        return math.isnan(fa) and math.isnan(fb)
    return abs(fa - fb) <= ulp * 1.0000001


def close_raw(x, y, rtol=RTOL):
    if math.isnan(x) or math.isnan(y):
        return math.isnan(x) and math.isnan(y)
    return abs(x - y) <= rtol * max(abs(x), abs(y)) + 1e-9


def compare_rows(want_rows, got_rows, want_raw=None, got_raw=None):
    """Returns (errors list, n_boundary_flips).  Rows are matched as multisets on their
    exact cells; numeric cells per the rules in the module docstring."""
    errors, flips = [], 0
    want = sorted(((key_of(split_row(r)), i) for i, r in enumerate(want_rows)))
    got = sorted(((key_of(split_row(r)), i) for i, r in enumerate(got_rows)))
    if [k for k, _ in want] != [k for k, _ in got]:
        wk, gk = [k for k, _ in want], [k for k, _ in got]
        for k in wk:
            if k not in gk:
                errors.append("missing row: %r" % (k[1:4] + k[-1:],))
        for k in gk:
            if k not in wk:
                errors.append("unexpected row: %r" % (k[1:4] + k[-1:],))
        if not errors:
            errors.append("row multiplicities differ")
        return errors, flips
    for (k, iw), (_, ig) in zip(want, got):
        cw, cg = split_row(want_rows[iw]), split_row(got_rows[ig])
        for cell, ulp in FLOAT_CELLS:
            if cw[cell] != cg[cell]:
                if _close_print(cw[cell], cg[cell], ulp):
                    flips += 1
                else:
                    errors.append("cell %d: %s != %s in %r" % (cell, cw[cell], cg[cell], k[1:4]))
        if want_raw is not None and got_raw is not None:
            for x, y in zip(want_raw[iw], got_raw[ig]):
                if not close_raw(float(x), float(y)):
                    errors.append("raw float %r vs %r in %r" % (x, y, k[1:4]))
    return errors, flips
