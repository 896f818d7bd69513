"""Shared by the emulation (CPU) and GPU parity tests: turn a BatchResult-shaped object into the
comparable record the goldens use (rows, raw floats, node set, alt sequences)."""
from km_b200 import engine
from km_b200.utils import PathQuant as upq
from km_b200.utils import common as uc


def record_of(res, t, db_name, query_name):
    fields = res.row_fields(t, db_name, query_name)
    paths = [upq.Path(*f) for f in fields]
    paths = sorted(paths, key=lambda x: uc.natsortkey(*x[11].split(" "), x[1], x[3], x[2], x[6], rev_ix=[0]))
    kmers, counts = res.kmers(t), res.counts(t)
    alt = []
    for p in res.paths(t):
        alt.append(kmers[p[0]] + "".join(kmers[i][-1] for i in p[1:]))
    return {"rows": [str(p) for p in paths],
            "raw": [[float(p.rVAF), float(p.expression), float(p.ref_expression)] for p in paths],
            "nodes": sorted([k, int(c)] for k, c in zip(kmers[:-2], counts[:-2])),
            "alt_sequences": sorted(alt)}
