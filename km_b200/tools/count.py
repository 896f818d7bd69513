"""`km count` -- the k-mer database step of km's workflow on the GPU: reads -> canonical k-mer counts
(km_table_count_reads) -> lower-count filter -> a Jellyfish binary/sorted file (km_table_write_jf) that
`km find_mutation` (this one or the reference) opens.  It stands where `jellyfish count -m 31 -C -L 2 -Q+`
stands in example/run_leucegene.sh:22; the reference itself has no such sub-command."""
import gzip
import sys

import numpy as np

from .. import engine

_SUFFIX = {"k": 10 ** 3, "K": 10 ** 3, "m": 10 ** 6, "M": 10 ** 6, "g": 10 ** 9, "G": 10 ** 9}


def parse_size(text):
    text = text.strip()
    if text and text[-1] in _SUFFIX:
        return int(float(text[:-1]) * _SUFFIX[text[-1]])
    return int(text)


def _open(fn):
    if fn == "-":
        return sys.stdin.buffer
    if fn.endswith(".gz"):
        return gzip.open(fn, "rb")
    return open(fn, "rb")


def read_sequences(fn, min_qual=None):
    """Yields the sequences of a FASTA or FASTQ file as bytes; with min_qual (a byte value) every base
    whose quality character is below it becomes N (jellyfish count -Q)."""
    with _open(fn) as f:
        first = f.read(1)
        if not first:
            return
        if first == b">":
            parts = []
            f.readline()
            for line in f:
                if line.startswith(b">"):
                    if parts:
                        yield b"".join(parts)
                    parts = []
                else:
                    parts.append(line.strip())
            if parts:
                yield b"".join(parts)
        elif first == b"@":
            f.readline()
            while True:
                seq = f.readline().strip()
                plus = f.readline()
                qual = f.readline().strip()
                if not plus:
                    break
                if min_qual is not None and len(qual) == len(seq):
                    s = np.frombuffer(seq, dtype=np.uint8).copy()
                    s[np.frombuffer(qual, dtype=np.uint8) < min_qual] = ord("N")
                    seq = s.tobytes()
                yield seq
                if not f.readline():        # next header
                    break
        else:
            raise ValueError("%s: neither FASTA nor FASTQ" % fn)


def count_into(table, files, min_qual=None, batch_bases=64 << 20, native=True):
    """Counts the reads of `files` into `table`; returns (reads, bases).  native: the library's own reader
    (km_table_count_file); otherwise the files are parsed here and handed over in batches."""
    n_reads = n_bases = 0
    chunk, size = [], 0

    def flush():
        nonlocal chunk, size
        if chunk:
            off = np.zeros(len(chunk) + 1, dtype=np.int64)
            np.cumsum([len(r) for r in chunk], out=off[1:])
            table.count_reads_blob(b"".join(chunk), off)
            chunk, size = [], 0
    for fn in files:
        if native:
            flush()
            r, b = table.count_file(fn, min_qual)          # the library reads and parses the file itself
            n_reads += r
            n_bases += b
            continue
        for seq in read_sequences(fn, min_qual):
            chunk.append(seq)
            size += len(seq)
            n_reads += 1
            n_bases += len(seq)
            if size >= batch_bases:
                flush()
    flush()
    return n_reads, n_bases


def main_count(args, argparser):
    if not 1 <= args.mer_len <= 31:
        argparser.error("-m must be in 1..31")
    min_qual = ord(args.min_qual_char[0]) if args.min_qual_char else None
    table = engine.Table.create(k=args.mer_len, canonical=args.canonical, capacity=parse_size(args.size), device=args.device)
    n_reads, n_bases = count_into(table, args.files, min_qual)
    distinct = table.info()["n_keys"]
    if args.lower_count > 1:
        distinct = table.drop_below(args.lower_count)
    table.write_jf(args.output, counter_len=args.out_counter_len)
    sys.stderr.write("%d reads, %d bases, %d distinct %d-mers written to %s\n" % (n_reads, n_bases, distinct, args.mer_len, args.output))
