"""Target / FASTA helpers and the natural sort key -- host mirror of km/utils/common.py.
Same function names, arguments and error behaviour; nothing here touches the GPU except
get_cov, which goes through the batched probe kernel."""
import os
import re


def args_2_list_files(args):
    """km/utils/common.py:7-17: several arguments are taken as files; a single directory is
    expanded with os.listdir (unsorted, every entry)."""
    if len(args) == 1 and os.path.isdir(args[0]):
        return [os.path.join(args[0], name) for name in os.listdir(args[0])]
    return args


def target_2_seqfiles(target_fn):
    return args_2_list_files(target_fn)


def fasta_parser(fa_f):
    """Yield (header, sequence) per FASTA record (common.py:25-32); lines before the first
    header are ignored, a header without sequence lines is not supported (as upstream)."""
    header, parts = None, []
    with open(fa_f, "r") as handle:
        for line in handle:
            if line.startswith(">"):
                if header is not None:
                    yield header, "".join(parts)
                header, parts = line.strip(), []
            elif header is not None:
                parts.append(line.strip())
    if header is not None:
        yield header, "".join(parts)


def file_2_seq(seq_f):
    """common.py:35-45 -> ([sequence upper-cased, ...], [{attribute: value}, ...])."""
    sequences, attributes = [], []
    for header, sequence in fasta_parser(seq_f):
        fields = header.replace(">", "location=", 1).split("|")
        attr = {}
        for field in fields:
            key, value = field.split("=")
            attr[key.strip()] = value.strip()
        attributes.append(attr)
        sequences.append(sequence.upper())
    return sequences, attributes


def get_ref_kmer(ref_seq, ref_name, k_len):
    """common.py:48-63: the target's k-mers in order; a repeat raises ValueError."""
    first_seen = {}
    kmers = []
    for pos in range(len(ref_seq) - k_len + 1):
        kmer = ref_seq[pos:pos + k_len]
        if kmer in first_seen:
            raise ValueError("%s found multiple times in reference %s, at pos. %d" % (kmer, ref_name, pos))
        first_seen[kmer] = pos
        kmers.append(kmer)
    return kmers


def mean(v):
    return float(sum(v)) / len(v) if len(v) else 0


def get_cov(db, ref_seq):
    """common.py:73-92, with the per-k-mer loop replaced by one batched probe.
    -> (sum, len(ref_seq), min, max, mean, n_kmers, n_zero)."""
    from .Jellyfish import Jellyfish
    jf = Jellyfish(db)
    kmers = [ref_seq[i:i + jf.k] for i in range(len(ref_seq) - jf.k + 1)]
    counts = [int(c) for c in jf.query_many(kmers)]
    return (sum(counts), len(ref_seq), min(counts), max(counts), mean(counts), len(counts),
            sum(1 for c in counts if c == 0))


_DIGITS = re.compile("([0-9]+)")


class _Descending:
    """Inverts the order of whatever it wraps (natsortkey's rev_ix)."""
    __slots__ = ("obj",)

    def __init__(self, obj):
        self.obj = obj

    def __eq__(self, other):
        return self.obj == other.obj

    def __lt__(self, other):
        return other.obj < self.obj


def natsortkey(*args, rev_ix=[]):
    """Natural sort key (common.py:95-116): digit runs compare as integers, the rest
    case-insensitively; positions listed in rev_ix sort descending."""
    key = []
    for pos, word in enumerate(args):
        chunks = [int(c) if c.isdigit() else c.lower() for c in _DIGITS.split(word)]
        key.append(_Descending(chunks) if pos in rev_ix else chunks)
    return tuple(key)
