"""`km linear_kmin` -- smallest k for which a target's k-mers are unique and form a linear
chain (km/tools/linear_kmin.py).  No database, no GPU: kept so the CLI is complete."""
import os
import sys
from collections import Counter

from ..utils import common as uc


def _is_linear(kmers):
    # every k-mer may have at most one other k-mer overlapping it on each side
    prefixes = Counter(km[:-1] for km in kmers)
    suffixes = Counter(km[1:] for km in kmers)
    for km in kmers:
        fwd = prefixes[km[1:]] - (1 if km[:-1] == km[1:] else 0)
        bwd = suffixes[km[:-1]] - (1 if km[:-1] == km[1:] else 0)
        if fwd > 1 or bwd > 1:
            return False
    return True


def find_kmin(ref_name, ref_seq, start):
    k_len = start - 1
    done = False
    while not done and k_len < len(ref_seq):
        k_len += 1
        try:
            kmers = uc.get_ref_kmer(ref_seq, ref_name, k_len)
        except ValueError:
            continue
        done = _is_linear(kmers)
    sys.stdout.write(ref_name + "\t" + str(k_len) + "\n")


def main_linear_kmin(args, argparser):
    sys.stdout.write("target_name\tlinear_kmin\n")
    for seq_f in uc.target_2_seqfiles(args.target_fn):
        ref_name = os.path.splitext(os.path.basename(seq_f))[0]
        seqs, _ = uc.file_2_seq(seq_f)
        find_kmin(ref_name, "".join(seqs), args.start)
