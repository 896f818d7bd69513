"""`km min_cov <target> <db.jf>...` -- drop-in for km/tools/min_cov.py; the k-mer lookups of each
database go through the batched GPU probe (common.get_cov)."""
import os
import sys

from ..utils import common as uc

HEADER = ("DB", "count", "length", "min", "max", "mean", "kmer_nb", "kmer_nb_0")


def main_min_cov(args, argparser):
    ref_seq = args.target_fn
    if os.path.isfile(args.target_fn):
        seqs, _ = uc.file_2_seq(args.target_fn)
        ref_seq = "".join(seqs)
    sys.stdout.write("\t".join(HEADER) + "\n")
    for jf_file in uc.args_2_list_files(args.jellyfish_fn):
        total, length, lo, hi, avg, n_kmer, n_zero = uc.get_cov(jf_file, ref_seq)
        sys.stdout.write("%s\t%d\t%d\t%d\t%d\t%.2f\t%d\t%d\n" % (jf_file, total, length, lo, hi, avg, n_kmer, n_zero))
