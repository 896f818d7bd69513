"""`km find_mutation <target(s)> <db.jf>` -- drop-in for km/tools/find_mutation.py.

Same stdout: the `#name:value` echo of every argument, the header, one sorted block of rows
per target in target order, `#Elapsed time:`.  The per-target loop of the reference
(find_mutation.py:47-58) becomes ONE km_find_batch launch over all targets; rows are formatted
by the library.
"""
import logging as log
import os
import sys
import time

from ..utils import MutationFinder as umf
from ..utils import Sequence as us
from ..utils import common as uc
from ..utils.Jellyfish import Jellyfish


def main_find_mut(args, argparser):
    time_start = time.time()
    if args.verbose:
        log.basicConfig(level=log.INFO, format="VERBOSE: %(message)s")
    if args.debug:
        log.basicConfig(level=log.DEBUG, format="VERBOSE: %(message)s")
    if getattr(args, "graphical", False):
        sys.exit("ERROR: -g/--graphical is not available in km_b200 (plots are outside the GPU path)")

    for name, value in vars(args).items():
        sys.stdout.write("#" + str(name) + ":" + str(value) + "\n")

    jf = Jellyfish(args.jellyfish_fn, cutoff=args.ratio, n_cutoff=args.count)
    umf.MutationFinder.output_header()

    refpaths = []
    for seq_f in uc.target_2_seqfiles(args.target_fn):
        ref_name = os.path.splitext(os.path.basename(seq_f))[0]
        ref_seqs, _ = uc.file_2_seq(seq_f)
        refpaths.append(us.RefSeq("".join(ref_seqs), ref_name, jf.k))    # multi-record FASTA is concatenated

    # rows of the targets before a failing one are still printed, then the error surfaces
    for finder in umf.find_batch(refpaths, jf, args.steps, args.branchs, args.nodes):
        sys.stdout.write(finder.format_rows())

    sys.stdout.write("#Elapsed time:" + str(time.time() - time_start) + "\n")
