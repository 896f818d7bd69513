"""Flags of `km find_cohort`: those of find_mutation (km/argparser/find_mutation.py:5-58), with the database
position taking any number of .jf files or directories of them."""
from .find_mutation import _INT_OPTS


def get_argparser_find_cohort(parser):
    def opt(short, long_, default, help_, typ):
        parser.add_argument(short, long_, help=help_, action="store", nargs="?", default=default, type=typ)

    opt(*_INT_OPTS[0], int)
    opt("-p", "--ratio", 0.05,
        "Minimum occurence ratio needed for exploration of alternative (default: -p 0.05)", float)
    for o in _INT_OPTS[1:]:
        opt(*o, int)
    parser.add_argument("-t", "--target", dest="target_fn", action="append", required=True,
                        help="Target sequence file or directory (repeatable).")
    parser.add_argument("--device", type=int, default=0, help="CUDA device [0].")
    parser.add_argument("--resident", action="store_true",
                        help="Load every database into HBM first and keep them all resident while the targets run "
                             "(default: one database at a time).")
    parser.add_argument("-f", "--format", dest="format", choices=["rows", "table"], default="rows",
                        help="rows: what the shell loop over `km find_mutation` prints [default]; table: per target, the "
                             "sample x variant matrix of `km find_report -f table` (find_report.py:290-327).")
    parser.add_argument("-m", "--min_cov", dest="min_cov", type=int, default=1,
                        help="With -f table: minimum coverage of a variant, as in find_report [1].")
    parser.add_argument("jellyfish_fn", nargs="+", help="Jellyfish databases (.jf files or directories of them).")
