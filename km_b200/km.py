"""Command line entry point: `km find_mutation | find_report | linear_kmin | min_cov`
(same sub-commands and dispatch as km/km.py:17-67), plus `count`, which stands in for the
`jellyfish count` step of km's workflow (example/run_leucegene.sh:22), and `find_cohort`, the
shell loop over samples (run_leucegene.sh:29-35) as one invocation."""
import argparse
import sys

from .argparser.count import get_argparser_count
from .argparser.find_cohort import get_argparser_find_cohort
from .argparser.find_mutation import get_argparser_find_mut
from .argparser.find_report import get_argparser_find_report
from .argparser.linear_kmin import get_argparser_linear_kmin
from .argparser.min_cov import get_argparser_min_cov
from .tools.count import main_count
from .tools.find_cohort import main_find_cohort
from .tools.find_mutation import main_find_mut
from .tools.find_report import main_find_report
from .tools.linear_kmin import main_linear_kmin
from .tools.min_cov import main_min_cov

COMMANDS = (
    ("find_mutation", "Identify and quantify mutations from a target sequence and a k-mer database.",
     main_find_mut, get_argparser_find_mut),
    ("find_report", "Parse find_mutation output and reformat it in a more user-friendly tabulated file.",
     main_find_report, get_argparser_find_report),
    ("linear_kmin", "Find min k-length to decompose a target sequence in a linear graph.",
     main_linear_kmin, get_argparser_linear_kmin),
    ("min_cov", "Compute coverage of target sequences.", main_min_cov, get_argparser_min_cov),
    ("count", "Count k-mers of FASTA/FASTQ reads on the GPU into a Jellyfish binary/sorted database.",
     main_count, get_argparser_count),
    ("find_cohort", "find_mutation over many databases in one invocation (targets packed once, one batched call per sample).",
     main_find_cohort, get_argparser_find_cohort),
)


def main():
    parser = argparse.ArgumentParser(prog="PROG")
    subparsers = parser.add_subparsers(help="sub-command help")
    for name, text, func, add_flags in COMMANDS:
        sub = subparsers.add_parser(name, help=text)
        sub.set_defaults(func=func)
        add_flags(sub)
    if len(sys.argv) == 1:
        parser.print_help(sys.stderr)
        sys.exit(1)
    args = parser.parse_args()
    args.func(args, parser)
