"""Flags of `km find_mutation` -- names, defaults and order are part of the drop-in contract
(km/argparser/find_mutation.py:5-58): every parsed value is echoed as `#name:value`."""

_INT_OPTS = (
    ("-c", "--count", 5, "Minimum occurence needed for exploration of alternative (default: -c 5)"),
    ("-s", "--steps", 500, "Maximum steps to discover a new branch on a target sequence (default: -s 500)"),
    ("-b", "--branchs", 10, "Maximum branchs until getback to target sequence (default: -b 10)"),
    ("-n", "--nodes", 10000, "Maximum nodes queried from jellyfish database (default: -n 5000)"),
)


def get_argparser_find_mut(parser):
    def opt(short, long_, default, help_, typ):
        parser.add_argument(short, long_, help=help_, action="store", nargs="?", default=default, type=typ)

    opt(*_INT_OPTS[0], int)
    opt("-p", "--ratio", 0.05,
        "Minimum occurence ratio needed for exploration of alternative (default: -p 0.05)", float)
    for o in _INT_OPTS[1:]:
        opt(*o, int)
    parser.add_argument("-g", "--graphical", help="Display coverage graph.", action="store_true")
    parser.add_argument("-v", "--verbose", help="Get more information.", action="store_true")
    parser.add_argument("-vv", "--debug", help="Get much more information.", action="store_true")
    # km_b200 only: the targets are dealt to this many GPUs of the box (the table is loaded on each); echoed as
    # `#gpus:N` only when it is not 1, so that the default output stays the reference's byte for byte
    parser.add_argument("--gpus", type=int, default=1, help="GPUs to spread the targets over (default: 1)")
    parser.add_argument("target_fn", help="Filename of the target sequence file or directory.", nargs="*")
    parser.add_argument("jellyfish_fn", help="Filename of the jellyfish database.")
