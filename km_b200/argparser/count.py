def get_argparser_count(parser):
    """The flags km's workflow passes to `jellyfish count` (example/run_leucegene.sh:22)."""
    parser.add_argument("files", nargs="+", help="FASTA / FASTQ files (plain or .gz), or - for stdin.")
    parser.add_argument("-m", "--mer-len", type=int, default=31, help="k-mer length (<= 31) [31].")
    parser.add_argument("-s", "--size", default="100M", help="Distinct k-mers to make room for (k, M, G suffixes) [100M].")
    parser.add_argument("-C", "--canonical", action="store_true", help="Count both strands, canonical representation.")
    parser.add_argument("-L", "--lower-count", type=int, default=0, help="Do not write k-mers with a count below this.")
    parser.add_argument("-Q", "--min-qual-char", default="", help="Bases with a quality character below this one are treated as N.")
    parser.add_argument("-c", "--counter-len", type=int, default=0, help="Accepted for compatibility (hash counter bits); ignored.")
    parser.add_argument("-t", "--threads", type=int, default=1, help="Accepted for compatibility; ignored.")
    parser.add_argument("--disk", action="store_true", help="Accepted for compatibility; ignored.")
    parser.add_argument("--out-counter-len", type=int, default=4, help="Bytes per count in the output file [4].")
    parser.add_argument("-o", "--output", default="mer_counts.jf", help="Output file [mer_counts.jf].")
    parser.add_argument("--device", type=int, default=0, help="CUDA device [0].")
