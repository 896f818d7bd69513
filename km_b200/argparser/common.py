import os


def is_valid_path(parser, p_file):
    full = os.path.abspath(p_file)
    if os.path.exists(full):
        return full
    parser.error("The path %s does not exist!" % full)


def is_valid_file(parser, n_file):
    if os.path.isfile(n_file):
        return n_file
    parser.error("The file %s does not exist!" % n_file)
