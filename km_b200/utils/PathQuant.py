"""Output row type -- host mirror of km/utils/PathQuant.py's ``Path``.

The numeric half of that file (class PathQuant: least squares + refinement) runs on the GPU
in km_b200/csrc/quant.h; ``PathQuant`` here is a thin holder for its results so code that
inspects ``quant.coef`` / ``quant.rVAF`` keeps working."""
import sys

COLUMNS = ("Database", "Query", "Type", "Variant_name", "rVAF", "Expression", "Min_coverage",
           "Start_offset", "Sequence", "Reference_expression", "Reference_sequence", "Info")


class Path:
    """One printable row (PathQuant.py:10-90)."""

    def __init__(self, db_f, ref_name, variant_name, ratio, expression, min_coverage, start_off, sequence,
                 ref_ratio, ref_expression, ref_sequence, note):
        self.db_name = db_f
        self.ref_name = ref_name
        self.variant_name = variant_name        # "Type\tstart:del/INS:end"
        self.rVAF = ratio
        self.expression = expression
        self.min_coverage = min_coverage
        self.start_off = start_off
        self.sequence = sequence
        self.ref_ratio = ref_ratio
        self.ref_expression = ref_expression
        self.ref_sequence = ref_sequence
        self.note = note

    def __str__(self):
        # PathQuant.py:37-49: 11 fields, the variant name itself holds a tab
        cells = (self.db_name, self.ref_name, self.variant_name, "%.3f" % self.rVAF, "%.1f" % self.expression,
                 "%d" % self.min_coverage, "%d" % self.start_off, self.sequence, "%.1f" % self.ref_expression,
                 self.ref_sequence, self.note)
        return "\t".join(cells)

    def __list__(self):
        return str(self).split("\t")

    def __getitem__(self, i):
        return self.__list__()[i]

    def get_min_cov(self):
        return self.min_coverage

    def get_sequence(self):
        return self.sequence

    def get_variant_name(self):
        return self.variant_name

    @staticmethod
    def output_header():
        sys.stdout.write("\t".join(COLUMNS) + "\n")


class PathQuant:
    """Results of one GPU least-squares problem: ``coef``, ``rVAF`` and the refine iteration
    count, in the column order the reference uses ([alt, ref] for vs_ref rows)."""

    def __init__(self, coef, rvaf, n_iter):
        self.coef = coef
        self.rVAF = rvaf
        self.n_iter = n_iter
