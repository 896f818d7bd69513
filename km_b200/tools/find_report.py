"""`km find_report -t target.fa [find_mutation output]` -- drop-in for km/tools/find_report.py.

Pure host-side post-processing of the handful of rows find_mutation prints (out
of scope for GPU work); the only database access is the optional `-e` exclusion table, which goes
through the batched GPU probe (utils.common.get_cov).  Output formats (default tabular, `-f vcf`,
`-f table`) and the insertion re-typing rules (ITD / I&I) follow the reference; each block cites
the lines it restates.
"""
import re
import sys

from ..utils import common as uc

REPORT_COLUMNS = ("Sample", "Region", "Location", "Type", "Removed", "Added", "Abnormal", "Normal", "rVAF",
                  "Min_coverage", "Exclu_min_cov", "Variant", "Target", "Info", "Variant_sequence",
                  "Reference_sequence")

VCF_HEADER = (
    '##fileformat=VCFv4.1\n'
    '##INFO=<ID=TYPE,Number=A,Type=String,Description='
    '"The type of variant, either Insertion, ITD, I&I, Deletion, Substitution or Indel.">\n'
    '##INFO=<ID=TARGET,Number=A,Type=String,Description='
    '"Name of the sequencing that contains the mutation.">\n'
    '##INFO=<ID=RATIO,Number=A,Type=String,Description="Ratio of mutation to reference.">\n'
    '##INFO=<ID=MINCOV,Number=A,Type=String,Description='
    '"Minimum k-mer coverage of alternative allele.">\n'
    '##INFO=<ID=REMOVED,Number=A,Type=String,Description="Number of removed bases.">\n'
    '##INFO=<ID=ADDED,Number=A,Type=String,Description="Number of added bases.">\n'
    '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n')

_COMPLEMENT = str.maketrans("ATGCU", "TACGA")


def print_line(*cells):
    sys.stdout.write("\t".join(cells) + "\n")


def print_vcf_header():
    sys.stdout.write(VCF_HEADER)


def print_vcf_line(chro, loc, ref_var, alt_var, type_var, target, ratio, min_cov, rem, ad):
    info = "TYPE=%s;TARGET=%s;RATIO=%s;MINCOV=%s;REMOVED=%s;ADDED=%s" % (type_var, target, ratio, min_cov, rem, ad)
    print_line(chro, str(loc), ".", ref_var, alt_var, ".", ".", info)


def init_ref_seq(arg_ref):
    """Genome coordinate of every base of the target, from the `>chr:start-end | strand=..`
    headers (find_report.py:38-76).  -> (coordinates, chromosome, strand)."""
    if not arg_ref:
        sys.exit("ERROR: Target file is empty\n")
    coords = []
    chro = strand = None
    with open(arg_ref, "r") as handle:
        for raw in handle:
            line = raw.strip()
            if not line.startswith(">"):
                continue
            loc = line.split(" ")[0]
            if ":" not in loc or "-" not in loc:
                sys.exit("ERROR: Fasta entries do not contain a correctly " +
                         "formatted location: {}\n".format(loc))
            attr = {}
            for field in line.replace(">", "location=", 1).split("|"):
                key, value = field.split("=")[0:2]
                attr[key.strip()] = value.strip()
            chro, span = attr["location"].split(":")
            first, last = span.split("-")
            if "strand" not in attr:
                attr["strand"] = "+"
                sys.stderr.write("WARNING: Strand is assumed to be '+' \n")
            strand = attr["strand"]
            exon = list(range(int(first), int(last) + 1))
            coords.extend(reversed(exon) if strand == "-" else exon)
    return coords, chro, strand


def _slide_left(variant, pos, ref):
    """find_report.py:84-87 (get_extremities), iteratively: how far a repeat-compatible variant
    can be shifted towards the start of `ref`.  Returns the index before the leftmost start."""
    while pos - 1 > 0 and ref[pos - 1] == variant[-1]:
        variant = ref[pos - 1] + variant[:-1]
        pos -= 1
    return pos - 1


class _Coords:
    def __init__(self, nts, chro, strand):
        self.nts, self.chro, self.strand = nts, chro, strand

    def span(self, pos, end):
        """Genome (start, end) of target offsets pos..end, strand aware (find_report.py:174-179)."""
        if self.strand == "+":
            return self.nts[pos], self.nts[end]
        if self.strand == "-":
            return self.nts[end], self.nts[pos]
        raise UnboundLocalError("strand %r" % self.strand)

    def whole(self):
        lo, hi = (self.nts[-1], self.nts[0]) if self.strand == "-" else (self.nts[0], self.nts[-1])
        return "{}:{}-{}".format(self.chro, lo, hi)


def create_report(args):
    if args.format == "vcf" and args.info == "cluster":
        sys.exit("ERROR: -f vcf and -i cluster options are incompatible")
    vcf = args.format == "vcf"
    table = args.format == "table"
    nts, chro, strand = init_ref_seq(args.target)
    geo = _Coords(nts, chro, strand)

    if vcf:
        print_vcf_header()
    elif not table:
        print_line(*REPORT_COLUMNS)

    tally, per_sample, ratios = {}, {}, {}          # -f table accumulators

    for line in args.infile:
        if line[0] == "#":
            continue
        tok = line.strip("\n").split("\t")
        if not re.search(args.info, line) or tok[0] == "Database" or len(tok) <= 1:
            continue
        samp, query, kind, name = tok[0], tok[1], tok[2], tok[3]
        ratio, alt_exp, min_cov, start_off = tok[4], tok[5], tok[6], tok[7]
        alt_seq, ref_exp, ref_seq_raw, info = tok[8], tok[9], tok[10], tok[11]
        ref_seq = ref_seq_raw.upper()

        min_exclu = ""
        if args.exclu != "" and alt_seq != "":
            min_exclu = str(uc.get_cov(args.exclu, alt_seq)[2])          # :137-139
        if int(min_cov) < args.min_cov:                                  # :141-142
            continue

        if kind == "Reference":                                          # :145-158
            mod = ""
            region = geo.whole()
            if vcf:
                continue
            if not table:
                print_line(samp, region, "-", kind, "0", "0", "0.0", alt_exp, tok[4], min_cov, min_exclu, "-",
                           query, tok[-1], "", "")
                continue
        else:
            start, mod, stop = name.split(":")
            delet, insert = mod.split("/")
            added, removed = str(len(insert)), str(len(delet))
            # 0-based offsets inside the printed (possibly clipped) sequence   (:169-173)
            pos = int(start) - 1 - int(start_off)
            end = int(stop) - 2 - int(start_off)
            start_pos, end_pos = geo.span(pos, end)
            region = "{}:{}-{}".format(chro, start_pos, end_pos + 1)
            ref_var, alt_var = delet.upper(), insert.upper()
            loc_var, end_var = start_pos, end_pos

            if not delet and insert:                                     # pure insertion (:188-227)
                start_pos, end_pos = geo.span(pos, end + 1)              # insertions end at the last position
                region = "{}:{}-{}".format(chro, start_pos, end_pos + 1)
                var = insert.upper()
                ibef = _slide_left(var, pos, ref_seq)
                before = ref_seq[ibef:pos]
                rev = ref_seq[::-1]
                iaft = _slide_left(var[::-1], len(ref_seq) - pos, rev)
                after = rev[iaft:len(ref_seq) - pos][::-1]
                iaft = len(ref_seq) - iaft - 1
                ref_var = before + after
                alt_var = before + var + after
                loc_var = nts[iaft] if strand == "-" else nts[ibef]
                end_var = nts[iaft - len(ref_var) + 1] if strand == "-" else nts[ibef + len(ref_var) - 1]
                if loc_var + len(ref_var) - 1 != end_var and vcf:
                    sys.stderr.write("NOTE: Mutation overlaps 2 exons or more, VCF output is disabled \n")
                    continue
                # re-type small insertions: identical to what precedes -> ITD, mostly so -> I&I
                upstream = alt_seq[pos - len(insert):pos]
                inside = pos - len(insert) >= 0
                match = 0
                if inside:
                    match = float(sum(1 for a, b in zip(insert, upstream) if a == b)) / len(insert)
                insert_type = "Insertion"
                if inside and len(insert) >= 3 and insert == upstream:
                    insert_type = "ITD"
                    added += " | " + str(end_pos - start_pos + 1)
                elif inside and len(insert) >= 3 and match > 0.5:
                    insert_type = "I&I"
                    added += " | " + str(end_pos - start_pos + 1)
                location = chro + ":" + str(end_pos)

            elif kind == "Deletion":                                     # :229-247
                location = ""
                insert_type = kind
                var = delet.upper()
                ibef = _slide_left(var, pos, ref_seq)
                before = ref_seq[ibef:pos]
                rev = ref_seq[::-1]
                tail = len(ref_seq) - pos - 1 - len(var) + 1
                iaft = _slide_left(var[::-1], tail, rev)
                after = rev[iaft:tail][::-1]
                iaft = len(ref_seq) - iaft - 1
                ref_var = before + var + after
                alt_var = before + after
                loc_var = nts[iaft] if strand == "-" else nts[ibef]
                end_var = nts[iaft - len(ref_var) + 1] if strand == "-" else nts[ibef + len(ref_var) - 1]
                if loc_var + len(ref_var) - 1 != end_var and vcf:
                    continue

            elif kind == "Substitution":                                 # :249-255
                location = chro + ":" + str(start_pos)
                insert_type = kind
                if loc_var + len(ref_var) - 1 != end_var and vcf:
                    sys.stderr.write("NOTE: Mutation overlaps 2 exons or more, VCF output is disabled \n")
                    continue

            elif kind == "Indel":                                        # :257-268
                location = chro + ":" + str(end_pos)
                insert_type = kind
                ref_var = ref_seq[pos - 1] + delet.upper() + ref_seq[end + 1]
                alt_var = ref_seq[pos - 1] + insert.upper() + ref_seq[end + 1]
                loc_var, end_var = start_pos - 1, end_pos + 1
                if loc_var + len(ref_var) - 1 != end_var and vcf:
                    sys.stderr.write("NOTE: Mutation overlaps 2 exons or more, VCF output is disabled \n")
                    continue

            else:                                                        # :270-275
                sys.stderr.write("WARNING: This variant isn't taken account\n")
                sys.stderr.write(" - variant: " + str(kind) + "\n")
                sys.stderr.write(" - line: " + line)
                sys.exit()

        if vcf:                                                          # :283-289
            if strand == "-":
                ref_var = ref_var.translate(_COMPLEMENT)[::-1]
                alt_var = alt_var.translate(_COMPLEMENT)[::-1]
            print_vcf_line(chro, loc_var, ref_var, alt_var, insert_type, query, ratio, min_cov, removed,
                           added.replace(" ", ""))
        elif table:                                                      # :291-306
            var_name = kind if "/" in kind else kind + "/" + query
            key = (var_name, region + ":" + mod if mod else region)
            tally[key] = tally.get(key, 0) + 1
            per_sample.setdefault(samp, set()).add(key)
            ratios.setdefault(samp, {})[key] = float(ratio)
        else:                                                            # :277-281
            print_line(samp, region, location, insert_type, removed, added, alt_exp, ref_exp, ratio, min_cov,
                       min_exclu, mod, query, info, alt_seq, ref_seq_raw)

    if table:                                                            # :308-327
        ordered = sorted(tally, key=tally.get, reverse=True)
        sys.stdout.write("Sample")
        for v in ordered:
            sys.stdout.write("\t" + (v[0] if v[0].split("/")[0] == "Reference" else v[1]))
        sys.stdout.write("\n")
        for samp, seen in per_sample.items():
            sys.stdout.write(samp)
            for v in ordered:
                if v in seen and ("Reference" in v[0] or ratios[samp][v]):
                    sys.stdout.write("\t" + str(ratios[samp][v]))
                else:
                    sys.stdout.write("\t" + ".")
            sys.stdout.write("\n")


def main_find_report(args, argparser):
    if args.infile.isatty() or args.target is None:
        argparser.print_help()
        sys.exit()
    create_report(args)
