"""find_report (host-only post-processing, km/tools/find_report.py) against the outputs of the
unmodified reference on the five pairs its own tests use: default, -f vcf, -f table, -i cluster."""
import io
import os
import sys
from argparse import Namespace
from contextlib import contextmanager

import pytest

from km_b200.tools import find_report as fr


@contextmanager
def captured():
    out, err = io.StringIO(), io.StringIO()
    old = sys.stdout, sys.stderr
    sys.stdout, sys.stderr = out, err
    try:
        yield out, err
    finally:
        sys.stdout, sys.stderr = old


def run_report(bundled, fm_text, target, fmt=None, info="vs_ref"):
    args = Namespace(target=os.path.join(bundled, "data/catalog/GRCh38", target + ".fa"),
                     infile=io.StringIO(fm_text), info=info, min_cov=1, exclu="", format=fmt)
    with captured() as (out, err):
        fr.main_find_report(args, None)
    return out.getvalue(), err.getvalue()


def test_reports_equal_reference_output(bundled, bundled_cli_golden):
    for case in bundled_cli_golden:
        for key, fmt, info in (("report", None, "vs_ref"), ("report_vcf", "vcf", "vs_ref"),
                               ("report_table", "table", "vs_ref"), ("report_cluster", None, "cluster")):
            out, err = run_report(bundled, case["find_mutation"], case["target"], fmt, info)
            assert out == case[key]["stdout"], (case["target"], case["sample"], key)
            assert err == case[key]["stderr"], (case["target"], case["sample"], key)


def test_known_answers_of_reference_tests(bundled, bundled_cli_golden):
    # km/tests/test_main.py:73-139 (NPM1), :417-448 (FLT3-TKD vcf)
    npm1 = [c for c in bundled_cli_golden if c["sample"] == "02H025_NPM1"][0]
    out, _ = run_report(bundled, npm1["find_mutation"], npm1["target"])
    row = out.split("\n")[2].split("\t")
    assert row[2:10] == ["chr5:171410544", "ITD", "0", "4 | 1", "2870.6", "3055.2", "0.484", "2428"]
    assert row[11] == "/TCTG"
    out, _ = run_report(bundled, npm1["find_mutation"], npm1["target"], "vcf")
    v = [l for l in out.split("\n") if l and l[0] != "#"][0].split("\t")
    assert (v[1], v[3], v[4]) == ("171410539", "CTCTGG", "CTCTGTCTGG")
    tkd = [c for c in bundled_cli_golden if c["sample"] == "05H094_FLT3-TKD_del"][0]
    out, _ = run_report(bundled, tkd["find_mutation"], tkd["target"], "vcf")
    v = [l for l in out.split("\n") if l and l[0] != "#"][0].split("\t")
    assert (v[1], v[3], v[4]) == ("28018497", "CATGATA", "CATA")


def test_vcf_with_cluster_is_refused(bundled, bundled_cli_golden):
    c = bundled_cli_golden[0]
    with pytest.raises(SystemExit) as ex:
        run_report(bundled, c["find_mutation"], c["target"], "vcf", "cluster")
    assert "incompatible" in str(ex.value)


def test_min_cov_filter_drops_all_zero_reference_rows(bundled):
    # a Reference row over all-zero counts has Min_coverage 0 < -m 1
    text = "Database\tQuery\tType\n" + "x.jf\tMYC\tReference\t\tnan\tnan\t0\t0\tACGT\tnan\tACGT\tvs_ref\n"
    out, _ = run_report(bundled, text, "MYC_T58A_P59R_exon2")
    assert out.count("\n") == 1      # header only


def test_linear_kmin_known_answer(bundled):
    # km/tests/test_main.py:563-579
    from km_b200.tools import linear_kmin as lk
    args = Namespace(start=5, target_fn=[os.path.join(bundled, "data/catalog/GRCh38/FLT3-ITD_exons_13-15.fa")])
    with captured() as (out, _):
        lk.main_linear_kmin(args, None)
    assert out.getvalue().split("\n")[1].split("\t")[1] == "10"
