#!/usr/bin/env python
"""Compact fixtures of the reference's tumor/normal indel testcases (needs /root/reference, i.e.
this container): reference sequence, candidate record, alignment properties, `expected:` block
and the valid BAM records of both samples, as tests/golden/tn/<name>.json.gz.
The restated preprocess/call glue (libprosic_b200/fixtures.py) runs them end to end."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from libprosic_b200 import fixtures as F  # noqa: E402

REF = "/root/reference/tests/resources/testcases"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tn")
# every tumor/normal indel testcase of the reference (tests/lib.rs:80-110) plus issue_154
CASES = ["test%02d" % k for k in range(2, 34)] + ["issue_154"]


def main():
    os.makedirs(OUT, exist_ok=True)
    gen = os.path.join(os.path.dirname(OUT), "gen")
    os.makedirs(gen, exist_ok=True)
    total = 0
    trio = os.path.join(os.path.dirname(OUT), "trio")
    os.makedirs(trio, exist_ok=True)
    for c in F.TRIO_CASES:  # pedigree SNV testcases (Mendelian prior through the prior table)
        p = os.path.join(trio, c + ".json.gz")
        F.save_fixture(F.load_testcase(os.path.join(REF, c)), p)
        total += os.path.getsize(p)
        print(c, os.path.getsize(p))
    for c in F.GENERIC_CASES:  # one-sample Generic testcases (SNVs, prior table)
        p = os.path.join(gen, c + ".json.gz")
        F.save_fixture(F.load_testcase(os.path.join(REF, c)), p)
        total += os.path.getsize(p)
        print(c, os.path.getsize(p))
    scn = os.path.join(os.path.dirname(OUT), "scn")
    os.makedirs(scn, exist_ok=True)
    for c in F.SCENARIO_CASES:  # Generic testcases driven by their scenario.yaml (grammar front end)
        p = os.path.join(scn, c + ".json.gz")
        F.save_fixture(F.load_testcase(os.path.join(REF, c)), p)
        total += os.path.getsize(p)
        print(c, os.path.getsize(p))
    scn_prior = os.path.join(os.path.dirname(OUT), "scn_prior")
    os.makedirs(scn_prior, exist_ok=True)
    for c in F.SCENARIO_CASES_PRIOR:  # priors over continuous allele frequencies (prior program)
        p = os.path.join(scn_prior, c + ".json.gz")
        F.save_fixture(F.load_testcase(os.path.join(REF, c)), p)
        total += os.path.getsize(p)
        print(c, os.path.getsize(p))
    tn_bnd = os.path.join(os.path.dirname(OUT), "tn_bnd")
    os.makedirs(tn_bnd, exist_ok=True)
    for c in F.TN_BND_CASES:  # tumor/normal mode, candidate = breakend event
        p = os.path.join(tn_bnd, c + ".json.gz")
        F.save_fixture(F.load_testcase(os.path.join(REF, c)), p)
        total += os.path.getsize(p)
        print(c, os.path.getsize(p))
    for c in CASES:
        d = os.path.join(REF, c)
        if not os.path.isdir(d):
            continue
        try:
            tc = F.load_testcase(d)
            F.make_variant(tc, 96)
        except Exception as e:  # not a tumor/normal indel testcase
            print(c, "skipped:", type(e).__name__, str(e)[:60])
            continue
        p = os.path.join(OUT, c + ".json.gz")
        F.save_fixture(tc, p)
        total += os.path.getsize(p)
        print(c, os.path.getsize(p))
    print("total", total)


if __name__ == "__main__":
    main()
