#!/usr/bin/env python
"""Run every Generic-mode testcase of the reference that comes with a scenario.yaml through the
grammar front end + restated preprocess/call glue on the CPU oracle and print which `expected:`
blocks hold (needs /root/reference).

    python tests/golden/run_scenario_testcases.py > tests/golden/scenario_results.txt
"""
import glob
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _oracle_backend import OracleBackend  # noqa: E402
from libprosic_b200 import fixtures as F  # noqa: E402

REF = "/root/reference/tests/resources/testcases"


def main():
    be = OracleBackend()
    n_pass = n_tot = 0
    only = sys.argv[1:]
    for d in sorted(glob.glob(os.path.join(REF, "*"))):
        c = os.path.basename(d)
        if only and c not in only:
            continue
        if not os.path.exists(os.path.join(d, "scenario.yaml")):
            continue
        try:
            r = F.run_scenario_testcase(d, be)
        except NotImplementedError as e:
            print("%-36s SKIP  %s" % (c, e))
            continue
        except Exception as e:  # noqa: BLE001 -- survey run: report and go on
            print("%-36s ERROR %s: %s" % (c, type(e).__name__, str(e)[:100]))
            if only:
                traceback.print_exc()
            continue
        v = F.check_expectations(r)
        ok = all(x[1] for x in v)
        if not v:
            print("%-36s RUNS  (the reference asserts no expectation)  obs %s  af %s" % (
                c, "/".join(str(x) for x in r["n_obs"].values()), "/".join("%.3f" % x for x in r["af"].values())))
            continue
        n_tot += 1
        n_pass += ok
        probs = " ".join("%s=%.4g" % (k[5:].lower(), x) for k, x in r.items() if k.startswith("PROB_"))
        print("%-36s %s  obs %s  af %s  %s  %s" % (
            c, "PASS" if ok else "FAIL", "/".join(str(x) for x in r["n_obs"].values()),
            "/".join("%.3f" % x for x in r["af"].values()), probs,
            "" if ok else "violated: " + "; ".join(e for e, o in v if not o)))
    print("%d of %d testcases with an `expected:` block satisfy it.  Not asserted upstream either: test38 "
          "(commented out, tests/lib.rs:121), test_giab_01 (commented out, 158-160), test_giab_04 (commented out, "
          "163-166), test_overlapping_events (`testcase_should_panic!`, 184)" % (n_pass, n_tot))


if __name__ == "__main__":
    main()
