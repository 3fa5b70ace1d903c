#!/usr/bin/env python
"""Run ALL tumor/normal indel testcases of the reference through the restated preprocess/call glue
on the CPU oracle and print which `expected:` blocks hold (needs /root/reference; the GPU suite
runs the committed subset under tests/golden/tn/ through the CUDA kernels).

    python tests/golden/run_tn_testcases.py [traceback_rule [fast]] > tests/golden/tn_results.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from _oracle_backend import OracleBackend  # noqa: E402
from libprosic_b200 import fixtures as F  # noqa: E402

REF = "/root/reference/tests/resources/testcases"


def main():
    if len(sys.argv) > 1:
        oracle.lib().vlro_set_traceback_rule(int(sys.argv[1]))
    be = OracleBackend(fast=len(sys.argv) > 2 and sys.argv[2] == "fast")
    n_pass = n_tot = 0
    for c in ["test%02d" % k for k in range(2, 34)] + ["issue_154"]:
        d = os.path.join(REF, c)
        if not os.path.isdir(d):
            continue
        try:
            r = F.run_testcase(d, be)
        except NotImplementedError as e:
            print("%-10s SKIP  %s" % (c, e))
            continue
        v = F.check_expectations(r)
        ok = all(x[1] for x in v)
        n_tot += 1
        n_pass += ok
        probs = " ".join("%s=%.4g" % (k[5:].lower(), x) for k, x in r.items() if k.startswith("PROB_"))
        print("%-10s %s  obs n/t %d/%d  af n/t %.3f/%.3f  %s  %s" % (
            c, "PASS" if ok else "FAIL", r["n_obs"]["normal"], r["n_obs"]["tumor"], r["af"]["normal"], r["af"]["tumor"],
            probs, "" if ok else "violated: " + "; ".join(e for e, o in v if not o)))
    print("%d of %d testcases satisfy the reference's expectations" % (n_pass, n_tot))


if __name__ == "__main__":
    main()
