import sys, glob, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from _oracle_backend import OracleBackend
import oracle as O
from libprosic_b200 import fixtures as F
REF="/root/reference/tests/resources/testcases/"
tn=["test%02d"%k for k in range(2,34)]+["issue_154","test42","test47"]
scn=sorted(os.path.basename(d) for d in glob.glob(REF+"*") if os.path.exists(d+"/scenario.yaml"))
skip_up={"test21","test38","test_giab_01","test_overlapping_events","test_giab_04"}
for rule in range(9):
    O.lib().vlro_set_traceback_rule(rule)
    fails=[]
    n=0
    for c in tn+scn:
        if c in skip_up: continue
        try:
            r=(F.run_testcase if c in tn else F.run_scenario_testcase)(REF+c, OracleBackend())
        except (NotImplementedError, FileNotFoundError):
            continue
        v=F.check_expectations(r)
        n+=1
        if not all(x[1] for x in v): fails.append(c)
    print(rule, n-len(fails), 'of', n, 'fail:', fails, flush=True)
