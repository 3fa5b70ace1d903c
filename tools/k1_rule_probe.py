import sys,time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import libprosic_b200 as P
from libprosic_b200 import synth
ctx=P.Context(0)
w=synth.make_pairhmm_workload(n_sites=int(sys.argv[1]) if len(sys.argv) > 1 else 1024, reads_per_site=140, seed=synth.SEED_C3)
gap=P.GapParams()
for flags in (P.HMM_SUM3_APPROX3, P.HMM_SUM3_APPROX, 0):
    for it in range(3):
        t=time.perf_counter()
        got=ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"], w["read_off"], gap, w["pair_hap"], w["pair_read"], flags=flags)
        dt=time.perf_counter()-t
    print(flags, len(w["pair_hap"]), 'pairs', dt*1e3,'ms', w["cells"]/dt/1e9,'GCUPS')
