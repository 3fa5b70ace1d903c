"""Generates tests/golden/pairhmm_golden.json from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference (Rust, bio 0.37) cannot be built in this
image, so these vectors freeze the RESTATEMENT, not the Rust binary: parity unpinned.
The first case is BASELINE.json config 1's hand-decoded candidate: tests/indels.bam read A.in1
(70M10I30M, inserted GCATCCTGCG after 0-based 546 of tests/chr17.prefix.fa) -- sequence
re-synthesised here, since fixtures may not be copied out of /root/reference."""
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402

rng = np.random.default_rng(20261019)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
GAPS = [[math.log(2.8e-6), math.log(5.1e-6), None, None],
        [math.log(0.02), math.log(0.03), math.log(0.3), math.log(0.3)]]
cases = []


def add(hap, read, qual, gap, flags):
    gl = np.array([x if x is not None else -math.inf for x in gap])
    hit = O.best_hit(hap, read)
    for med in (-1, hit["dist"] + 2):
        cases.append(dict(hap=bytes(hap).decode(), read=bytes(read).decode(),
                          qual=[int(x) for x in qual], gap_ln=gap, max_edit_dist=int(med),
                          flags=flags,
                          ln_prob=O.pairhmm_prob_related(hap, read, qual, gl, med, flags),
                          hit=[hit["dist"], hit["start"], hit["end"], hit["n_best"]]))


# insertion-carrying read vs ref window and vs alt window (config-1 shaped)
ref = ACGT[rng.integers(0, 4, 192)]
ins = np.frombuffer(b"GCATCCTGCG", dtype=np.uint8)
alt = np.concatenate([ref[:96], ins, ref[96:]])
read = alt[40:140].copy()
qual = np.clip(np.rint(rng.normal(32, 5, 100)), 2, 41).astype(np.uint8)
for flags in (0, 1):
    add(ref, read, qual, GAPS[0], flags)
    add(alt, read, qual, GAPS[0], flags)
for k in range(12):
    lx, ly = int(rng.integers(30, 200)), int(rng.integers(10, 100))
    hap = ACGT[rng.integers(0, 4, lx)]
    s = int(rng.integers(0, max(1, lx - ly)))
    r = hap[s:s + ly].copy()
    if len(r) < ly:
        r = np.concatenate([r, ACGT[rng.integers(0, 4, ly - len(r))]])
    for _ in range(int(rng.integers(0, 4))):
        r[rng.integers(0, ly)] = ACGT[rng.integers(0, 4)]
    q = rng.integers(2, 42, ly).astype(np.uint8)
    add(hap, r, q, GAPS[k % 2], k % 2)
# homopolymer repeat: sum-of-paths in repeats, where ln_sum3_exp_approx matters
hap = np.frombuffer(b"ACGT" + b"A" * 30 + b"CGTTGCA" + b"T" * 20 + b"GGC", dtype=np.uint8)
r = np.delete(hap[2:60].copy(), 12)
add(hap, r, np.full(len(r), 30, np.uint8), GAPS[0], 0)
add(hap, r, np.full(len(r), 30, np.uint8), GAPS[0], 1)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pairhmm_golden.json")
json.dump(dict(generator="tests/golden/make_golden.py", note="oracle restatement; parity unpinned",
               cases=cases), open(out, "w"), indent=0)
print(len(cases), "cases ->", out)
