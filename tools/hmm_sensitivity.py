#!/usr/bin/env python
"""Measured sensitivity of the UNPINNED recollections of bio 0.37's pair-HMM (the crate source is
not in /root/reference and there is no Rust toolchain): how much does each choice move the results?

Choices (oracle flags, oracle/vlr_oracle.h):
  sum3   exact | two-way ln_sum3_exp_approx (VLRO_HMM_SUM3_APPROX) | three-way (VLRO_HMM_SUM3_APPROX3)
  reset  all four column arrays after a column swap (default) | fm only (VLRO_HMM_RESET_FM_ONLY)
  close  close-gap constants paired as recollected | as PathHMMRealigner (VLRO_HMM_PATH_CLOSE)

Reported:
  (i)   max |delta ln P| of the raw allele probabilities over every realignment job (read window x
        ref / alt haplotype, banded dist + 2 on the shrunk window) of the 32 tumor/normal testcases;
  (ii)  max |delta ln P| over N unbanded C3 pairs (150 x 300);
  (iii) the expectation tables of the tumor/normal and the scenario-driven testcases per variant,
        with the four "floor" thresholds (test08 >= 590, test10 >= 702, test14 >= 196, test20 >= 318).
CPU only (oracle); needs /root/reference for (iii) on the full testcase set.

    python tools/hmm_sensitivity.py [n_c3_pairs] > tests/golden/hmm_sensitivity.txt
"""
import glob
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from _oracle_backend import OracleBackend  # noqa: E402
from libprosic_b200 import fixtures as F, synth  # noqa: E402
import libprosic_b200 as P  # noqa: E402

VARIANTS = [("exact sum", 0), ("two-way", O.HMM_SUM3_APPROX), ("three-way (default)", O.HMM_SUM3_APPROX3),
            ("three-way + fm-only reset", O.HMM_SUM3_APPROX3 | O.HMM_RESET_FM_ONLY),
            ("three-way + PathHMM close pairing", O.HMM_SUM3_APPROX3 | O.HMM_PATH_CLOSE)]
FLOORS = [("test08", "PROB_SOMATIC_TUMOR", 590.0), ("test10", "PROB_SOMATIC_TUMOR", 702.0),
          ("test14", "PROB_SOMATIC_TUMOR", 196.0), ("test20", "PROB_SOMATIC_TUMOR", 318.0)]


class Recorder:
    """OracleBackend that also keeps the raw values of every job."""

    def __init__(self, flags):
        self.be = OracleBackend(flags=flags)
        self.raw = []

    def allele_support(self, jobs):
        out = []
        for j in jobs:
            w = O.allele_support_region([np.frombuffer(j["ref"], np.uint8)], [np.frombuffer(a, np.uint8) for a, _ in j["alts"]],
                                        np.frombuffer(j["bases"], np.uint8), np.frombuffer(j["quals"], np.uint8),
                                        self.be.gap, shrink_extra=[0] + [e for _, e in j["alts"]], flags=self.be.flags)
            self.raw.append((w["raw_ref"], w["raw_alt"], w["prob_ref"] != w["prob_alt"]))
            out.append(dict(prob_ref=w["prob_ref"], prob_alt=w["prob_alt"], indel_ops=[int(x) for x in w["indel_ops"]]))
        return out

    def __getattr__(self, k):
        return getattr(self.be, k)


def tn_table(flags):
    rec = Recorder(flags)
    res = {}
    n_pass = n_tot = 0
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "tn", "*.json.gz"))):
        name = os.path.basename(p).split(".")[0]
        try:
            r = F.run_testcase(p, rec)
        except NotImplementedError:
            continue
        v = F.check_expectations(r)
        n_tot += 1
        n_pass += all(x[1] for x in v)
        res[name] = r
    return rec.raw, res, n_pass, n_tot


def scn_table(flags):
    be = OracleBackend(flags=flags)
    n_pass = n_tot = 0
    for sub in ("scn", "scn_prior", "gen", "trio"):
        for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", sub, "*.json.gz"))):
            try:
                if sub == "gen":
                    r = F.run_generic_case(p, be)
                elif sub == "trio":
                    r = F.run_trio_testcase(p, be)
                else:
                    r = F.run_scenario_testcase(p, be)
            except NotImplementedError:
                continue
            v = F.check_expectations(r)
            if v:
                n_tot += 1
                n_pass += all(x[1] for x in v)
    return n_pass, n_tot


def _c3_chunk(args):
    lo, hi, flags, n_sites = args
    w = synth.make_pairhmm_workload(n_sites=n_sites, reads_per_site=8, seed=synth.SEED_C3)
    gap = P.GapParams().as_ln_array()
    out = np.empty(hi - lo)
    for k in range(lo, hi):
        h, r = w["pair_hap"][k], w["pair_read"][k]
        out[k - lo] = O.pairhmm_prob_related(w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]],
                                             w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
                                             w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]], gap, -1, flags)
    return out


def c3_values(flags, n_pairs, pool, n_proc):
    n_sites = (n_pairs + 15) // 16
    step = (n_pairs + 4 * n_proc - 1) // (4 * n_proc)
    jobs = [(lo, min(lo + step, n_pairs), flags, n_sites) for lo in range(0, n_pairs, step)]
    return np.concatenate(pool.map(_c3_chunk, jobs))


def main():
    n_c3 = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    n_proc = os.cpu_count() or 1
    print("# sensitivity of the unpinned bio 0.37 recollections (tools/hmm_sensitivity.py, oracle only)")
    base_raw = base_c3 = None
    rows = []
    with Pool(n_proc) as pool:
        for name, flags in VARIANTS:
            raw, res, n_pass, n_tot = tn_table(flags)
            sp, st = scn_table(flags)
            c3 = c3_values(flags, n_c3, pool, n_proc) if not (flags & (O.HMM_RESET_FM_ONLY | O.HMM_PATH_CLOSE)) else None
            a = np.array([(x[0], x[1]) for x in raw])
            keep = np.array([x[2] for x in raw])
            if base_raw is None:
                base_raw, base_keep, base_c3 = a, keep, c3
            fin = np.isfinite(a) & np.isfinite(base_raw)
            d1 = float(np.max(np.abs(np.where(fin, a - base_raw, 0.0))))
            flips = int(np.sum(keep != base_keep))
            d2 = float(np.max(np.abs(c3 - base_c3))) if c3 is not None else float("nan")
            floors = " ".join("%s=%.1f" % (t, res[t][k]) for t, k, _ in FLOORS)
            rows.append((name, len(raw), d1, flips, d2, n_pass, n_tot, sp, st, floors))
    print("variant | C1 jobs | max |d ln P| vs exact, C1 (banded) | informative-flag flips vs exact | "
          "max |d ln P| vs exact, %d C3 pairs (unbanded) | TN testcases | other testcases | floors" % n_c3)
    for r in rows:
        print("%s | %d | %.3g | %d | %.3g | %d/%d | %d/%d | %s" % r)


if __name__ == "__main__":
    main()
