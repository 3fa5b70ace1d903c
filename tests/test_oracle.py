"""CPU tests of the oracle (oracle/vlr_oracle.c): the restated arithmetic against
(a) a brute-force enumeration of alignment paths, (b) the reference's own unit-test identities
(src/variants/model/likelihood.rs:267-386), (c) committed golden vectors (tests/golden/),
(d) structural properties.  No GPU needed."""
import json
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
GAPS = [np.array([math.log(2.8e-6), math.log(5.1e-6), -math.inf, -math.inf]),
        np.log([0.02, 0.03, 0.3, 0.3]), np.log([0.02, 0.03, 0.2, 0.4])]


def test_logprob_helpers(oracle):
    L = oracle.lib()
    assert L.vlro_ln_add_exp(-math.inf, -math.inf) == -math.inf
    assert L.vlro_ln_add_exp(-1.0, -math.inf) == -1.0
    assert abs(L.vlro_ln_add_exp(math.log(0.25), math.log(0.5)) - math.log(0.75)) < 1e-15
    assert L.vlro_ln_one_minus_exp(0.0) == -math.inf
    assert L.vlro_ln_one_minus_exp(-math.inf) == 0.0
    assert abs(L.vlro_ln_one_minus_exp(math.log(0.3)) - math.log(0.7)) < 1e-15
    v = np.array([-3.0, -1.0, -math.inf, -2.0])
    want = math.log(sum(math.exp(x) for x in v if x > -math.inf))
    assert abs(L.vlro_ln_sum_exp(v.ctypes.data_as(oracle._DP), 4) - want) < 1e-15
    # bases.rs LUT: Q30 -> eps = 1e-3
    assert abs(L.vlro_ln_prob_miscall(30) - math.log(1e-3)) < 1e-12
    assert abs(L.vlro_ln_prob_call(30) - math.log(1 - 1e-3)) < 1e-15
    assert L.vlro_ln_prob_call(0) == -math.inf  # Q0: ln(1 - 1) (bases.rs:29-31,41-47)


@pytest.mark.parametrize("flags", [0, 2])
def test_forward_equals_path_enumeration(oracle, flags):
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(150):
        lx, ly = int(rng.integers(1, 7)), int(rng.integers(1, 5))
        hap, read = ACGT[rng.integers(0, 4, lx)], ACGT[rng.integers(0, 4, ly)]
        q = rng.integers(2, 41, ly).astype(np.uint8)
        for g in GAPS:
            a = oracle.pairhmm_prob_related(hap, read, q, g, -1, flags)
            b = oracle.pairhmm_bruteforce(hap, read, q, g, flags)
            worst = max(worst, abs(a - b))
    assert worst < 1e-12


def test_band_only_removes_mass_and_wide_band_is_exact(oracle):
    rng = np.random.default_rng(2)
    for _ in range(40):
        lx, ly = int(rng.integers(20, 80)), int(rng.integers(5, 40))
        hap = ACGT[rng.integers(0, 4, lx)]
        start = int(rng.integers(0, max(1, lx - ly)))
        read = hap[start:start + ly].copy()
        if len(read) < ly:
            read = np.concatenate([read, ACGT[rng.integers(0, 4, ly - len(read))]])
        read[rng.integers(0, ly)] = ord("A")
        q = rng.integers(10, 41, ly).astype(np.uint8)
        full = oracle.pairhmm_prob_related(hap, read, q, GAPS[0], -1, 0)
        wide = oracle.pairhmm_prob_related(hap, read, q, GAPS[0], lx + ly, 0)
        assert abs(full - wide) < 1e-12
        hit = oracle.best_hit(hap, read)
        band = oracle.pairhmm_prob_related(hap, read, q, GAPS[0], hit["dist"] + 2, 0)
        assert band <= full + 1e-12


def test_myers_best_hit(oracle):
    hap = np.frombuffer(b"TTTTACGTACGGACGTTTTT", dtype=np.uint8)
    read = np.frombuffer(b"ACGTACGTACGT", dtype=np.uint8)  # one substitution G->T at read pos 7
    hit = oracle.best_hit(hap, read, want_alignments=True)
    assert hit["dist"] == 1 and hit["start"] == 4 and hit["n_best"] == 1
    assert hit["end"] == min(4 + 12 + 1, 20)
    assert list(hit["alignments"][0]) == [0] * 7 + [1] + [0] * 4
    # deletion in the read relative to the haplotype -> one Del op (2)
    read2 = np.frombuffer(b"ACGTAGGACGT", dtype=np.uint8)  # "C" of ACGG.. missing
    hit2 = oracle.best_hit(hap, read2, want_alignments=True)
    assert hit2["dist"] == 1 and list(hit2["indel_ops"]) == [2]
    # perfect match: dist 0, no indel ops; lower-case haplotype is upper-cased (edit_distance.rs:61-62)
    hit3 = oracle.best_hit(np.frombuffer(b"ttttacgtacggacgttttt", dtype=np.uint8),
                           np.frombuffer(b"ACGTACGG", dtype=np.uint8))
    assert hit3["dist"] == 0 and hit3["start"] == 4 and len(hit3["indel_ops"]) == 0


def test_allele_support_rules(oracle):
    """realignment/mod.rs:283-301, 368-372: normalisation, certainty short-circuit, 0.5 fallback."""
    rng = np.random.default_rng(3)
    ref = ACGT[rng.integers(0, 4, 200)]
    alt = np.concatenate([ref[:100], ref[110:]])          # 10 bp deletion
    read_alt = alt[60:140].copy()                          # perfect alt read
    q = np.full(80, 30, np.uint8)
    s = oracle.allele_support_region([ref], [alt], read_alt, q, GAPS[0])
    assert s["alt_dist"] == 0 and s["ref_dist"] > 0
    certainty = 80 * math.log(1 - 1e-3)
    assert abs(s["raw_alt"] - certainty) < 1e-9            # dist == 0 -> certainty_est (Q7)
    assert abs(math.exp(s["prob_ref"]) + math.exp(s["prob_alt"]) - 1.0) < 1e-12
    assert s["prob_alt"] > s["prob_ref"] and s["informative"]
    # read far from the variant matches both alleles perfectly -> equal, uninformative
    s2 = oracle.allele_support_region([ref], [alt], ref[5:60].copy(), np.full(55, 30, np.uint8), GAPS[0])
    assert not s2["informative"] and abs(s2["prob_ref"] - math.log(0.5)) < 1e-12
    # Q0 everywhere: both alleles ln 0 -> 0.5 / 0.5 fallback (mod.rs:292-301)
    s3 = oracle.allele_support_region([ref], [alt], ref[5:60].copy(), np.zeros(55, np.uint8), GAPS[0])
    assert s3["prob_ref"] == math.log(0.5) and s3["prob_alt"] == math.log(0.5)
    # fast mode (PathHMM, mod.rs:499-576): for a substitution-only best path the single-path
    # probability cannot exceed the forward sum (alt allele here); the ref allele needs a 10 bp
    # deletion, which PathHMM prices with its re-open edge (mod.rs:471-476) while the forward
    # recursion has no X->X re-open -- so no ordering holds there (ambiguity 2)
    read_mm = read_alt.copy(); read_mm[10] = ord("A") if read_mm[10] != ord("A") else ord("C")
    ex = oracle.allele_support_region([ref], [alt], read_mm, q, GAPS[0], flags=0)
    fa = oracle.allele_support_region([ref], [alt], read_mm, q, GAPS[0], fast_mode=True)
    assert fa["raw_alt"] <= ex["raw_alt"] + 1e-9
    # ref path: 80 aligned bases (79 matches, 1 substitution), 10 Del ops, 78 M->M transitions
    want_ref = (10 * math.log(5.1e-6) + 79 * math.log(1 - 1e-3) + math.log(1e-3) + math.log(0.3333)
                + 78 * math.log(1 - (2.8e-6 + 5.1e-6)))
    # (the traceback splits the deletion into runs, which moves a few no-gap factors: 1e-4 slack)
    assert abs(fa["raw_ref"] - want_ref) < 1e-4


def test_minilogprob_matches_numpy_half(oracle):
    """utils/mod.rs:381-407; half::f16::from_f64 == numpy float16 cast (round to nearest even)."""
    rng = np.random.default_rng(4)
    L = oracle.lib()
    vals = np.concatenate([-np.exp(rng.uniform(-12, 12, 4000)), [-10.0, -10.0000001, -65504.0, -65520.0, -1e5, 0.0]])
    with np.errstate(over="ignore"):
        for v in vals:
            h = np.float64(v).astype(np.float16)
            assert L.vlro_f64_to_f16_bits(float(v)) == int(h.view(np.uint16))
            proj = float(h)
            same = (not math.isinf(proj)) and math.floor(proj) == math.floor(v)
            want = proj if (v < -10.0 and same) else float(np.float32(v))
            got = L.vlro_minilogprob_roundtrip(float(v))
            assert got == want or (math.isinf(got) and math.isinf(want))


# ---- the reference's own likelihood identities (likelihood.rs:247-387, model/mod.rs:269-290) ----
def _observation_pileup(oracle, rows):
    """rows: list of (prob_mapping, prob_alt, prob_ref) -> pileup built like tests::observation()."""
    L = oracle.lib()
    cols = {n: [] for n in oracle.OBS_COLUMNS}
    for pm, pa, pr in rows:
        cols["prob_mapping"].append(pm)
        cols["prob_mismapping"].append(L.vlro_ln_one_minus_exp(pm))
        cols["prob_alt"].append(pa)
        cols["prob_ref"].append(pr)
        cols["prob_missed_allele"].append(L.vlro_ln_add_exp(pr, pa) - math.log(2.0))
        cols["prob_sample_alt"].append(0.0)
        cols["prob_double_overlap"].append(0.0)
        cols["prob_single_overlap"].append(L.vlro_ln_one_minus_exp(0.0))
        cols["prob_hit_base"].append(math.log(0.01))
    cat = oracle.pack_cat([2] * len(rows), [8] * len(rows), [1] * len(rows), [0] * len(rows),
                          [2] * len(rows), [1] * len(rows))
    return oracle.PileupView({k: np.array(v) for k, v in cols.items()}, cat)


def _rel_eq(a, b, eps=1e-12):
    return abs(a - b) <= eps * max(abs(a), abs(b), 1e-300)


def test_reference_likelihood_identities(oracle):
    import ctypes as C
    L = oracle.lib()
    none = oracle.make_biases()
    ninf = -math.inf
    # test_likelihood_observation_absent_single / _absent (likelihood.rs:267-293)
    pv1 = _observation_pileup(oracle, [(0.0, ninf, 0.0)])
    any1 = L.vlro_biases_prob_any(C.byref(none), C.byref(pv1.c), 0)
    assert _rel_eq(oracle.likelihood_single(pv1, 0.0, none), any1)
    assert _rel_eq(oracle.likelihood_contaminated(pv1, 1.0, 0.0, none, 0.0, none), any1)
    # test_likelihood_pileup_absent / _absent_single (295-350): 10 alt-zero reads at AF=0
    pv10 = _observation_pileup(oracle, [(0.0, ninf, 0.0)] * 10)
    want = sum(L.vlro_biases_prob_any(C.byref(none), C.byref(pv10.c), i) for i in range(10))
    assert _rel_eq(oracle.likelihood_contaminated(pv10, 1.0, 0.0, none, 0.0, none), want)
    assert _rel_eq(oracle.likelihood_single(pv10, 0.0, none), want)
    # test_likelihood_pileup (352-386): 5 alt + 5 ref reads -> AF 0.5 maximises the likelihood
    pv = _observation_pileup(oracle, [(0.0, 0.0, ninf)] * 5 + [(0.0, ninf, 0.0)] * 5)
    lh = oracle.likelihood_contaminated(pv, 1.0, 0.5, none, 0.0, none)
    for af in np.linspace(0.0, 1.0, 10):
        if af != 0.5:
            assert lh > oracle.likelihood_contaminated(pv, 1.0, float(af), none, 0.0, none)


def test_grid_and_observable_bounds(oracle):
    L = oracle.lib()
    # generic.rs:109-122
    assert L.vlro_grid_points(100, 100) == 101 and L.vlro_grid_points(40, 5) == 5
    assert L.vlro_grid_points(30, 100) == 31 and L.vlro_grid_points(0, 100) == 5
    assert L.vlro_grid_points(3, 100) == 5 and L.vlro_grid_points(5, 100) == 7
    # formula.rs:1029-1071, worked example of SURVEY 8a: ]0,1] with 100 obs -> [1/100, 1]
    assert L.vlro_observable_min(0.0, 1.0, 1, 0, 100) == 0.01
    assert L.vlro_observable_max(0.0, 1.0, 1, 0, 100) == 1.0
    assert L.vlro_observable_min(0.0, 0.5, 1, 1, 40) == 1 / 40
    assert L.vlro_observable_max(0.0, 0.5, 1, 1, 40) == 19 / 40
    assert L.vlro_observable_min(0.0, 1.0, 1, 0, 5) == 0.0  # quirk Q2: n_obs < 10 keeps raw start


def test_golden_vectors(oracle):
    """Committed outputs of the oracle itself (tests/golden/make_golden.py): guards against
    accidental changes of the restatement; NOT a pin against the Rust binary (parity unpinned)."""
    path = os.path.join(GOLDEN, "pairhmm_golden.json")
    g = json.load(open(path))
    for c in g["cases"]:
        hap = np.frombuffer(c["hap"].encode(), dtype=np.uint8)
        read = np.frombuffer(c["read"].encode(), dtype=np.uint8)
        q = np.array(c["qual"], dtype=np.uint8)
        gap = np.array([x if x is not None else -math.inf for x in c["gap_ln"]])
        got = oracle.pairhmm_prob_related(hap, read, q, gap, c["max_edit_dist"], c["flags"])
        assert abs(got - c["ln_prob"]) <= 1e-12 * max(1.0, abs(c["ln_prob"])), c
        hit = oracle.best_hit(hap, read)
        assert [hit["dist"], hit["start"], hit["end"], hit["n_best"]] == c["hit"]
