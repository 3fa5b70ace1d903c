"""Config C1: part 1 on the reference's bundled tumor/normal indel testcases.

tests/golden/c1_realign.json holds the (read window, ref window, alt window) triples that the
restated host glue (tests/golden/make_c1_golden.py: candidate_region, Deletion / Insertion
haplotypes) produces from the REAL reads of ten reference testcases.  The GPU path must agree with
the oracle on them (allele log-probabilities within 1e-6, edit distances and indel operations
bit-exact, exact and fast realignment modes); on the CPU the oracle's per-read verdicts are
checked against what the testcases are known to contain."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def c1():
    return json.load(open(os.path.join(HERE, "golden", "c1_realign.json")))


def _arr(s):
    return np.frombuffer(s.encode(), dtype=np.uint8)


def _oracle_counts(oracle, case, gap_ln):
    res = {"tumor": [0, 0, 0], "normal": [0, 0, 0]}
    alt = _arr(case["alt"])
    for r in case["regions"]:
        w = oracle.allele_support_region([_arr(case["ref_windows"][r["ref"]])], [alt], _arr(r["bases"]),
                                         np.array(r["quals"], np.uint8), gap_ln, shrink_extra=[0, case["alt_extra"]])
        d = w["prob_alt"] - w["prob_ref"]
        res[r["sample"]][0 if d > 1 else (1 if d < -1 else 2)] += 1
    return res


def test_oracle_verdicts_match_the_testcases(c1):
    """testcase.yaml truths: test02 / test16 are somatic tumor-only deletions (no alt read in the
    normal sample), test03 / test24 are germline heterozygous 1-bp deletions (alt reads in both)."""
    import oracle
    import libprosic_b200 as P
    gap = P.GapParams().as_ln_array()
    by_name = {c["case"]: c for c in c1}
    for name in ("test02", "test16"):
        r = _oracle_counts(oracle, by_name[name], gap)
        assert r["normal"][0] == 0 and r["tumor"][0] >= 1 and r["normal"][1] >= 20, (name, r)
    for name in ("test03", "test24"):
        r = _oracle_counts(oracle, by_name[name], gap)
        assert r["normal"][0] >= 6 and r["tumor"][0] >= 6 and r["normal"][1] >= 6 and r["tumor"][1] >= 6, (name, r)


def _flatten(c1):
    haps, extra, reads, quals = [], [], [], []
    region_read, region_haps, region_hap_off, region_n_ref, meta = [], [], [0], [], []
    for case in c1:
        h0 = len(haps)
        for w in case["ref_windows"]:
            haps.append(_arr(w))
            extra.append(0)
        alt_id = len(haps)
        haps.append(_arr(case["alt"]))
        extra.append(case["alt_extra"])
        for r in case["regions"]:
            region_read.append(len(reads))
            reads.append(_arr(r["bases"]))
            quals.append(np.array(r["quals"], np.uint8))
            region_haps += [h0 + r["ref"], alt_id]
            region_hap_off.append(len(region_haps))
            region_n_ref.append(1)
            meta.append((case, r))
    hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    return dict(haps=haps, extra=np.array(extra, np.int32), reads=reads, quals=quals, meta=meta,
                hap_bytes=np.concatenate(haps), hap_off=hap_off, read_bases=np.concatenate(reads),
                read_quals=np.concatenate(quals), read_off=read_off,
                region_read=np.array(region_read, np.int32), region_haps=np.array(region_haps, np.int32),
                region_hap_off=np.array(region_hap_off, np.int32), region_n_ref=np.array(region_n_ref, np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [False, True])
def test_c1_allele_support_gpu_vs_oracle(ctx, oracle, c1, fast):
    import libprosic_b200 as P
    R = _flatten(c1)
    gap = P.GapParams()
    flags = P.HMM_DEFAULT | (P.REALIGN_FAST if fast else 0)
    got = ctx.allele_support_batch(R["region_read"], R["region_hap_off"], R["region_haps"], R["region_n_ref"],
                                   R["hap_bytes"], R["hap_off"], R["read_bases"], R["read_quals"], R["read_off"],
                                   gap, shrink_extra=R["extra"], flags=flags)
    n_inf = 0
    for k, (case, r) in enumerate(R["meta"]):
        ids = R["region_haps"][2 * k:2 * k + 2]
        want = oracle.allele_support_region([R["haps"][ids[0]]], [R["haps"][ids[1]]], R["reads"][k], R["quals"][k],
                                            gap.as_ln_array(), shrink_extra=[0, case["alt_extra"]], fast_mode=fast)
        for key in ("prob_ref", "prob_alt", "raw_ref", "raw_alt"):
            g, w = got[key][k], want[key]
            assert (np.isneginf(g) and np.isneginf(w)) or abs(g - w) <= 1e-6, (case["case"], r["name"], key, g, w)
        assert got["ref_dist"][k] == want["ref_dist"] and got["alt_dist"][k] == want["alt_dist"], (case["case"], k)
        assert got["n_indel_ops"][k] == len(want["indel_ops"]), (case["case"], k)
        assert np.array_equal(got["indel_ops"][k][:len(want["indel_ops"])], want["indel_ops"]), (case["case"], k)
        n_inf += want["informative"]
    assert n_inf > 300
