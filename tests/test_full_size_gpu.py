"""The BASELINE.json configurations at FULL size, checked through size-independent properties
(the oracle needs ~30 s per 1000 C3 pairs, so it only sees samples): a result must not depend on
where in the batch its unit sits, duplicates must agree bit for bit, and a sample must match the
oracle.  C3 part 1: 7168 sites x 140 reads x 2 haplotypes = 2 007 040 pairs of 150 x 300;
allele support on 262 144 of those reads; C2 part 2: 262 144 sites x 30 observations."""
import numpy as np
import pytest

import libprosic_b200 as P
from libprosic_b200 import scenario, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    return synth.make_pairhmm_workload(n_sites=7168, reads_per_site=140, seed=synth.SEED_C3)


def _sub_batch(w, pairs):
    """The listed pairs as a self-contained small batch (own CSR arrays)."""
    hs, rs = w["pair_hap"][pairs], w["pair_read"][pairs]
    haps = [w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]] for h in hs]
    rb = [w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]] for r in rs]
    rq = [w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]] for r in rs]
    ho = np.concatenate([[0], np.cumsum([len(x) for x in haps])]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum([len(x) for x in rb])]).astype(np.int64)
    return np.concatenate(haps), ho, np.concatenate(rb), np.concatenate(rq), ro


def test_c3_pairhmm_full_size(ctx, oracle, c3):
    w = c3
    n = len(w["pair_hap"])
    assert n == 2007040 and w["cells"] == n * 150 * 300
    gap = P.GapParams()
    got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"], w["read_off"], gap,
                            w["pair_hap"], w["pair_read"])
    assert np.isfinite(got).all() and (got <= 0.0).all()
    # position independence: 4096 pairs drawn from all over the batch, recomputed as their own batch
    # (other CSR offsets, other bucket positions, other warps) -- bit for bit
    rng = np.random.default_rng(9)
    pick = np.sort(rng.choice(n, 4096, replace=False))
    hb, ho, rb, rq, ro = _sub_batch(w, pick)
    again = ctx.pairhmm_batch(hb, ho, rb, rq, ro, gap)
    assert np.array_equal(again.view(np.int64), got[pick].view(np.int64))
    # every read is paired with both haplotypes of its site and was sampled from one of them: the
    # two probabilities differ by the indel, the larger one belongs to a plausible alignment
    two = got.reshape(-1, 2)
    assert (np.max(two, axis=1) > -60.0).mean() > 0.99
    # a sample against the oracle
    for k in pick[:48]:
        h, r = int(w["pair_hap"][k]), int(w["pair_read"][k])
        want = oracle.pairhmm_prob_related(w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]],
                                           w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
                                           w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]],
                                           gap.as_ln_array(), -1, P.HMM_DEFAULT)
        assert abs(got[k] - want) <= 1e-6, (k, got[k], want)


def test_c3_allele_support_full_size(ctx, oracle, c3):
    w = c3
    n_reads = 262144
    gap = P.GapParams()
    rr = np.arange(n_reads, dtype=np.int32)
    rh = (2 * w["site_of"][:n_reads, None] + np.arange(2)[None, :]).reshape(-1).astype(np.int32)
    rho = np.arange(0, 2 * n_reads + 1, 2, dtype=np.int32)
    ones = np.ones(n_reads, np.int32)
    got = ctx.allele_support_batch(rr, rho, rh, ones, w["hap_bytes"], w["hap_off"], w["read_bases"],
                                   w["read_quals"], w["read_off"], gap)
    pr, pa = got["prob_ref"], got["prob_alt"]
    assert np.isfinite(pr).all() and np.isfinite(pa).all()
    both = np.isfinite(got["raw_ref"]) & np.isfinite(got["raw_alt"]) if "raw_ref" in got else np.ones(n_reads, bool)
    # normalised: exp(prob_ref) + exp(prob_alt) = 1 wherever both alleles had a hit
    s = np.exp(pr) + np.exp(pa)
    assert np.abs(s[both] - 1.0).max() <= 1e-9
    # position independence on 2048 reads taken from all over the batch, as their own batch
    rng = np.random.default_rng(10)
    pick = np.sort(rng.choice(n_reads, 2048, replace=False)).astype(np.int32)
    sub_rh = rh.reshape(-1, 2)[pick].reshape(-1)
    sub = ctx.allele_support_batch(pick, np.arange(0, 2 * len(pick) + 1, 2, dtype=np.int32), sub_rh,
                                   np.ones(len(pick), np.int32), w["hap_bytes"], w["hap_off"], w["read_bases"],
                                   w["read_quals"], w["read_off"], gap)
    for key in ("prob_ref", "prob_alt"):
        assert np.array_equal(sub[key].view(np.int64), got[key][pick].view(np.int64)), key
    for key in ("ref_dist", "alt_dist", "n_indel_ops"):
        assert np.array_equal(sub[key], got[key][pick]), key
    # a sample against the oracle, kept-observation decision (prob_ref != prob_alt) included
    for r in pick[:40]:
        hp = [w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]] for h in rh[2 * r:2 * r + 2]]
        ws = oracle.allele_support_region([hp[0]], [hp[1]], w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
                                          w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]], gap.as_ln_array())
        assert abs(ws["prob_ref"] - pr[r]) <= 1e-6 and abs(ws["prob_alt"] - pa[r]) <= 1e-6, r
        assert (ws["prob_ref"] != ws["prob_alt"]) == (pr[r] != pa[r]), r


def test_c2_posterior_full_size(ctx, oracle):
    scn = scenario.single_sample(continuous=False)
    n_sites = 262144
    cols, cat, off, _ = synth.make_pileups(n_sites, [30], seed=synth.SEED_C2, kind="snv", artifact_frac=0.05)
    got = ctx.posterior_batch(scn, cols, cat, off)
    ev, marg = got["event_ln"], got["marginal"]
    assert np.isfinite(marg).all()
    # the marginal is the ln-sum of the event posteriors (Model::compute)
    m = np.max(ev, axis=1)
    lse = m + np.log(np.sum(np.exp(ev - m[:, None]), axis=1))
    assert np.abs(lse - marg).max() <= 1e-9
    # position independence: 4096 sites from all over the batch as their own batch, bit for bit
    rng = np.random.default_rng(11)
    pick = np.sort(rng.choice(n_sites, 4096, replace=False))
    rows = np.concatenate([np.arange(off[s], off[s + 1]) for s in pick])
    sub_off = np.concatenate([[0], np.cumsum(off[pick + 1] - off[pick])]).astype(np.int64)
    sub = ctx.posterior_batch(scn, {k: v[rows] for k, v in cols.items()}, cat[rows], sub_off)
    assert np.array_equal(sub["event_ln"].view(np.int64), ev[pick].view(np.int64))
    assert np.array_equal(sub["map_af"], got["map_af"][pick], equal_nan=True)
    # a sample against the oracle
    spec = oracle.ModelSpec(scn["n_samples"], scn["resolution"], scn["contaminated"], scn["by"], scn["purity"],
                            scn["nodes"], scn["events"], [oracle.make_biases(**b) for b in scn["biases"]])
    r = oracle.model_compute_batch(spec, {k: v[rows] for k, v in cols.items()}, cat[rows], sub_off, 256, 4)
    fin = np.isfinite(r["event_ln"]) & np.isfinite(sub["event_ln"][:256])
    assert np.abs(r["event_ln"][fin] - sub["event_ln"][:256][fin]).max() <= 1e-6
    assert np.array_equal(r["map_af"], sub["map_af"][:256], equal_nan=True)
