"""Size-independent properties of the GPU path at the sizes of the benchmark configurations
(BASELINE.json configs C2 / C3 per bench step): where the oracle cannot follow in seconds, the path
is checked through invariants -- batch-order and batch-composition independence (bit for bit),
unbanded == banded with an unreachable band, normalisation of the per-read allele support and of
the event posteriors, invariance against the order of the observations of a pileup -- plus an
oracle spot check on a random sample."""
import numpy as np
import pytest

import libprosic_b200 as P
from libprosic_b200 import scenario, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    return synth.make_pairhmm_workload(1792, 140, seed=synth.SEED_C3 + 5)   # 501 760 pairs of 150 x 300


def _hmm(ctx, w, ph, pr, med=None, flags=P.HMM_DEFAULT):
    return ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"], w["read_off"],
                             P.GapParams(), ph, pr, max_edit_dist=med, flags=flags)


def test_pairhmm_full_size_invariants(ctx, oracle, c3):
    w = c3
    ph, pr = w["pair_hap"], w["pair_read"]
    n = len(ph)
    base = _hmm(ctx, w, ph, pr)
    assert base.shape == (n,) and np.all(np.isfinite(base)) and np.all(base <= 0.0)
    # order of the pairs (bucketing by read length, dynamic scheduling) does not matter -- bit for bit
    rng = np.random.default_rng(1)
    perm = rng.permutation(n)
    assert np.array_equal(_hmm(ctx, w, ph[perm], pr[perm]), base[perm])
    # neither does the composition of the batch
    sub = rng.choice(n, 4096, replace=False)
    assert np.array_equal(_hmm(ctx, w, ph[sub], pr[sub]), base[sub])
    # a band no alignment can leave is the unbanded recursion (PairHMM::prob_related with and without max_edit_dist)
    some = rng.choice(n, 65536, replace=False)
    banded = _hmm(ctx, w, ph[some], pr[some], med=np.full(len(some), 100000, np.int32))
    assert np.max(np.abs(banded - base[some])) <= 1e-9
    # the read's own haplotype explains it at least as well as the other one for all but a few reads
    pairs = base.reshape(-1, 2)
    assert np.mean(np.abs(pairs[:, 0] - pairs[:, 1]) > 1.0) > 0.4       # informative reads exist
    # oracle spot check
    gap = P.GapParams().as_ln_array()
    for k in rng.choice(n, 48, replace=False):
        h, r = int(ph[k]), int(pr[k])
        want = oracle.pairhmm_prob_related(w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]],
                                           w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
                                           w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]], gap, -1, P.HMM_DEFAULT)
        assert abs(base[k] - want) <= 1e-6


def test_allele_support_full_size_invariants(ctx, c3):
    w = c3
    n_reads = 131072
    site_of = w["site_of"][:n_reads]
    rr = np.arange(n_reads, dtype=np.int32)
    rh = (2 * site_of[:, None] + np.arange(2)[None, :]).reshape(-1).astype(np.int32)
    rho = np.arange(0, 2 * n_reads + 1, 2, dtype=np.int32)
    rnr = np.ones(n_reads, dtype=np.int32)
    rb, rq, ro = (w["read_bases"][:int(w["read_off"][n_reads])], w["read_quals"][:int(w["read_off"][n_reads])],
                  w["read_off"][:n_reads + 1])
    out = ctx.allele_support_batch(rr, rho, rh, rnr, w["hap_bytes"], w["hap_off"], rb, rq, ro, P.GapParams())
    pr_, pa = out["prob_ref"], out["prob_alt"]
    assert np.all(pr_ <= 0.0) and np.all(pa <= 0.0) and not np.isnan(pr_).any() and not np.isnan(pa).any()
    # Realigner::allele_support normalises: e^ref + e^alt = 1 (both -inf cannot occur: ln .5 each then)
    assert np.max(np.abs(np.logaddexp(pr_, pa))) <= 1e-12
    # distance 0 <=> the certainty short cut: raw probability is the sum of ln(1 - eps), never above 0
    assert np.all(out["raw_ref"][out["ref_dist"] == 0] <= 0.0)
    # batch composition does not matter -- bit for bit (reads of the first 512 regions alone)
    m = 512
    sub = ctx.allele_support_batch(rr[:m], rho[:m + 1], rh[:2 * m], rnr[:m], w["hap_bytes"], w["hap_off"],
                                   rb[:int(ro[m])], rq[:int(ro[m])], ro[:m + 1], P.GapParams())
    for k in ("prob_ref", "prob_alt", "ref_dist", "alt_dist", "n_indel_ops"):
        assert np.array_equal(sub[k], out[k][:m]), k


@pytest.mark.parametrize("name", ["germline", "continuous", "trio", "tumor_normal"])
def test_posterior_full_size_invariants(ctx, name):
    if name == "tumor_normal":
        scn, depths, n_sites, kind = scenario.tumor_normal(purity=0.75, indel=True), [40, 100], 16384, "indel"
        kw = dict(af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]], purity=[1.0, 0.75])
    elif name == "trio":
        scn, depths, n_sites, kind = scenario.trio(with_biases=True), [30, 30, 30], 65536, "snv"
        kw = dict(af_sets=[[(0, .9), (.5, .07), (1, .03)]] * 3)
    else:
        scn, depths, n_sites, kind = scenario.single_sample(continuous=name == "continuous"), [30], 262144, "snv"
        kw = {}
    ns, ne = scn["n_samples"], len(scn["events"])
    cols, cat, off, _ = synth.make_pileups(n_sites, depths, seed=99, kind=kind, artifact_frac=0.05, **kw)
    got = ctx.posterior_batch(scn, cols, cat, off)
    ev, marg = got["event_ln"].reshape(n_sites, ne), got["marginal"]
    assert not np.isnan(ev).any() and np.all(np.isfinite(marg))
    # Model::compute: the marginal is the ln-sum of the event posteriors (gated-out events are -inf)
    m = ev.max(axis=1)
    with np.errstate(invalid="ignore"):
        lse = m + np.log(np.exp(np.where(np.isfinite(ev), ev - m[:, None], -np.inf)).sum(axis=1))
    assert np.max(np.abs(lse - marg)) <= 1e-9
    # sites are independent: a permuted batch gives the permuted results -- bit for bit
    rng = np.random.default_rng(3)
    perm = rng.permutation(n_sites)[:20000]
    rows = np.concatenate([np.arange(off[s * ns], off[(s + 1) * ns]) for s in perm])
    lens = np.concatenate([np.diff(off[s * ns:(s + 1) * ns + 1]) for s in perm])
    off2 = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cols2 = {k: v[rows] for k, v in cols.items()}
    got2 = ctx.posterior_batch(scn, cols2, cat[rows], off2)
    assert np.array_equal(got2["event_ln"].reshape(-1, ne), ev[perm])
    assert np.array_equal(got2["map_af"].reshape(-1, ns), got["map_af"].reshape(n_sites, ns)[perm], equal_nan=True)
    # the order of the observations inside a pileup changes the summation order only
    rows3 = rows.copy()
    p = 0
    for ln_ in lens:
        rows3[p:p + ln_] = rows3[p:p + ln_][::-1]
        p += ln_
    got3 = ctx.posterior_batch(scn, {k: v[rows3] for k, v in cols.items()}, cat[rows3], off2)
    e3, e0 = got3["event_ln"].reshape(-1, ne), ev[perm]
    assert np.array_equal(np.isfinite(e3), np.isfinite(e0))
    fin = np.isfinite(e0)
    assert np.max(np.abs(e3[fin] - e0[fin])) <= 1e-9
    same = np.isclose(got3["map_af"].reshape(-1, ns), got["map_af"].reshape(n_sites, ns)[perm], equal_nan=True).all(axis=1)
    assert same.mean() > 0.999   # exact ties between allele frequencies may resolve differently
