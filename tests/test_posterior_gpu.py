"""GPU parity of K3/K4 (pileup likelihood on the AF grid + event integration) against the CPU
oracle, through the C ABI (vlr_posterior_batch / vlr_likelihood_batch).
Tolerances: |delta| <= 1e-6 in natural-log space for event posteriors and marginals; MAP allele
frequencies and bias calls identical."""
import math

import numpy as np
import pytest

import libprosic_b200 as P
from libprosic_b200 import scenario, synth

pytestmark = pytest.mark.gpu
TOL = 1e-6


def _oracle_spec(oracle, scn):
    return oracle.ModelSpec(scn["n_samples"], scn["resolution"], scn["contaminated"], scn["by"],
                            scn["purity"], scn["nodes"], scn["events"],
                            [oracle.make_biases(**b) for b in scn["biases"]])


def _oracle_run(oracle, scn, cols, cat, off, sites):
    spec = _oracle_spec(oracle, scn)
    ns = scn["n_samples"]
    out = []
    for s in sites:
        pv = [oracle.PileupView(cols, cat, int(off[s * ns + k]), int(off[s * ns + k + 1]))
              for k in range(ns)]
        r = oracle.model_compute(spec, pv, prior=scn.get("prior_fn"))
        r["map_bias_cfg"] = np.array([spec.flat_bias_idx[b] if b >= 0 else -1 for b in r["map_bias"]])
        out.append(r)
    return out


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    same_inf = (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))) | (np.isnan(a) & np.isnan(b))
    d = np.abs(np.where(same_inf, 0.0, a - b))
    assert not np.isnan(d).any(), (a, b)
    assert d.max(initial=0.0) <= tol, "max |delta| = %g (got %r want %r)" % (
        d.max(), a.ravel()[d.argmax()], b.ravel()[d.argmax()])


def _check(ctx, oracle, scn, cols, cat, off, sites=None):
    ns = scn["n_samples"]
    n_sites = (len(off) - 1) // ns
    got = ctx.posterior_batch(scn, cols, cat, off)
    sites = range(n_sites) if sites is None else sites
    want = _oracle_run(oracle, scn, cols, cat, off, sites)
    for s, w in zip(sites, want):
        _close(got["event_ln"][s], w["event_ln"])
        _close(got["marginal"][s], w["marginal"])
        assert np.array_equal(got["map_af"][s], w["map_af"], equal_nan=True), (s, got["map_af"][s], w["map_af"])
        assert np.array_equal(got["map_bias"][s], w["map_bias_cfg"]), (s, got["map_bias"][s], w["map_bias_cfg"])
    return got, want


@pytest.mark.parametrize("continuous", [False, True])
def test_c2_single_sample(ctx, oracle, continuous):
    scn = scenario.single_sample(continuous=continuous)
    cols, cat, off, _ = synth.make_pileups(300, [30], seed=synth.SEED_C2, artifact_frac=0.1)
    got, want = _check(ctx, oracle, scn, cols, cat, off)
    # some sites must be called present, some absent, some artifact for the test to mean anything
    post = got["event_ln"] - got["marginal"][:, None]
    assert (post.argmax(1) == 0).sum() > 100 and (post.argmax(1) != 0).sum() > 10
    assert (got["map_bias"] >= 0).any()


def test_c2_ragged_depths_and_empty(ctx, oracle):
    scn = scenario.single_sample(continuous=True)
    cols, cat, off, _ = synth.make_pileups(200, [30], seed=7, ragged=True)
    assert (np.diff(off) == 0).any()
    _check(ctx, oracle, scn, cols, cat, off)


def test_tumor_normal_indel(ctx, oracle):
    scn = scenario.tumor_normal(purity=0.75, indel=True)
    cols, cat, off, _ = synth.make_pileups(
        24, [40, 100], seed=synth.SEED_C3, kind="indel",
        af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]], purity=[1.0, 0.75])
    _check(ctx, oracle, scn, cols, cat, off)


def test_tumor_normal_snv_ragged(ctx, oracle):
    scn = scenario.tumor_normal(purity=0.6, indel=False, tumor_resolution=31)
    cols, cat, off, _ = synth.make_pileups(
        24, [20, 45], seed=11, kind="snv", ragged=True, artifact_frac=0.2,
        af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]], purity=[1.0, 0.6])
    _check(ctx, oracle, scn, cols, cat, off)


def test_deep_pileups_beyond_max_depth(ctx, oracle):
    """600 observations per pileup (default --max-depth is 200): shared-memory staging scales."""
    scn = scenario.single_sample(continuous=True, resolution=21)
    cols, cat, off, _ = synth.make_pileups(6, [600], seed=13)
    _check(ctx, oracle, scn, cols, cat, off)


def test_trio_with_branching_tree(ctx, oracle):
    scn = scenario.trio(with_biases=True)
    cols, cat, off, _ = synth.make_pileups(60, [30, 30, 30], seed=synth.SEED_C5)
    _check(ctx, oracle, scn, cols, cat, off)


def test_unquantised_extreme_probabilities(ctx, oracle):
    """Allele probabilities down to e^-2000 and MAPQ 255 (prob_mismapping = ln 0): the scaled
    linear product leaves fp64 range and the log-space path must take over."""
    scn = scenario.single_sample(continuous=False)
    cols, cat, off, _ = synth.make_pileups(40, [12], seed=3, quantise=False)
    n = len(cat)
    rng = np.random.default_rng(0)
    k = rng.random(n) < 0.5
    cols["prob_ref"] = np.where(k & (cols["prob_alt"] > cols["prob_ref"]), -2000.0, cols["prob_ref"])
    cols["prob_alt"] = np.where(k & (cols["prob_alt"] < cols["prob_ref"]), -1500.0, cols["prob_alt"])
    cols["prob_mapping"] = np.where(rng.random(n) < 0.5, 0.0, cols["prob_mapping"])
    cols["prob_mismapping"] = synth._ln1mexp(cols["prob_mapping"])
    cols["prob_missed_allele"] = np.logaddexp(cols["prob_ref"], cols["prob_alt"]) - math.log(2.0)
    _check(ctx, oracle, scn, cols, cat, off)


def test_all_other_indel_ops_drops_unbiased_events(ctx, oracle):
    """Quirk Q16: if every observation is IndelOperations::Other, DivIndelBias::None is impossible
    and only artifact events survive."""
    scn = scenario.single_sample(continuous=False, snv=False)
    cols, cat, off, _ = synth.make_pileups(6, [25], seed=5, kind="indel", af_sets=[[(0.5, 1.0)]])
    cat = (cat & ~np.uint16(3 << 8)) | np.uint16(1 << 8)
    got, _ = _check(ctx, oracle, scn, cols, cat, off)
    unbiased = [i for i, e in enumerate(scn["events"]) if e["biases"] == [0]]
    assert np.isneginf(got["event_ln"][:, unbiased]).all()


def test_likelihood_items_and_reference_identities(ctx, oracle):
    """vlr_likelihood_batch vs the oracle, plus the reference's own identities
    (likelihood.rs:267-386) evaluated on the GPU."""
    cols, cat, off, _ = synth.make_pileups(20, [15, 25], seed=9, kind="indel",
                                           af_sets=[[(0, .5), (.5, .5)], [(0.2, 1.0)]])
    biases = [scenario.bias()] + scenario.artifact_biases(True, True, True, True, True)
    rng = np.random.default_rng(1)
    items = []
    for _ in range(400):
        items.append(dict(site=int(rng.integers(0, 20)), sample=int(rng.integers(0, 2)),
                          bias=int(rng.integers(0, len(biases))), contaminated=int(rng.integers(0, 2)),
                          af=float(rng.choice([0.0, 0.13, 0.5, 1.0, rng.random()])),
                          af_secondary=float(rng.choice([0.0, 0.5, 1.0, rng.random()])),
                          purity=float(rng.uniform(0.3, 1.0)), other_rate=float(rng.choice([0.0, 0.25, 0.6]))))
    got = ctx.likelihood_batch(20, 2, cols, cat, off, biases, items)
    for it, g in zip(items, got):
        k = it["site"] * 2 + it["sample"]
        pv = oracle.PileupView(cols, cat, int(off[k]), int(off[k + 1]))
        b = oracle.make_biases(**biases[it["bias"]])
        b.other_rate = it["other_rate"]
        if it["contaminated"]:
            w = oracle.likelihood_contaminated(pv, it["purity"], it["af"], b, it["af_secondary"], b)
        else:
            w = oracle.likelihood_single(pv, it["af"], b)
        if math.isnan(w):
            assert math.isnan(g)     # Strand::None x strand bias is unreachable!() in the reference
        else:
            _close(g, w)
    # reference identities: 10 alt-zero reads at AF 0 -> sum of prob_any; AF .5 maximises 5+5 reads
    ninf = -math.inf

    def obs_rows(rows):
        c = {n: [] for n in synth.OBS_COLUMNS}
        for pm, pa, pr in rows:
            c["prob_mapping"].append(pm); c["prob_mismapping"].append(float(synth._ln1mexp(pm)))
            c["prob_alt"].append(pa); c["prob_ref"].append(pr)
            c["prob_missed_allele"].append(float(np.logaddexp(pr, pa)) - math.log(2.0))
            c["prob_sample_alt"].append(0.0); c["prob_double_overlap"].append(0.0)
            c["prob_single_overlap"].append(ninf); c["prob_hit_base"].append(math.log(0.01))
        n = len(rows)
        return ({k: np.array(v) for k, v in c.items()},
                synth.pack_cat([2] * n, [8] * n, [1] * n, [0] * n, [2] * n, [1] * n))
    c1, k1 = obs_rows([(0.0, ninf, 0.0)] * 10)
    none = [scenario.bias()]
    any_sum = 10 * (math.log(0.5) + math.log(0.5) + math.log(0.01))
    g = ctx.likelihood_batch(1, 1, c1, k1, [0, 10], none,
                             [dict(site=0, sample=0, bias=0, af=0.0),
                              dict(site=0, sample=0, bias=0, af=0.0, contaminated=1, af_secondary=0.0, purity=1.0)])
    _close(g, [any_sum, any_sum], tol=1e-9)
    c2, k2 = obs_rows([(0.0, 0.0, ninf)] * 5 + [(0.0, ninf, 0.0)] * 5)
    afs = [0.5] + [a for a in np.linspace(0, 1, 10)]
    g = ctx.likelihood_batch(1, 1, c2, k2, [0, 10], none,
                             [dict(site=0, sample=0, bias=0, af=float(a), contaminated=1, af_secondary=0.0,
                                   purity=1.0) for a in afs])
    assert (g[0] > g[1:]).all()


def test_posterior_errors_are_loud(ctx):
    scn = scenario.single_sample()
    cols, cat, off, _ = synth.make_pileups(1, [5000], seed=1)
    with pytest.raises(P.VlrError):
        ctx.posterior_batch(scn, cols, cat, off)  # 5000 observations > 4096 per pileup
    bad = scenario.single_sample()
    bad["nodes"][0]["sample"] = 5
    cols, cat, off, _ = synth.make_pileups(2, [10], seed=1)
    with pytest.raises(P.VlrError):
        ctx.posterior_batch(bad, cols, cat, off)


def _mendel_prior(afs, het=0.001):
    """Pedigree-shaped prior over diploid genotypes (child, father, mother): population prior of
    the parents and Mendelian transmission to the child -- a stand-in for Prior::compute
    (prior.rs:288-431, 674-786); what matters here is that it is a non-flat, AF-vector dependent
    ln prior evaluated by the host."""
    def pop(a):
        return {0.0: math.log(1 - 1.5 * het), 0.5: math.log(het), 1.0: math.log(0.5 * het)}[a]

    def gam(a):
        return {0.0: (1.0, 0.0), 0.5: (0.5, 0.5), 1.0: (0.0, 1.0)}[a]
    c, f, m = afs
    pf, pm = gam(f), gam(m)
    n_alt = int(round(2 * c))
    p = sum(pf[i] * pm[j] for i in range(2) for j in range(2) if i + j == n_alt)
    p = 0.999 * p + 0.001 / 3.0  # de-novo mass keeps every genotype possible
    return pop(f) + pop(m) + math.log(p)


def test_trio_with_prior_table(ctx, oracle):
    """Non-flat prior: the host evaluates its Prior for the leaves vlr_model_leaves reports and
    passes the table; event posteriors and MAP genotypes must follow the oracle's callback."""
    scn = dict(scenario.trio(with_biases=True))
    ev, afs = ctx.model_leaves(scn)
    assert len(ev) == 1 + 2 * (9 + 9 + 1) and afs.shape[1] == 3  # absent + (event, artifact event) x leaves
    scn["prior_ln"] = np.array([_mendel_prior(list(a)) for a in afs])
    cols, cat, off, _ = synth.make_pileups(96, [25, 30, 20], seed=5150, kind="snv",
                                           af_sets=[[(0, .5), (.5, .35), (1, .15)]] * 3, ragged=True,
                                           artifact_frac=0.05)
    got = ctx.posterior_batch(scn, cols, cat, off)
    flat = ctx.posterior_batch({k: v for k, v in scn.items() if k != "prior_ln"}, cols, cat, off)
    assert np.nanmax(np.abs(got["event_ln"] - flat["event_ln"])) > 1.0  # the prior does matter
    spec = _oracle_spec(oracle, scn)
    for s in range(96):
        pv = [oracle.PileupView(cols, cat, int(off[s * 3 + k]), int(off[s * 3 + k + 1])) for k in range(3)]
        w = oracle.model_compute(spec, pv, prior=_mendel_prior)
        _close(got["event_ln"][s], w["event_ln"])
        _close(got["marginal"][s], w["marginal"])
        assert np.array_equal(got["map_af"][s], w["map_af"], equal_nan=True), (s, got["map_af"][s], w["map_af"])


def test_prior_table_rejected_for_ranges(ctx):
    scn = dict(scenario.single_sample(continuous=True))
    scn["prior_ln"] = np.zeros(4)
    cols, cat, off, _ = synth.make_pileups(4, [10], seed=1, kind="snv")
    with pytest.raises(P.VlrError):
        ctx.posterior_batch(scn, cols, cat, off)


@pytest.mark.parametrize("snv", [True, False])
@pytest.mark.parametrize("continuous", [True, False])
def test_real_pileups_from_reference_records(ctx, oracle, snv, continuous):
    """Part 2 on REAL observation records (written by the reference's `preprocess`, extracted from
    its testcases into tests/golden/obs_records.json and decoded by the library's codec): GPU
    posteriors, marginals and MAP allele frequencies against the oracle."""
    import json
    import os
    from libprosic_b200 import obs_codec as lc
    recs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obs_records.json")))
    names = lc.COLUMNS
    parts = {c: [] for c in names}
    cats, off = [], [0]
    for r in recs:
        cols, cat = lc.decode_record(r["fields"])
        for c in names:
            parts[c].append(cols[c])
        cats.append(cat)
        off.append(off[-1] + len(cat))
    cols = {c: np.concatenate(parts[c]) for c in names}
    scn = scenario.single_sample(continuous=continuous, snv=snv)
    got, want = _check(ctx, oracle, scn, cols, np.concatenate(cats), np.array(off, np.int64))
    # the records are real calls: most sites must put their mass on a present-variant event
    present = sum(1 for w in want if np.argmax(w["event_ln"]) != 0)
    assert present >= len(recs) // 2


def test_call_record_values(ctx, oracle):
    """PROB_* (PHRED f32), AF (f32) and DP of the final call record (calling/variants/mod.rs:185-187,
    309-333): bit-identical f32 / int values for the same posterior outputs, float-missing for
    sites without observations."""
    from oracle import call_records as ocr
    scn = scenario.tumor_normal(purity=0.75, indel=True)
    cols, cat, off, _ = synth.make_pileups(64, [12, 25], seed=99, kind="indel", ragged=True,
                                           af_sets=[[(0, .6), (.5, .4)], [(0, .4), (.3, .6)]], purity=[1.0, 0.75])
    off = off.copy()
    post = ctx.posterior_batch(scn, cols, cat, off)
    got = ctx.call_records(scn, post, cols["prob_mapping"], off)
    assert got["names"] == ["absent", "somatic_tumor", "somatic_normal", "germline_het", "germline_hom", "artifact"]
    is_art = [any(scn["biases"][b][k] for k in ("strand", "orientation", "read_position", "softclip", "divindel"))
              for b in (ev["biases"][0] for ev in scn["events"])]
    n_missing = 0
    for s in range(64):
        pms = [cols["prob_mapping"][off[s * 2 + m]:off[s * 2 + m + 1]] for m in range(2)]
        prob, af, dp = ocr.call_record(post["event_ln"][s], post["marginal"][s], post["map_af"][s], is_art, pms)
        assert np.array_equal(got["dp"][s], dp), (s, got["dp"][s], dp)
        assert np.array_equal(got["af"][s].view(np.uint32), af.view(np.uint32)), s
        if all(len(p) == 0 for p in pms):
            n_missing += 1
            assert (got["prob"][s].view(np.uint32) == 0x7F800001).all(), s
            continue
        # f32 rounding of values that agree to 1e-13: allow the last f32 bit
        a, b = got["prob"][s].astype(np.float64), prob.astype(np.float64)
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), s
        assert np.all(np.abs(a[fin] - b[fin]) <= 1.2e-7 * np.maximum(1.0, np.abs(b[fin]))), (s, a, b)
    assert np.mean(got["prob"][np.isfinite(got["prob"])] >= 0) == 1.0


def test_chunked_pipeline_and_f32_entry(ctx, oracle):
    """The host-buffer entry points stream the batch in chunks (H2D of chunk c+1 overlaps the
    kernels of chunk c): a batch large enough for several chunks, ragged depths, must give the
    same numbers as the oracle on a sample of sites, and the compact f32 entry (derived columns on
    the device) the same as the f64 entry."""
    scn = scenario.tumor_normal(purity=0.75, indel=True)
    cols, cat, off, _ = synth.make_pileups(6000, [40, 100], seed=321, kind="indel", ragged=True,
                                           af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]],
                                           purity=[1.0, 0.75], artifact_frac=0.05)
    assert len(cat) * 74 > 2 * 24 * 2 ** 20  # at least three chunks of ~24 MB
    got = ctx.posterior_batch(scn, cols, cat, off)
    sites = list(range(0, 6000, 499)) + [5999]
    want = _oracle_run(oracle, scn, cols, cat, off, sites)
    for s, w in zip(sites, want):
        _close(got["event_ln"][s], w["event_ln"])
        assert np.array_equal(got["map_af"][s], w["map_af"], equal_nan=True), s
    cols32 = {k: v.astype(np.float32) for k, v in cols.items()}
    for k in ("prob_mapping", "prob_alt", "prob_ref", "prob_missed_allele", "prob_sample_alt",
              "prob_double_overlap", "prob_hit_base"):
        assert np.array_equal(cols32[k].astype(np.float64), cols[k], equal_nan=True), k  # quantised inputs
    got32 = ctx.posterior_batch_f32(scn, cols32, cat, off)
    _close(got32["event_ln"], got["event_ln"], 1e-9)
    _close(got32["marginal"], got["marginal"], 1e-9)
    assert np.array_equal(got32["map_af"], got["map_af"], equal_nan=True)
    assert np.array_equal(got32["map_bias"], got["map_bias"])


def test_wire_entry_decodes_on_the_device(ctx, oracle):
    """vlr_posterior_batch_wire: the records as `preprocess` stores them (bincode words per field)
    are uploaded in chunks and decoded by a kernel; the posteriors are those of the f64 entry on the
    host-decoded columns, bit for bit, for ragged two-sample pileups spanning several chunks."""
    from libprosic_b200 import obs_codec as lc
    scn = scenario.tumor_normal(purity=0.75, indel=True)
    cols, cat, off, _ = synth.make_pileups(9000, [40, 100], seed=77, kind="indel", ragged=True,
                                           af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]],
                                           purity=[1.0, 0.75], artifact_frac=0.05)
    wire = lc.encode_batch(cols, cat, off)
    assert wire.nbytes > 2 * 32 * 2 ** 20  # at least three chunks
    got = ctx.posterior_batch_wire(scn, wire)
    dcols, dcat, doff = lc.decode_batch(ctx, wire)
    assert np.array_equal(doff, off) and np.array_equal(dcat, cat)
    for k in lc.COLUMNS:
        if k not in ("prob_mismapping", "prob_single_overlap"):
            assert np.array_equal(dcols[k], cols[k], equal_nan=True), k   # synth pileups are MiniLogProb-quantised
    want = ctx.posterior_batch(scn, dcols, dcat, doff)
    assert np.array_equal(got["event_ln"], want["event_ln"], equal_nan=True)
    assert np.array_equal(got["marginal"], want["marginal"], equal_nan=True)
    assert np.array_equal(got["map_af"], want["map_af"], equal_nan=True)
    assert np.array_equal(got["map_bias"], want["map_bias"])
    sites = [0, 4500, 8999]
    for s, w in zip(sites, _oracle_run(oracle, scn, cols, cat, off, sites)):
        _close(got["event_ln"][s], w["event_ln"])


def _random_scenario(rng, n_samples):
    """Random event universe: 2-6 events, each 1-3 tree roots; a root is a chain over the samples in
    random order whose nodes are random Sets / Ranges (some chains end in a fan-out of two children,
    some carry a FALSE or SKIP node); artifact events over a random subset of bias components."""
    b = scenario._Builder()
    grid = [0.0, 0.1, 0.25, 0.5, 0.75, 1.0]

    def spectrum(s):
        if rng.random() < 0.45:
            lo, hi = sorted(rng.choice(len(grid), 2, replace=False))
            return (s, scenario.NODE_RANGE, [grid[lo], grid[hi], float(rng.random() < 0.5), float(rng.random() < 0.5)])
        k = int(rng.integers(1, 4))
        return (s, scenario.NODE_SET, sorted(float(v) for v in rng.choice(grid, k, replace=False)))

    def root():
        order = list(rng.permutation(n_samples))
        if n_samples >= 2 and rng.random() < 0.3:   # fan-out below the first sample
            first = spectrum(order[0])
            kids = [b.chain([spectrum(s) for s in order[1:]]) for _ in range(2)]
            if rng.random() < 0.3:
                kids.append(b.node(scenario.NODE_FALSE))
            top = b.node(first[1], first[0], first[2], kids)
            return b.node(scenario.NODE_SKIP, 0, [], [top]) if rng.random() < 0.3 else top
        return b.chain([spectrum(s) for s in order])

    named = [("e%d" % k, [root() for _ in range(int(rng.integers(1, 4)))]) for k in range(int(rng.integers(1, 6)))]
    flags = [bool(rng.random() < 0.6) for _ in range(5)]
    arts = scenario.artifact_biases(*flags) if any(flags) else []
    events, biases = scenario._with_events(b, named, n_samples, arts)
    cont = [0] * n_samples
    by = [0] * n_samples
    purity = [1.0] * n_samples
    if n_samples >= 2 and rng.random() < 0.5:
        c = int(rng.integers(0, n_samples))
        cont[c], by[c], purity[c] = 1, int((c + 1 + rng.integers(0, n_samples - 1)) % n_samples), float(rng.choice([0.3, 0.75, 0.9]))
    return dict(n_samples=n_samples, resolution=[int(rng.choice([5, 9, 20, 100])) for _ in range(n_samples)],
                contaminated=cont, by=by, purity=purity, nodes=b.nodes, events=events, biases=biases)


@pytest.mark.parametrize("seed", range(24))
def test_random_models(ctx, oracle, seed):
    """Fuzz of the model flattening, the term-major integration with its segmented flush (events of
    1 .. several hundred terms, several events per warp round, gated-out configurations, FALSE
    branches), every kernel instantiation (1-4 samples, with and without contamination, one-warp
    and CTA teams) and ragged / empty pileups."""
    rng = np.random.default_rng(1000 + seed)
    ns = int(rng.integers(1, 5))
    scn = _random_scenario(rng, ns)
    depths = [int(rng.choice([3, 12, 30, 70])) for _ in range(ns)]
    cols, cat, off, _ = synth.make_pileups(40, depths, seed=seed, kind=str(rng.choice(["snv", "indel"])),
                                           ragged=bool(rng.random() < 0.5), artifact_frac=0.2)
    _check(ctx, oracle, scn, cols, cat, off)


def _random_prior_program(rng, ns):
    """A random vlr_prior_prog (dict form): uniform / somatic / germline samples, somatic terms of
    every kind, a random germline table with some impossible count vectors."""
    kind, ploidy, som_kind, parent, ln_rate, som_zero, uni = [], [], [], [], [], [], []
    for s in range(ns):
        k = int(rng.choice([0, 1, 1, 2]))
        kind.append(k)
        ploidy.append(-1 if k == 0 and rng.random() < 0.5 else int(rng.choice([1, 2, 2, 3])))
        uni.append([(1.0, 0.0, 1.0, float(rng.random() < 0.3), 0.0)] if k == 0 else [])
        sk = 0 if k == 2 else int(rng.choice([0, 1, 1, 2, 3])) if ns > 1 else int(rng.choice([0, 1]))
        if k == 0:
            sk = 0
        som_kind.append(sk)
        parent.append(int((s + 1 + rng.integers(0, max(1, ns - 1))) % ns) if sk >= 2 else -1)
        ln_rate.append(float(np.log(rng.choice([1e-10, 1e-7, 1e-4]))))
        som_zero.append(float(-rng.random() * 1e-3))
    n_germ = 1
    for s in range(ns):
        n_germ *= 1 if (kind[s] == 0 or ploidy[s] <= 0) else ploidy[s] + 1
    germ_ln = (-20.0 * rng.random(n_germ)).tolist()
    for k in rng.choice(n_germ, n_germ // 4, replace=False):
        germ_ln[int(k)] = -np.inf
    return dict(kind=kind, ploidy=ploidy, som_kind=som_kind, parent=parent, ln_rate=ln_rate, som_zero=som_zero,
                ln_genome_size=float(np.log(3.5e9)), uni=uni, germ_ln=germ_ln)


@pytest.mark.parametrize("seed", range(16))
def test_random_models_with_prior_programs_on_ranges(ctx, oracle, seed):
    """Prior::compute on the device for universes with Ranges (vlr_prior_prog): random models whose
    tree paths bind Ranges and Sets, random prior programs (somatic mutation terms of every kind,
    germline tables, uniform samples); the oracle takes the host evaluation of the same program as
    its prior callback (prior.rs:288-431 restated in libprosic_b200/prior.py)."""
    from libprosic_b200 import prior as prior_mod
    rng = np.random.default_rng(5000 + seed)
    ns = int(rng.integers(1, 5))
    scn = _random_scenario(rng, ns)
    scn["resolution"] = [int(rng.choice([5, 9, 20])) for _ in range(ns)]
    prog = _random_prior_program(rng, ns)
    scn["prior_prog"] = prog
    scn["prior_fn"] = lambda afs: prior_mod.eval_program(prog, list(afs))
    depths = [int(rng.choice([3, 12, 30])) for _ in range(ns)]
    cols, cat, off, _ = synth.make_pileups(12, depths, seed=seed, kind=str(rng.choice(["snv", "indel"])),
                                           ragged=bool(rng.random() < 0.5), artifact_frac=0.2)
    _check(ctx, oracle, scn, cols, cat, off)
