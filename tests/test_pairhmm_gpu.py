"""GPU parity of K1 (pair-HMM forward) against the CPU oracle, through the C ABI
(vlr_pairhmm_batch).  Tolerance: |delta| <= 1e-6 in natural-log space (BASELINE.json north_star).
"""
import numpy as np
import pytest

import libprosic_b200 as P
from libprosic_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-6


def _gap_sets():
    return [P.GapParams(),  # CLI defaults 2.8e-6 / 5.1e-6 / 0 / 0 (cli.rs:207-230)
            P.GapParams(0.02, 0.03, 0.3, 0.3),   # C4-style raised rates, symmetric extension
            P.GapParams(1e-3, 2e-3, 0.1, 0.1)]


def _compare(got, want, tol=TOL):
    got, want = np.asarray(got), np.asarray(want)
    both_inf = np.isneginf(got) & np.isneginf(want)
    diff = np.abs(np.where(both_inf, 0.0, got - want))
    assert not np.isnan(got).any(), "NaN from the GPU path"
    assert np.nanmax(diff) <= tol, "max |delta| = %g at %d (got %r want %r)" % (
        np.nanmax(diff), int(np.nanargmax(diff)), got[np.nanargmax(diff)], want[np.nanargmax(diff)])


@pytest.mark.parametrize("flags", [P.HMM_DEFAULT, P.HMM_SUM3_APPROX, 0])
def test_c3_shape_unbanded(ctx, oracle, flags):
    w = synth.make_pairhmm_workload(n_sites=6, reads_per_site=16, seed=synth.SEED_C3)
    gap = P.GapParams()
    got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                            w["read_off"], gap, w["pair_hap"], w["pair_read"], flags=flags)
    want = np.array([oracle.pairhmm_prob_related(
        w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]],
        w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
        w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]], gap.as_ln_array(), -1, flags)
        for h, r in zip(w["pair_hap"], w["pair_read"])])
    _compare(got, want)


@pytest.mark.parametrize("gi", [0, 1, 2])
def test_ragged_unbanded(ctx, oracle, gi):
    gap = _gap_sets()[gi]
    w = synth.make_ragged_pairs(300, seed=100 + gi, max_ly=248, max_lx=600)
    got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                            w["read_off"], gap)
    want = oracle.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                                w["read_off"], gap.as_ln_array())
    _compare(got, want)


@pytest.mark.parametrize("gi", [0, 1])
def test_ragged_banded(ctx, oracle, gi):
    """Band k = dist + 2 as the reference runs it (edit_distance.rs:180-182)."""
    gap = _gap_sets()[gi]
    w = synth.make_ragged_pairs(300, seed=200 + gi, max_ly=128, max_lx=330)
    n = len(w["hap_off"]) - 1
    med = np.empty(n, dtype=np.int32)
    for k in range(n):
        hit = oracle.best_hit(w["hap_bytes"][w["hap_off"][k]:w["hap_off"][k + 1]],
                              w["read_bases"][w["read_off"][k]:w["read_off"][k + 1]])
        med[k] = hit["dist"] + 2
    med[::17] = -1  # a few unbanded pairs inside a banded batch
    got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                            w["read_off"], gap, max_edit_dist=med)
    want = oracle.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                                w["read_off"], gap.as_ln_array(), med)
    _compare(got, want)


def test_underflow_goes_through_log_space_pass(ctx, oracle):
    """Unrelated read/haplotype with Q60 bases: P ~ 1e-900, outside scaled fp64 range."""
    rng = np.random.default_rng(5)
    n = 8
    hap = np.frombuffer(b"A" * 300, dtype=np.uint8)
    haps = np.tile(hap, n)
    reads = np.tile(np.frombuffer(b"C" * 150, dtype=np.uint8), n)
    quals = np.full(n * 150, 60, dtype=np.uint8)
    quals[:150] = rng.integers(50, 61, 150)
    hap_off = np.arange(0, (n + 1) * 300, 300, dtype=np.int64)
    read_off = np.arange(0, (n + 1) * 150, 150, dtype=np.int64)
    gap = P.GapParams()
    got = ctx.pairhmm_batch(haps, hap_off, reads, quals, read_off, gap)
    want = oracle.pairhmm_batch(haps, hap_off, reads, quals, read_off, gap.as_ln_array())
    assert (want < -1500).all()
    _compare(got, want, tol=1e-6 * 10)  # |ln p| ~ 2000: allow relative 5e-9


def test_edge_cases(ctx, oracle):
    gap = P.GapParams()
    # single-base read, single-base haplotype, Q0 base (ln(1-eps) = -inf), read longer than hap
    cases = [(b"A", b"A", [30]), (b"A", b"C", [30]), (b"ACGT", b"G", [0]), (b"AC", b"ACGTACGT", [20] * 8),
             (b"acgtacgtac", b"ACGTACGTAC", [40] * 10), (b"NNNN", b"NN", [10, 10])]
    haps = np.frombuffer(b"".join(c[0] for c in cases), dtype=np.uint8)
    reads = np.frombuffer(b"".join(c[1] for c in cases), dtype=np.uint8)
    quals = np.array(sum((c[2] for c in cases), []), dtype=np.uint8)
    hap_off = np.concatenate([[0], np.cumsum([len(c[0]) for c in cases])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(c[1]) for c in cases])]).astype(np.int64)
    got = ctx.pairhmm_batch(haps, hap_off, reads, quals, read_off, gap)
    want = oracle.pairhmm_batch(haps, hap_off, reads, quals, read_off, gap.as_ln_array())
    _compare(got, want)


def test_errors_are_loud(ctx):
    gap = P.GapParams()
    hap = np.frombuffer(b"ACGT" * 100, dtype=np.uint8)
    read = np.frombuffer(b"A" * 300, dtype=np.uint8)
    with pytest.raises(P.VlrError):
        ctx.pairhmm_batch(hap, np.array([0, 400]), read, np.full(300, 30, np.uint8),
                          np.array([300, 0]), gap, max_edit_dist=np.array([5], np.int32))  # offsets not monotone
    with pytest.raises(P.VlrError):
        ctx.pairhmm_batch(hap, np.array([0, 400]), read[:10], np.full(10, 30, np.uint8),
                          np.array([0, 10]), gap, pair_hap=np.array([3], np.int32),
                          pair_read=np.array([0], np.int32))


@pytest.mark.parametrize("gi,flags", [(0, P.HMM_DEFAULT), (1, P.HMM_DEFAULT), (1, 0)])
def test_long_reads_strips(ctx, oracle, gi, flags):
    """Config C4 shape in miniature: read windows of 250..1400 rows against longer haplotypes
    (strip-wise sweep with per-strip renormalisation), mixed with short pairs in one batch."""
    gap = _gap_sets()[gi]
    rng = np.random.default_rng(40 + gi)
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    haps, reads, quals = [], [], []
    for k in range(24):
        ly = int(rng.integers(249, 1400)) if k % 4 else int(rng.integers(20, 200))
        lx = ly + int(rng.integers(10, 600))
        hap = A[rng.integers(0, 4, lx)]
        s = int(rng.integers(0, lx - ly))
        read = hap[s:s + ly].copy()
        q = np.clip(np.rint(rng.normal(15, 4, ly)), 3, 30).astype(np.uint8)   # long-read error model
        sub = rng.random(ly) < 0.08
        read[sub] = A[rng.integers(0, 4, int(sub.sum()))]
        keep = rng.random(ly) > 0.02
        read = read[keep]
        q = q[keep]
        ins = np.sort(rng.integers(0, len(read), int(0.02 * len(read))))
        read = np.insert(read, ins, A[rng.integers(0, 4, len(ins))])
        q = np.insert(q, ins, 10)
        if k == 5:
            read = A[rng.integers(0, 4, len(read))]  # unrelated: deep underflow -> log-space pass
        haps.append(hap); reads.append(read); quals.append(q)
    hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    hb, rb, rq = np.concatenate(haps), np.concatenate(reads), np.concatenate(quals)
    got = ctx.pairhmm_batch(hb, hap_off, rb, rq, read_off, gap, flags=flags)
    want = oracle.pairhmm_batch(hb, hap_off, rb, rq, read_off, gap.as_ln_array(), flags=flags)
    assert (want < -700).any()   # beyond plain fp64 range: renormalisation is exercised
    _compare(got, want, tol=1e-6)


@pytest.mark.parametrize("flags", [P.HMM_DEFAULT, 0])
def test_non_acgt_bases(ctx, oracle, flags):
    """Reads and haplotypes with N / IUPAC codes and lower case: read-side codes route the pair to
    the byte-compare kernel (the emission LUT has columns for A, C, G, T and 'anything else');
    N == N is a match (pairhmm.rs:190-191)."""
    rng = np.random.default_rng(77)
    alphabet = np.frombuffer(b"ACGTNRYacgtn", dtype=np.uint8)
    read_alpha = np.frombuffer(b"ACGTNR", dtype=np.uint8)
    haps, reads, quals = [], [], []
    for k in range(96):
        lx, ly = int(rng.integers(40, 330)), int(rng.integers(5, 200))
        pw = np.array([.22, .22, .22, .22, .03, .01, .01, .02, .02, .01, .01, .01])
        hap = alphabet[rng.choice(len(alphabet), lx, p=pw / pw.sum())]
        if k % 3 == 0:    # read bases all ACGT, haplotype with codes: stays on the LUT kernel
            read = read_alpha[rng.integers(0, 4, ly)]
        else:
            read = read_alpha[rng.choice(6, ly, p=[.23, .23, .23, .23, .06, .02])]
        if ly <= lx and k % 2 == 0:  # make the read resemble the haplotype
            s = int(rng.integers(0, lx - ly + 1))
            up = hap[s:s + ly].copy()
            up[(up >= 97)] -= 32
            keep = rng.random(ly) < 0.9
            read = np.where(keep, up, read)
        haps.append(hap); reads.append(read)
        quals.append(rng.integers(2, 42, ly).astype(np.uint8))
    hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    gap = P.GapParams(1e-3, 2e-3, 0.1, 0.1)
    got = ctx.pairhmm_batch(np.concatenate(haps), hap_off, np.concatenate(reads), np.concatenate(quals),
                            read_off, gap, flags=flags)
    want = np.array([oracle.pairhmm_prob_related(haps[k], reads[k], quals[k], gap.as_ln_array(), -1, flags)
                     for k in range(96)])
    _compare(got, want)


def _band_cases(rng, n, kind):
    """(haplotype window, read, qualities, band) the way allele support produces them: the read is a
    copy of a stretch of the window with a few edits, the window extends a little beyond the hit,
    the band is a little more than the number of edits.  kind: 'random' sequence, 'homopolymer',
    'repeat' (short tandem repeat) -- the low-complexity kinds keep many diagonals alive."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    haps, reads, quals, med = [], [], [], []
    for _ in range(n):
        ly = int(rng.integers(20, 200))
        extra = int(rng.integers(0, 40))
        lx = ly + extra
        if kind == "random":
            hap = acgt[rng.integers(0, 4, lx)]
        elif kind == "homopolymer":
            hap = np.full(lx, acgt[rng.integers(0, 4)], np.uint8)
            hap[rng.random(lx) < 0.03] = acgt[rng.integers(0, 4)]
        else:
            unit = acgt[rng.integers(0, 4, int(rng.integers(2, 5)))]
            hap = np.tile(unit, lx // len(unit) + 1)[:lx].copy()
        start = int(rng.integers(0, extra + 1))
        read = list(hap[start:start + ly])
        n_edits = int(rng.integers(0, 9))
        for _e in range(n_edits):
            pos = int(rng.integers(0, len(read)))
            op = rng.random()
            if op < 0.5:
                read[pos] = int(acgt[rng.integers(0, 4)])
            elif op < 0.75 and len(read) > 10:
                del read[pos]
            else:
                read.insert(pos, int(acgt[rng.integers(0, 4)]))
        read = np.array(read[:248], np.uint8)
        haps.append(hap)
        reads.append(read)
        quals.append(np.clip(np.rint(rng.normal(30, 8, len(read))), 2, 41).astype(np.uint8))
        med.append(int(rng.integers(0, 14)))
    hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    return (np.concatenate(haps), hap_off, np.concatenate(reads), np.concatenate(quals), read_off,
            np.array(med, np.int32))


@pytest.mark.parametrize("kind", ["random", "homopolymer", "repeat"])
@pytest.mark.parametrize("flags", [P.HMM_DEFAULT, 0])
def test_band_only_kernel(ctx, oracle, kind, flags):
    """pairhmm_band.cuh (diagonals as lanes, only the band is computed) against the oracle's full
    matrix with the reference's pruning rule: bands of 3 .. 96 diagonals (one to three diagonals
    per lane), windows from len_y to len_y + 40, low-complexity sequence where most of the band
    stays alive.  The kernel leaves out mass that needs k + 1 gaps in one direction: the observed
    difference stays far below the 1e-6 bar."""
    rng = np.random.default_rng({"random": 1, "homopolymer": 2, "repeat": 3}[kind] + (flags != 0))
    hb, ho, rb, rq, ro, med = _band_cases(rng, 400, kind)
    gap = P.GapParams()
    got = ctx.pairhmm_batch(hb, ho, rb, rq, ro, gap, max_edit_dist=med, flags=flags)
    want = oracle.pairhmm_batch(hb, ho, rb, rq, ro, gap.as_ln_array(), med, flags)
    _compare(got, want, tol=1e-7)


@pytest.mark.parametrize("gi,flags", [(0, P.HMM_SUM3_APPROX3), (0, P.HMM_SUM3_APPROX), (0, 0),
                                      (1, P.HMM_SUM3_APPROX3), (2, P.HMM_SUM3_APPROX | P.HMM_PATH_CLOSE),
                                      (1, P.HMM_SUM3_APPROX3 | P.HMM_PATH_CLOSE)])
def test_literal_kernel_is_bit_identical_to_the_oracle(ctx, oracle, gi, flags):
    """VLR_HMM_LITERAL: natural-log space in the reference's operation order with the deterministic
    libm -- every result equals the oracle's BIT FOR BIT (banded with the reference's cell set,
    unbanded, gap extension, any alphabet, Q0 bases).  This is the evaluation the fused allele
    support falls back to when prob_ref and prob_alt are nearly tied (mod.rs:303)."""
    gap = _gap_sets()[gi]
    w = synth.make_ragged_pairs(240, seed=700 + gi, max_ly=248, max_lx=420)
    n = len(w["hap_off"]) - 1
    med = np.full(n, -1, dtype=np.int32)
    for k in range(0, n, 2):  # every other pair banded as the reference does (dist + 2)
        hit = oracle.best_hit(w["hap_bytes"][w["hap_off"][k]:w["hap_off"][k + 1]],
                              w["read_bases"][w["read_off"][k]:w["read_off"][k + 1]])
        med[k] = hit["dist"] + 2
    got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                            w["read_off"], gap, max_edit_dist=med, flags=flags | P.HMM_LITERAL)
    want = oracle.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                                w["read_off"], gap.as_ln_array(), med, flags)
    assert np.array_equal(got.view(np.int64), want.view(np.int64)), \
        "first difference at %d: %r vs %r" % (int(np.argmax(got != want)), got[np.argmax(got != want)],
                                              want[np.argmax(got != want)])


def _long_banded(oracle, n, seed, unbanded_every):
    w = synth.make_ragged_pairs(n, seed=seed, min_ly=249, max_ly=700, min_lx=200, max_lx=900,
                                related_frac=0.85)
    med = np.empty(n, dtype=np.int32)
    for k in range(n):
        hit = oracle.best_hit(w["hap_bytes"][w["hap_off"][k]:w["hap_off"][k + 1]],
                              w["read_bases"][w["read_off"][k]:w["read_off"][k + 1]])
        med[k] = hit["dist"] + 2
    med[::unbanded_every] = -1
    return w, med


@pytest.mark.parametrize("gi,flags", [(0, P.HMM_DEFAULT), (1, P.HMM_DEFAULT), (2, 0), (0, P.HMM_SUM3_APPROX)])
def test_long_banded_windows(ctx, oracle, gi, flags):
    """Banded read windows of 249-700 rows (reachable through merged multi-locus regions and breakend
    alleles, bio's Myers::Long): the anti-diagonal log-space kernel (pairhmm_diag.cuh) with the
    reference's cell set; the unbanded entries of the same batch take the strip-wise kernel."""
    gap = _gap_sets()[gi]
    w, med = _long_banded(oracle, 40, 900 + gi, 9)
    got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                            w["read_off"], gap, max_edit_dist=med, flags=flags)
    want = oracle.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                                w["read_off"], gap.as_ln_array(), med, flags)
    _compare(got, want)
    assert np.isfinite(want).sum() > 20


@pytest.mark.parametrize("gi,flags", [(0, P.HMM_SUM3_APPROX3), (1, P.HMM_SUM3_APPROX3), (2, 0)])
def test_literal_long_windows_are_bit_identical_to_the_oracle(ctx, oracle, gi, flags):
    """VLR_HMM_LITERAL above 248 rows: the literal instantiation of the anti-diagonal kernel."""
    gap = _gap_sets()[gi]
    w, med = _long_banded(oracle, 14, 950 + gi, 7)
    got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                            w["read_off"], gap, max_edit_dist=med, flags=flags | P.HMM_LITERAL)
    want = oracle.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                                w["read_off"], gap.as_ln_array(), med, flags)
    assert np.array_equal(got.view(np.int64), want.view(np.int64)), \
        "first difference at %d: %r vs %r" % (int(np.argmax(got != want)), got[np.argmax(got != want)],
                                              want[np.argmax(got != want)])


@pytest.mark.parametrize("literal", [False, True])
def test_long_banded_edge_cases(ctx, oracle, literal):
    """Corners of the anti-diagonal kernel: the Myers limit (1024 rows) and beyond it, a haplotype
    shorter than the read, a five-base haplotype, band 0 on an exact copy, a read of N's whose band dies
    after a few rows (ln 0), lower-case haplotype, Q0 bases."""
    rng = np.random.default_rng(5150)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def related(ly, lx, n_sub):
        hap = acgt[rng.integers(0, 4, lx)]
        start = int(rng.integers(0, max(1, lx - ly + 1)))
        read = hap[start:start + ly].copy()
        if len(read) < ly:
            read = np.concatenate([read, acgt[rng.integers(0, 4, ly - len(read))]])
        pos = rng.choice(ly, n_sub, replace=False)
        read[pos] = acgt[rng.integers(0, 4, n_sub)]
        return hap, read
    cases = [related(1024, 1100, 6), related(1500, 1490, 12), related(300, 5, 0), related(260, 260, 0),
             (acgt[rng.integers(0, 4, 500)], np.full(400, ord("N"), np.uint8)), related(700, 760, 3)]
    cases[5] = (np.where(cases[5][0] < 96, cases[5][0] + 32, cases[5][0]).astype(np.uint8), cases[5][1])
    haps, reads = [c[0] for c in cases], [c[1] for c in cases]
    quals = [rng.integers(0, 42, len(r)).astype(np.uint8) for r in reads]
    quals[3][:] = 30
    hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    med = np.array([oracle.best_hit(h, r)["dist"] + 2 for h, r in zip(haps, reads)], dtype=np.int32)
    med[3] = 0
    med[4] = 3
    args = (np.concatenate(haps), hap_off, np.concatenate(reads), np.concatenate(quals), read_off)
    gap = P.GapParams()
    flags = P.HMM_DEFAULT | (P.HMM_LITERAL if literal else 0)
    got = ctx.pairhmm_batch(*args, gap, max_edit_dist=med, flags=flags)
    want = oracle.pairhmm_batch(*args, gap.as_ln_array(), med, P.HMM_DEFAULT)
    assert np.isneginf(want[4]) and np.isfinite(want[3]) and np.isfinite(want[0])
    if literal:
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (got, want)
    else:
        _compare(got, want)


def test_three_way_and_two_way_rules_differ_where_expected(ctx, oracle):
    """The two recollections of ln_sum3_exp_approx: results differ by at most ~e^-10 per cell
    (<= 1e-3 in ln on these reads) and both match their oracle variants."""
    w = synth.make_pairhmm_workload(n_sites=4, reads_per_site=16, seed=77)
    gap = P.GapParams()
    res = {}
    for flags in (P.HMM_SUM3_APPROX3, P.HMM_SUM3_APPROX, 0):
        got = ctx.pairhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                                w["read_off"], gap, w["pair_hap"], w["pair_read"], flags=flags)
        want = np.array([oracle.pairhmm_prob_related(
            w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]],
            w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
            w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]], gap.as_ln_array(), -1, flags)
            for h, r in zip(w["pair_hap"], w["pair_read"])])
        _compare(got, want)
        res[flags] = got
    assert np.max(np.abs(res[P.HMM_SUM3_APPROX3] - res[0])) <= 1e-3
    assert np.max(np.abs(res[P.HMM_SUM3_APPROX] - res[0])) <= 1e-3
