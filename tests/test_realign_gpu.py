"""GPU parity of K2 (Myers best hit + traceback) and of the fused allele-support pipeline
(vlr_besthit_batch / vlr_allele_support_batch) against the CPU oracle.
Integer outputs (distances, hit windows, indel operations) must be bit-exact; allele
log-probabilities within 1e-6."""
import numpy as np
import pytest

import libprosic_b200 as P
from libprosic_b200 import synth

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def test_besthit_ragged(ctx, oracle):
    w = synth.make_ragged_pairs(400, seed=31, max_ly=248, max_lx=500)
    got = ctx.besthit_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_off"])
    for k in range(400):
        want = oracle.best_hit(w["hap_bytes"][w["hap_off"][k]:w["hap_off"][k + 1]],
                               w["read_bases"][w["read_off"][k]:w["read_off"][k + 1]])
        assert [got["dist"][k], got["start"][k], got["end"][k], got["n_best"][k]] == \
               [want["dist"], want["start"], want["end"], want["n_best"]], k
        assert got["n_indel_ops"][k] == len(want["indel_ops"]), k
        n = len(want["indel_ops"])
        assert np.array_equal(got["indel_ops"][k][:n], want["indel_ops"][:n]), k


def test_besthit_long_windows(ctx, oracle):
    """Read windows beyond 256 rows (bio's Myers::Long, edit_distance.rs:42-46): the 16-word
    instantiation takes 257..1024 rows; a batch mixing short and long windows runs both kernels.
    Related pairs (a few edits), unrelated ones (distances of hundreds) and the truncated second
    scan must all agree with the oracle's explicit matrix, operations included."""
    w = synth.make_ragged_pairs(160, seed=77, min_ly=200, max_ly=640, min_lx=300, max_lx=1100)
    lens = np.diff(w["read_off"])
    assert (lens > 256).sum() > 60 and (lens <= 256).sum() > 10
    got = ctx.besthit_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_off"])
    for k in range(160):
        want = oracle.best_hit(w["hap_bytes"][w["hap_off"][k]:w["hap_off"][k + 1]],
                               w["read_bases"][w["read_off"][k]:w["read_off"][k + 1]])
        assert [got["dist"][k], got["start"][k], got["end"][k], got["n_best"][k]] == \
               [want["dist"], want["start"], want["end"], want["n_best"]], k
        assert got["n_indel_ops"][k] == len(want["indel_ops"]), k
        if not got["indel_overflow"][k]:
            n = len(want["indel_ops"])
            assert np.array_equal(got["indel_ops"][k][:n], want["indel_ops"][:n]), k
    with pytest.raises(P.VlrError):   # beyond 1024 rows: refused, not truncated
        big = synth.make_ragged_pairs(2, seed=3, min_ly=1100, max_ly=1200, min_lx=1300, max_lx=1400)
        ctx.besthit_batch(big["hap_bytes"], big["hap_off"], big["read_bases"], big["read_off"])


def test_besthit_repeats_and_ties(ctx, oracle):
    """Homopolymers / tandem repeats: many equally good end positions and indel placements."""
    rng = np.random.default_rng(4)
    haps, reads = [], []
    for _ in range(120):
        unit = ACGT[rng.integers(0, 4, int(rng.integers(1, 4)))]
        rep = np.tile(unit, int(rng.integers(5, 30)))
        hap = np.concatenate([ACGT[rng.integers(0, 4, 20)], rep, ACGT[rng.integers(0, 4, 20)]])
        s = int(rng.integers(0, 25))
        read = np.delete(hap[s:s + 60].copy(), rng.integers(5, 40, int(rng.integers(0, 3))))
        haps.append(hap)
        reads.append(read)
    hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    got = ctx.besthit_batch(np.concatenate(haps), hap_off, np.concatenate(reads), read_off)
    for k in range(120):
        want = oracle.best_hit(haps[k], reads[k])
        assert [got["dist"][k], got["start"][k], got["end"][k], got["n_best"][k]] == \
               [want["dist"], want["start"], want["end"], want["n_best"]], k
        assert np.array_equal(got["indel_ops"][k][:len(want["indel_ops"])], want["indel_ops"]), k


def _make_regions(rng, n_sites, reads_per_site, multi_alt=False, rl=100):
    """Indel candidates with reference-like windows: ref window 192 bp around the locus, alt
    window with the variant spliced in (insertions grow by ins_len: shrink_extra), read windows
    of <= 128 bases (rl = read length before trimming; the windows scale with it)."""
    haps, extra, reads, quals = [], [], [], []
    region_read, region_haps, region_hap_off, region_n_ref = [], [], [0], []
    W = rl - 4
    for _ in range(n_sites):
        contig = ACGT[rng.integers(0, 4, 5 * rl)]
        locus = 5 * rl // 2
        L = int(rng.integers(1, 25))
        is_ins = rng.random() < 0.5
        ref = contig[locus - W:locus + W]
        if is_ins:
            ins = ACGT[rng.integers(0, 4, L)]
            full_alt = np.concatenate([contig[:locus + 1], ins, contig[locus + 1:]])
            alt = full_alt[locus - W:locus + W + L]
            alt_extra = L
        else:
            full_alt = np.concatenate([contig[:locus + 1], contig[locus + 1 + L:]])
            alt = full_alt[locus - W:locus + W]
            alt_extra = 0
        h0 = len(haps)
        haps += [ref, alt]
        extra += [0, alt_extra]
        alt_ids = [h0 + 1]
        if multi_alt:  # a second, unshrinkable alt candidate (breakend-like)
            haps.append(np.concatenate([alt[:60], ACGT[rng.integers(0, 4, 30)], alt[90:]]))
            extra.append(-1)
            alt_ids.append(h0 + 2)
        for _r in range(reads_per_site):
            src = full_alt if rng.random() < 0.5 else contig
            start = locus - int(rng.integers(20, rl))
            read = src[start:start + rl].copy()
            q = np.clip(np.rint(rng.normal(30, 8, rl)), 0, 41).astype(np.uint8)
            err = rng.random(rl) < np.power(10.0, -q / 10.0)
            read[err] = ACGT[rng.integers(0, 4, int(err.sum()))]
            if rng.random() < 0.05:
                q[:] = 0
            lo = int(rng.integers(0, 10))
            hi = rl - int(rng.integers(0, 10))
            region_read.append(len(reads))
            reads.append(read[lo:hi])
            quals.append(q[lo:hi])
            region_haps += [h0] + alt_ids
            region_hap_off.append(len(region_haps))
            region_n_ref.append(1)
    hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    return dict(haps=haps, extra=np.array(extra, np.int32), reads=reads, quals=quals,
                hap_bytes=np.concatenate(haps), hap_off=hap_off, read_bases=np.concatenate(reads),
                read_quals=np.concatenate(quals), read_off=read_off,
                region_read=np.array(region_read, np.int32), region_haps=np.array(region_haps, np.int32),
                region_hap_off=np.array(region_hap_off, np.int32),
                region_n_ref=np.array(region_n_ref, np.int32))


@pytest.mark.parametrize("multi_alt", [False, True])
@pytest.mark.parametrize("flags", [P.HMM_DEFAULT, 0])
def test_allele_support_regions(ctx, oracle, multi_alt, flags):
    rng = np.random.default_rng(17 + int(multi_alt))
    R = _make_regions(rng, 12, 10, multi_alt)
    gap = P.GapParams()
    got = ctx.allele_support_batch(R["region_read"], R["region_hap_off"], R["region_haps"],
                                   R["region_n_ref"], R["hap_bytes"], R["hap_off"], R["read_bases"],
                                   R["read_quals"], R["read_off"], gap, shrink_extra=R["extra"],
                                   flags=flags)
    n_inf = 0
    for r in range(len(R["region_read"])):
        ids = R["region_haps"][R["region_hap_off"][r]:R["region_hap_off"][r + 1]]
        rd = R["region_read"][r]
        want = oracle.allele_support_region(
            [R["haps"][ids[0]]], [R["haps"][h] for h in ids[1:]], R["reads"][rd], R["quals"][rd],
            gap.as_ln_array(), shrink_extra=[int(R["extra"][h]) for h in ids], flags=flags)
        for key in ("prob_ref", "prob_alt", "raw_ref", "raw_alt"):
            g, w = got[key][r], want[key]
            assert (np.isneginf(g) and np.isneginf(w)) or abs(g - w) <= 1e-6, (r, key, g, w)
        assert got["ref_dist"][r] == want["ref_dist"] and got["alt_dist"][r] == want["alt_dist"], r
        assert got["n_indel_ops"][r] == len(want["indel_ops"]), r
        assert np.array_equal(got["indel_ops"][r][:len(want["indel_ops"])], want["indel_ops"]), r
        n_inf += want["informative"]
    assert n_inf > 20  # most reads must discriminate the alleles for the test to mean anything


@pytest.mark.parametrize("rl", [300, 620])
def test_allele_support_long_read_windows(ctx, oracle, rl):
    """Read windows of 280-620 rows in the exact mode (Myers::Long hits, then the banded pair-HMM
    through the anti-diagonal kernel, near ties through its literal instantiation): same values,
    distances and indel operations as the oracle."""
    rng = np.random.default_rng(4100 + rl)
    R = _make_regions(rng, 4, 6, multi_alt=(rl == 300), rl=rl)
    gap = P.GapParams()
    got = ctx.allele_support_batch(R["region_read"], R["region_hap_off"], R["region_haps"],
                                   R["region_n_ref"], R["hap_bytes"], R["hap_off"], R["read_bases"],
                                   R["read_quals"], R["read_off"], gap, shrink_extra=R["extra"])
    n_inf = 0
    for r in range(len(R["region_read"])):
        ids = R["region_haps"][R["region_hap_off"][r]:R["region_hap_off"][r + 1]]
        rd = R["region_read"][r]
        want = oracle.allele_support_region(
            [R["haps"][ids[0]]], [R["haps"][h] for h in ids[1:]], R["reads"][rd], R["quals"][rd],
            gap.as_ln_array(), shrink_extra=[int(R["extra"][h]) for h in ids], flags=P.HMM_DEFAULT)
        for key in ("prob_ref", "prob_alt", "raw_ref", "raw_alt"):
            g, w = got[key][r], want[key]
            assert (np.isneginf(g) and np.isneginf(w)) or abs(g - w) <= 1e-6, (r, key, g, w)
        assert got["ref_dist"][r] == want["ref_dist"] and got["alt_dist"][r] == want["alt_dist"], r
        assert got["n_indel_ops"][r] == len(want["indel_ops"]), r
        n_inf += want["informative"]
    assert n_inf > 10


def test_pathhmm_fast_mode(ctx, oracle):
    """--pairhmm-mode fast (PathHMMRealigner, realignment/mod.rs:499-576): probability of the best
    Myers alignment path, all three gap settings (extension and re-open terms included)."""
    for gi, gap in enumerate([P.GapParams(), P.GapParams(0.02, 0.03, 0.3, 0.3), P.GapParams(1e-3, 2e-3, 0.1, 0.2)]):
        w = synth.make_ragged_pairs(150, seed=900 + gi, max_ly=200, max_lx=400)
        got, dist = ctx.pathhmm_batch(w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                                      w["read_off"], gap)
        for k in range(150):
            hap = w["hap_bytes"][w["hap_off"][k]:w["hap_off"][k + 1]]
            read = w["read_bases"][w["read_off"][k]:w["read_off"][k + 1]]
            qual = w["read_quals"][w["read_off"][k]:w["read_off"][k + 1]]
            want = oracle.pathhmm_prob(hap, read, qual, gap.as_ln_array())
            if want is None:
                assert dist[k] == -1 and np.isnan(got[k]), k
                continue
            assert dist[k] == oracle.best_hit(hap, read)["dist"], k
            assert (np.isneginf(got[k]) and np.isneginf(want)) or abs(got[k] - want) <= 1e-6, (gi, k, got[k], want)


def test_allele_support_fast_mode(ctx, oracle):
    rng = np.random.default_rng(23)
    R = _make_regions(rng, 10, 10, True)
    gap = P.GapParams()
    got = ctx.allele_support_batch(R["region_read"], R["region_hap_off"], R["region_haps"],
                                   R["region_n_ref"], R["hap_bytes"], R["hap_off"], R["read_bases"],
                                   R["read_quals"], R["read_off"], gap, shrink_extra=R["extra"],
                                   flags=P.HMM_DEFAULT | P.REALIGN_FAST)
    for r in range(len(R["region_read"])):
        ids = R["region_haps"][R["region_hap_off"][r]:R["region_hap_off"][r + 1]]
        rd = R["region_read"][r]
        want = oracle.allele_support_region(
            [R["haps"][ids[0]]], [R["haps"][h] for h in ids[1:]], R["reads"][rd], R["quals"][rd],
            gap.as_ln_array(), shrink_extra=[int(R["extra"][h]) for h in ids], fast_mode=True)
        for key in ("prob_ref", "prob_alt", "raw_ref", "raw_alt"):
            g, w = got[key][r], want[key]
            assert (np.isneginf(g) and np.isneginf(w)) or abs(g - w) <= 1e-6, (r, key, g, w)
        assert got["ref_dist"][r] == want["ref_dist"] and got["alt_dist"][r] == want["alt_dist"], r
