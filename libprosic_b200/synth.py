"""Synthetic workloads of SURVEY.md section 8(d) (fixed PCG64 seeds; shapes of BASELINE.json configs).

Pure numpy input generation shared by tests/ and bench.py; nothing here computes a likelihood.
"""
import numpy as np

SEED_C2 = 20261019
SEED_C3 = 20261020
SEED_C4 = 20261021
SEED_C5 = 20261022

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _quals(rng, shape, mean=32.0, sd=5.0, lo=2, hi=41):
    return np.clip(np.rint(rng.normal(mean, sd, size=shape)), lo, hi).astype(np.uint8)


def make_pairhmm_workload(n_sites, reads_per_site, seed=SEED_C3, read_len=150, hap_len=300,
                          max_indel=30, indel_err=1e-4, qual_mean=32.0, qual_sd=5.0, qual_lo=2,
                          qual_hi=41, sub_scale=1.0):
    """Config C3 part 1: per site a ref and an alt haplotype window (alt = ref with one indel of
    length U[1,max_indel] at the centre), reads of `read_len` sampled 50/50 from ref or alt at a
    uniform offset with substitution errors at the base-quality rate and rare indel errors.
    Every read is paired with both haplotypes of its site (2 pairs per read).

    Returns dict(hap_bytes, hap_off, read_bases, read_quals, read_off, pair_hap, pair_read,
    cells) in the CSR layout of include/vlr_gpu.h.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    pad = max_indel + 8
    contig_len = hap_len + 2 * pad
    contigs = _ACGT[rng.integers(0, 4, size=(n_sites, contig_len))]
    centre = pad + hap_len // 2
    ind_len = rng.integers(1, max_indel + 1, size=n_sites)
    is_ins = rng.random(n_sites) < 0.5
    ins_seq = _ACGT[rng.integers(0, 4, size=(n_sites, max_indel))]
    haps = np.empty((n_sites, 2, hap_len), dtype=np.uint8)
    haps[:, 0, :] = contigs[:, pad:pad + hap_len]
    half = hap_len // 2
    for s in range(n_sites):
        L = int(ind_len[s])
        c = contigs[s]
        if is_ins[s]:
            alt = np.concatenate([c[pad:centre], ins_seq[s, :L], c[centre:centre + (hap_len - half - L)]])
        else:
            alt = np.concatenate([c[pad:centre], c[centre + L:centre + L + (hap_len - half)]])
        haps[s, 1, :] = alt[:hap_len]
    n_reads = n_sites * reads_per_site
    site_of = np.repeat(np.arange(n_sites), reads_per_site)
    from_alt = rng.random(n_reads) < 0.5
    offs = rng.integers(0, hap_len - read_len + 1, size=n_reads)
    idx = offs[:, None] + np.arange(read_len)[None, :]
    reads = haps[site_of, from_alt.astype(np.int64)][np.arange(n_reads)[:, None], idx]
    quals = _quals(rng, (n_reads, read_len), qual_mean, qual_sd, qual_lo, qual_hi)
    eps = np.power(10.0, -quals.astype(np.float64) / 10.0) * sub_scale
    sub = rng.random((n_reads, read_len)) < eps
    shift = rng.integers(1, 4, size=(n_reads, read_len))
    code = np.zeros(256, dtype=np.int64)
    code[_ACGT] = np.arange(4)
    reads = np.where(sub, _ACGT[(code[reads] + shift) % 4], reads).astype(np.uint8)
    # rare indel errors: delete or duplicate one base, keep the length by shifting in a random base
    n_err = rng.binomial(n_reads * read_len, indel_err)
    for _ in range(int(n_err)):
        r = int(rng.integers(0, n_reads))
        p = int(rng.integers(1, read_len - 1))
        if rng.random() < 0.5:
            reads[r, p:-1] = reads[r, p + 1:]
            reads[r, -1] = _ACGT[rng.integers(0, 4)]
        else:
            reads[r, p + 1:] = reads[r, p:-1].copy()
    hap_bytes = haps.reshape(-1)
    hap_off = np.arange(0, (2 * n_sites + 1) * hap_len, hap_len, dtype=np.int64)
    read_off = np.arange(0, (n_reads + 1) * read_len, read_len, dtype=np.int64)
    pair_read = np.repeat(np.arange(n_reads, dtype=np.int32), 2)
    pair_hap = (2 * site_of[:, None] + np.arange(2)[None, :]).reshape(-1).astype(np.int32)
    return dict(hap_bytes=np.ascontiguousarray(hap_bytes), hap_off=hap_off,
                read_bases=np.ascontiguousarray(reads.reshape(-1)),
                read_quals=np.ascontiguousarray(quals.reshape(-1)), read_off=read_off,
                pair_hap=pair_hap, pair_read=pair_read, from_alt=from_alt, site_of=site_of,
                cells=int(len(pair_hap)) * read_len * hap_len)


def make_ragged_pairs(n_pairs, seed, min_ly=1, max_ly=256, min_lx=1, max_lx=700, lower_frac=0.05,
                      related_frac=0.7, qual_lo=0, qual_hi=60):
    """Ragged (read, haplotype) pairs for edge-case parity: random lengths, some reads derived from
    their haplotype with edits, lower-case and N bases, quality 0 .. 60."""
    rng = np.random.Generator(np.random.PCG64(seed))
    haps, reads, quals = [], [], []
    alphabet = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for _ in range(n_pairs):
        lx = int(rng.integers(min_lx, max_lx + 1))
        ly = int(rng.integers(min_ly, max_ly + 1))
        hap = alphabet[rng.choice(5, size=lx, p=[0.245, 0.245, 0.245, 0.245, 0.02])]
        if rng.random() < related_frac and lx >= 2:
            start = int(rng.integers(0, max(1, lx - 1)))
            seg = hap[start:start + ly]
            read = list(seg)
            # a few edits
            for _e in range(int(rng.integers(0, 6))):
                if not read:
                    break
                p = int(rng.integers(0, len(read)))
                k = rng.random()
                if k < 0.5:
                    read[p] = int(alphabet[rng.integers(0, 4)])
                elif k < 0.75:
                    del read[p]
                else:
                    read.insert(p, int(alphabet[rng.integers(0, 4)]))
            while len(read) < ly:
                read.append(int(alphabet[rng.integers(0, 4)]))
            read = np.array(read[:ly], dtype=np.uint8)
        else:
            read = alphabet[rng.integers(0, 4, size=ly)]
        if rng.random() < lower_frac:
            hap = np.where(hap < 96, hap + 32, hap).astype(np.uint8)  # lower-case haplotype
        q = rng.integers(qual_lo, qual_hi + 1, size=ly).astype(np.uint8)
        haps.append(hap)
        reads.append(read)
        quals.append(q)
    hap_off = np.zeros(n_pairs + 1, dtype=np.int64)
    hap_off[1:] = np.cumsum([len(h) for h in haps])
    read_off = np.zeros(n_pairs + 1, dtype=np.int64)
    read_off[1:] = np.cumsum([len(r) for r in reads])
    return dict(hap_bytes=np.concatenate(haps), hap_off=hap_off, read_bases=np.concatenate(reads),
                read_quals=np.concatenate(quals), read_off=read_off)


# ------------------------------------------------------------------------------------------------
# pileups (part 2 inputs): Observation columns of include/vlr_gpu.h
# ------------------------------------------------------------------------------------------------
OBS_COLUMNS = ("prob_mapping", "prob_mismapping", "prob_alt", "prob_ref", "prob_missed_allele",
               "prob_sample_alt", "prob_double_overlap", "prob_single_overlap", "prob_hit_base")


def minilogprob(p):
    """MiniLogProb wire quantisation (reference src/utils/mod.rs:381-407): f16 if p < -10 and
    floor(f16(p)) == floor(p), else f32.  The likelihood model only ever sees these values."""
    p = np.asarray(p, dtype=np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        h = p.astype(np.float16).astype(np.float64)
        use16 = (p < -10.0) & (np.floor(h) == np.floor(p))
        return np.where(use16, h, p.astype(np.float32).astype(np.float64))


def _ln1mexp(p):
    p = np.asarray(p, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(p < -0.693, np.log1p(-np.exp(p)), np.log(-np.expm1(p)))


def pack_cat(strand, orientation, read_position, softclipped, indel_ops, paired):
    return (np.asarray(strand, np.uint16) | (np.asarray(orientation, np.uint16) << 2)
            | (np.asarray(read_position, np.uint16) << 6)
            | (np.asarray(softclipped, np.uint16) << 7) | (np.asarray(indel_ops, np.uint16) << 8)
            | (np.asarray(paired, np.uint16) << 10)).astype(np.uint16)


def make_pileups(n_sites, depths, seed=SEED_C2, kind="snv", af_sets=None, purity=None,
                 quantise=True, ragged=False, artifact_frac=0.0):
    """Synthetic pileups per SURVEY.md section 8(d), config C2 (kind="snv") / C3 part 2
    (kind="indel", tumor/normal).

    depths: observations per sample (list, one entry per sample, sample order of the scenario).
    af_sets: per sample list of (af, probability) for the site's true allele frequency; default
             {0: .9, .5: .07, 1: .03} (C2).  purity: per-sample mixing with sample 0's AF (tumor
             contaminated by normal), or None.
    ragged: depths vary per site (Poisson around the given depth, zero allowed).
    Returns (cols dict, cat, obs_off[n_sites*n_samples+1], true_af[n_sites, n_samples]).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    ns = len(depths)
    if af_sets is None:
        af_sets = [[(0.0, 0.90), (0.5, 0.07), (1.0, 0.03)]] * ns
    true_af = np.empty((n_sites, ns))
    for s in range(ns):
        vals = np.array([a for a, _ in af_sets[s]])
        pr = np.array([p for _, p in af_sets[s]])
        true_af[:, s] = vals[rng.choice(len(vals), size=n_sites, p=pr / pr.sum())]
    if ragged:
        n_obs = np.stack([rng.poisson(d, size=n_sites) for d in depths], axis=1)
        n_obs = np.minimum(n_obs, 250)
        n_obs[rng.random(n_sites) < 0.03] = 0
    else:
        n_obs = np.tile(np.asarray(depths, dtype=np.int64), (n_sites, 1))
    obs_off = np.zeros(n_sites * ns + 1, dtype=np.int64)
    obs_off[1:] = np.cumsum(n_obs.reshape(-1))
    n = int(obs_off[-1])
    site_of = np.repeat(np.arange(n_sites * ns) // ns, n_obs.reshape(-1))
    samp_of = np.repeat(np.arange(n_sites * ns) % ns, n_obs.reshape(-1))
    eff_af = true_af[site_of, samp_of]
    if purity is not None:
        pur = np.asarray(purity)[samp_of]
        eff_af = pur * eff_af + (1.0 - pur) * true_af[site_of, 0]
    is_alt = rng.random(n) < eff_af
    ln_conf = np.log(0.3333)
    if kind == "snv":
        q = np.clip(np.rint(rng.normal(32, 5, n)), 2, 41)
        ln_eps = q * (-np.log(10.0) / 10.0)
        sup, oth = _ln1mexp(ln_eps), ln_eps + ln_conf      # prob_read_base, bases.rs:13-21
        indel = np.full(n, 2)                               # IndelOperations::None
    else:
        # realignment-like: normalised pair of allele probabilities, some uninformative reads
        d = np.power(10.0, -rng.uniform(0.5, 30.0, n))
        sup, oth = np.log1p(-d), np.log(d)
        uninf = rng.random(n) < 0.08
        sup = np.where(uninf, np.log(0.5), sup)
        oth = np.where(uninf, np.log(0.5), oth)
        indel = np.where(is_alt & ~uninf, np.where(rng.random(n) < 0.9, 0, 1), 2)
    p_alt = np.where(is_alt, sup, oth)
    p_ref = np.where(is_alt, oth, sup)
    mapq = np.where(rng.random(n) < 0.92, 60.0, rng.integers(1, 60, n).astype(np.float64))
    prob_mapping = np.log1p(-np.power(10.0, -mapq / 10.0))
    strand = rng.integers(0, 2, n)
    both = rng.random(n) < 0.05
    strand = np.where(both, 2, strand)
    orient = rng.integers(0, 2, n)
    orient = np.where(rng.random(n) < 0.03, 8, orient)      # a few without orientation info
    softclip = rng.random(n) < 0.02
    # ReadPosition::Major = the most common read position of the pileup (observation.rs:314-326);
    # synthetic: every pileup has one over-represented position
    major_pos = rng.integers(0, 150, n_sites * ns)[np.repeat(np.arange(n_sites * ns), n_obs.reshape(-1))]
    readpos = np.where(rng.random(n) < 0.12, major_pos, rng.integers(0, 150, n))
    major = (readpos == major_pos) & (kind == "snv")
    if artifact_frac > 0:
        # strand artifact: at some sites every alt-supporting read is on the forward strand
        art_site = rng.random(n_sites) < artifact_frac
        force = art_site[site_of] & is_alt
        strand = np.where(force, 0, strand)
    prob_double_overlap = np.where(strand == 2, 0.0, -np.inf)
    with np.errstate(divide="ignore"):
        missed = np.logaddexp(p_ref, p_alt) - np.log(2.0)
    cols = dict(prob_mapping=prob_mapping, prob_alt=p_alt, prob_ref=p_ref, prob_missed_allele=missed,
                prob_sample_alt=np.zeros(n) if kind == "snv" else np.log(rng.uniform(0.7, 1.0, n)),
                prob_double_overlap=prob_double_overlap, prob_hit_base=np.full(n, -np.log(150.0)))
    if quantise:
        cols = {k: minilogprob(v) for k, v in cols.items()}
    cols["prob_mismapping"] = _ln1mexp(cols["prob_mapping"])          # prob_mapping_mismapping()
    cols["prob_single_overlap"] = _ln1mexp(cols["prob_double_overlap"])  # prob_overlap()
    cat = pack_cat(strand, orient, np.where(major, 0, 1), softclip, indel, np.ones(n, dtype=int))
    return cols, cat, obs_off, true_af


def make_long_read_pairs(n_pairs, seed=SEED_C4, read_len=15000, hap_len=20000, sub=0.08, ins=0.02,
                         dele=0.02):
    """Config C4: long reads (Q ~ clip(N(15,4),3,30), 8 % substitutions, 2 % insertions, 2 %
    deletions) against longer haplotype windows; one pair per read."""
    rng = np.random.Generator(np.random.PCG64(seed))
    haps = _ACGT[rng.integers(0, 4, size=(n_pairs, hap_len))]
    reads, quals = [], []
    for k in range(n_pairs):
        s = int(rng.integers(0, hap_len - read_len + 1))
        r = haps[k, s:s + read_len].copy()
        m = rng.random(read_len) < sub
        r[m] = _ACGT[rng.integers(0, 4, int(m.sum()))]
        r = r[rng.random(read_len) > dele]
        pos = np.sort(rng.integers(0, len(r), int(ins * len(r))))
        r = np.insert(r, pos, _ACGT[rng.integers(0, 4, len(pos))])[:read_len]
        reads.append(r)
        quals.append(_quals(rng, len(r), 15.0, 4.0, 3, 30))
    read_off = np.zeros(n_pairs + 1, dtype=np.int64)
    read_off[1:] = np.cumsum([len(r) for r in reads])
    hap_off = np.arange(0, (n_pairs + 1) * hap_len, hap_len, dtype=np.int64)
    cells = int(sum(len(r) for r in reads)) * hap_len
    return dict(hap_bytes=np.ascontiguousarray(haps.reshape(-1)), hap_off=hap_off,
                read_bases=np.concatenate(reads), read_quals=np.concatenate(quals),
                read_off=read_off, cells=cells)
