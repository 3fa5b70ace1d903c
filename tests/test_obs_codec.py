"""Observation wire format (preprocessing/mod.rs:452-598, utils/mod.rs:381-407).

Golden vectors: tests/golden/obs_records.json holds 21 records WRITTEN BY THE REFERENCE ITSELF
(the candidate VCFs of its testcases still carry the INFO fields of an upstream `varlociraptor
preprocess` run; extracted by tests/golden/make_obs_golden.py).  They pin the oracle restatement
(oracle/obs_codec.py) and, through it, the library's codec.  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def records():
    return json.load(open(os.path.join(HERE, "golden", "obs_records.json")))


def test_oracle_decodes_reference_records_exactly(records):
    """Every field is consumed to the byte, all fields agree on the observation count, and the
    oracle's ENCODER reproduces the reference's integers (MiniLogProb f16/f32 choice included)."""
    from oracle import obs_codec as oc
    assert len(records) >= 20
    n_f16 = n_f32 = 0
    for r in records:
        cols, cat = oc.decode_record(r["fields"])
        assert len(cat) == r["fields"]["PROB_MAPPING"][0]
        for c in ("prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele", "prob_sample_alt",
                  "prob_double_overlap", "prob_hit_base"):
            assert (cols[c] <= 1e-6).all(), (r["name"], c)  # ln probabilities
        enc = oc.encode_record(cols, cat)
        for tag, want in r["fields"].items():
            assert enc[tag] == want, (r["name"], tag)
        kinds = [oc.minilogprob(p)[0] for p in cols["prob_alt"]]
        n_f16 += kinds.count(0)
        n_f32 += kinds.count(1)
    assert n_f16 > 100 and n_f32 > 100  # both MiniLogProb variants occur in the fixtures


def test_library_codec_matches_oracle_on_reference_records(records):
    from libprosic_b200 import obs_codec as lc
    from oracle import obs_codec as oc
    for r in records:
        want_cols, want_cat = oc.decode_record(r["fields"])
        cols, cat = lc.decode_record(r["fields"])
        assert np.array_equal(cat, want_cat), r["name"]
        for c in want_cols:
            assert np.array_equal(cols[c], want_cols[c], equal_nan=True) or \
                np.allclose(cols[c], want_cols[c], rtol=1e-15, atol=0, equal_nan=True), (r["name"], c)
        enc = lc.encode_record(cols, cat)
        for tag, want in r["fields"].items():
            assert list(enc[tag]) == want, (r["name"], tag)  # byte-identical to the reference's record
        if "INDEL_OPERATIONS" not in r["fields"]:  # format 7 -> IndelOperations::None
            assert ((cat >> 8) & 3 == 2).all()


def test_minilogprob_quantisation_and_round_trip():
    """MiniLogProb::new over a wide range (f16 subnormals, overflow, -inf, the -10 boundary):
    library encoder == oracle encoder, decode(encode(x)) == quantised x, empty records."""
    from libprosic_b200 import obs_codec as lc
    from oracle import obs_codec as oc
    rng = np.random.default_rng(12)
    n = 4000
    vals = np.concatenate([-np.exp(rng.uniform(-12, 12, n)), -rng.uniform(9.5, 10.5, 200),
                           [0.0, -0.0, -10.0, -np.inf, -65504.0, -65520.0, -70000.0, -1e-8, -2.0 ** -25]])
    n = len(vals)
    cols = {c: rng.permutation(vals) for c in ("prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele",
                                               "prob_sample_alt", "prob_double_overlap", "prob_hit_base")}
    cat = (rng.integers(0, 4, n) | (rng.integers(0, 9, n) << 2) | (rng.integers(0, 2, n) << 6) |
           (rng.integers(0, 2, n) << 7) | (rng.integers(0, 3, n) << 8) | (rng.integers(0, 2, n) << 10)).astype(np.uint16)
    enc_l, enc_o = lc.encode_record(cols, cat), oc.encode_record(cols, cat)
    for tag in lc.FIELDS:
        assert list(enc_l[tag]) == list(enc_o[tag]), tag
    dec, dcat = lc.decode_record(enc_l)
    assert np.array_equal(dcat, cat)
    for c in cols:
        want = np.array([oc.minilogprob(p)[1] for p in cols[c]])
        assert np.array_equal(dec[c], want), c
    assert np.allclose(dec["prob_mismapping"], oc.ln_one_minus_exp(dec["prob_mapping"]), rtol=1e-14, equal_nan=True)
    # empty pileup
    e = lc.encode_record({c: np.zeros(0) for c in cols}, np.zeros(0, np.uint16))
    assert list(e["SOFTCLIPPED"]) == oc.encode_bits([]) and len(lc.decode_record(e)[1]) == 0


def test_malformed_records_are_rejected(records):
    from libprosic_b200 import VlrError
    from libprosic_b200 import obs_codec as lc
    f = dict(records[0]["fields"])
    f["STRAND"] = f["STRAND"][:-3]
    with pytest.raises(VlrError):
        lc.decode_record(f)
    f = dict(records[0]["fields"])
    f["PROB_ALT"] = list(f["PROB_ALT"])
    f["PROB_ALT"][4] = 7  # MiniLogProb variant 7
    with pytest.raises(VlrError):
        lc.decode_record(f)


def test_batch_encoder_reproduces_reference_records(records):
    """vlr_obs_encode_batch (host, write_observations for a batch) on the decoded reference records:
    every field of every record comes back word for word, offsets included."""
    from libprosic_b200 import obs_codec as lc
    recs = [r for r in records if "INDEL_OPERATIONS" in r["fields"]] or records
    cols_all, cat_all, off = {c: [] for c in lc.COLUMNS}, [], [0]
    for r in recs:
        cols, cat = lc.decode_record(r["fields"])
        for c in lc.COLUMNS:
            cols_all[c].append(cols[c])
        cat_all.append(cat)
        off.append(off[-1] + len(cat))
    cols_all = {c: np.concatenate(v) for c, v in cols_all.items()}
    wire = lc.encode_batch(cols_all, np.concatenate(cat_all), np.array(off))
    assert wire.n_records == len(recs)
    for k, r in enumerate(recs):
        for f, tag in enumerate(lc.FIELDS):
            got = wire.words[f][wire.word_off[f][k]:wire.word_off[f][k + 1]]
            want = r["fields"].get(tag)
            if want is None:   # format 7 record re-encoded as format 8: IndelOperations::None everywhere
                assert tag == "INDEL_OPERATIONS"
                continue
            assert list(got) == want, (r["name"], tag)


@pytest.mark.gpu
def test_device_decoder_matches_host_codec_on_reference_records(records):
    """vlr_obs_decode_batch (one kernel, a thread per record and field) against the host codec and the
    oracle on the 21 records written by the reference: every column bit for bit."""
    import libprosic_b200 as P
    from libprosic_b200 import obs_codec as lc
    from oracle import obs_codec as oc
    ctx = P.Context(0)
    for group in ([r for r in records if "INDEL_OPERATIONS" in r["fields"]],
                  [r for r in records if "INDEL_OPERATIONS" not in r["fields"]]):
        if not group:
            continue
        wire = lc.WireBatch.from_records([r["fields"] for r in group])
        cols, cat, off = lc.decode_batch(ctx, wire)
        for k, r in enumerate(group):
            want_cols, want_cat = oc.decode_record(r["fields"])
            lo, hi = off[k], off[k + 1]
            assert hi - lo == len(want_cat)
            assert np.array_equal(cat[lo:hi], want_cat), r["name"]
            for c in ("prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele", "prob_sample_alt",
                      "prob_double_overlap", "prob_hit_base"):
                assert np.array_equal(cols[c][lo:hi], want_cols[c], equal_nan=True), (r["name"], c)
            for c in ("prob_mismapping", "prob_single_overlap"):   # derived on the device (CUDA libm)
                assert np.allclose(cols[c][lo:hi], want_cols[c], rtol=1e-14, atol=0, equal_nan=True), (r["name"], c)


@pytest.mark.gpu
def test_device_decoder_random_batches_and_malformed_input():
    import libprosic_b200 as P
    from libprosic_b200 import obs_codec as lc
    ctx = P.Context(0)
    rng = np.random.default_rng(5)
    n_rec = 3000
    n_obs = rng.integers(0, 70, n_rec)
    n_obs[:5] = [0, 1, 7, 8, 9]
    off = np.concatenate([[0], np.cumsum(n_obs)])
    n = int(off[-1])
    vals = np.concatenate([-np.exp(rng.uniform(-12, 8, n)), [0.0, -10.0, -np.inf, -65504.0, -70000.0]])
    cols = {c: rng.choice(vals, n) for c in ("prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele",
                                             "prob_sample_alt", "prob_double_overlap", "prob_hit_base")}
    cat = (rng.integers(0, 4, n) | (rng.integers(0, 9, n) << 2) | (rng.integers(0, 2, n) << 6) |
           (rng.integers(0, 2, n) << 7) | (rng.integers(0, 3, n) << 8) | (rng.integers(0, 2, n) << 10)).astype(np.uint16)
    wire = lc.encode_batch(cols, cat, off)
    dcols, dcat, doff = lc.decode_batch(ctx, wire)
    assert np.array_equal(doff, off) and np.array_equal(dcat, cat)
    for k in (0, 1, 2, 17, n_rec - 1):   # against the per-record host decoder
        rec = {t: wire.words[f][wire.word_off[f][k]:wire.word_off[f][k + 1]].astype(np.int32) for f, t in enumerate(lc.FIELDS)}
        hc, hcat = lc.decode_record(rec)
        for c in hc:
            got = dcols[c][off[k]:off[k + 1]]
            assert np.array_equal(got, hc[c], equal_nan=True) or np.allclose(got, hc[c], rtol=1e-14, atol=0, equal_nan=True), (k, c)
    # malformed: a variant index of 7 inside PROB_ALT of record 2; a truncated field
    bad = lc.WireBatch([w.copy() for w in wire.words], [o.copy() for o in wire.word_off])
    bad.words[2][bad.word_off[2][2] + 4] = 7
    with pytest.raises(P.VlrError):
        lc.decode_batch(ctx, bad)
    bad = lc.WireBatch([w.copy() for w in wire.words], [o.copy() for o in wire.word_off])
    bad.word_off[0][-1] -= 1
    with pytest.raises(P.VlrError):
        lc.decode_batch(ctx, bad)
