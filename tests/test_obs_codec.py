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
