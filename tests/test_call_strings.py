"""OBS / SOBS generalized-CIGAR strings and bias flag characters of the call record (SURVEY 8f-3;
reference src/calling/variants/mod.rs:147-260, 335-392, src/utils/mod.rs:101-146) against
oracle/call_records.py.  The reference's order among items with equal counts is a HashMap's, so the
comparison is per (class, count) multiset; the library's own order is deterministic (code order)."""
import ctypes as C
import re

import numpy as np
import pytest

from libprosic_b200 import _ffi, synth
from oracle import call_records as ocr


def _parse(s, width):
    if s == ".":
        return []
    out = []
    for m in re.finditer(r"(\d+)(\D.{%d})" % (width - 1), s):
        out.append((int(m.group(1)), m.group(2)))
    assert "".join("%d%s" % kv for kv in out) == s
    return out


def _klass(item):
    return 2 if item.startswith("N") else (1 if item.startswith("E") else 0)


def _same(got, want):
    assert sorted(got) == sorted(want), (got, want)
    # ordering rule: class ascending, then count descending
    keys = [(_klass(it), -n) for n, it in got]
    assert keys == sorted(keys), got


def test_cigar_formatting_on_the_host():
    """vlr_obs_cigar needs no GPU: codes in, string out."""
    lib = _ffi.load()
    # codes: V upper paired + > ^ . * (x3), n lower single - ! * $ # (x2), E (x1)
    def code(letter, lower, paired, strand, orient, readpos, soft, indel):
        return letter | lower << 3 | paired << 4 | strand << 5 | orient << 7 | readpos << 9 | soft << 10 | indel << 11
    codes = np.array([code(5, 0, 1, 2, 0, 0, 0, 0)] * 3 + [code(0, 1, 0, 1, 3, 1, 1, 1)] * 2 + [code(1, 0, 1, 0, 2, 1, 0, 2)],
                     np.uint16)
    buf = C.create_string_buffer(256)
    n = lib.vlr_obs_cigar(codes.ctypes.data_as(C.c_void_p), len(codes), 0, buf, 256)
    assert n > 0 and buf.value.decode() == "3Vp+>^.*2ns-!*$#1Ep***.."
    n = lib.vlr_obs_cigar(codes.ctypes.data_as(C.c_void_p), len(codes), 1, buf, 256)
    assert buf.value.decode() == "3V2n1E"
    assert lib.vlr_obs_cigar(None, 0, 0, buf, 256) == 1 and buf.value == b"."
    assert lib.vlr_obs_cigar(codes.ctypes.data_as(C.c_void_p), len(codes), 0, buf, 8) < 0  # capacity


def test_bias_chars_on_the_host():
    lib = _ffi.load()
    for b in (dict(strand=0, orientation=0, read_position=0, softclip=0, divindel=0),
              dict(strand=1, orientation=2, read_position=1, softclip=0, divindel=1),
              dict(strand=2, orientation=1, read_position=0, softclip=1, divindel=0)):
        bb = _ffi.Biases(b["strand"], b["orientation"], b["read_position"], b["softclip"], b["divindel"], 0, 0.0)
        out = C.create_string_buffer(5)
        assert lib.vlr_bias_chars(C.byref(bb), out) == 0
        assert out.raw.decode() == ocr.bias_chars(b)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["snv", "indel"])
def test_obs_and_sobs_strings(ctx, kind):
    cols, cat, off, _ = synth.make_pileups(60, [30, 12], seed=5, kind=kind, ragged=True, artifact_frac=0.2)
    got = ctx.call_obs_strings(cols, cat, off)
    assert len(got) == len(off) - 1
    n_nonempty = 0
    for k, (obs, sobs) in enumerate(got):
        lo, hi = int(off[k]), int(off[k + 1])
        items = ocr.obs_items(cols["prob_alt"][lo:hi], cols["prob_ref"][lo:hi], cols["prob_mapping"][lo:hi], cat[lo:hi])
        _same(_parse(obs, 7), ocr.generalized_cigar([a for a, _ in items]))
        _same(_parse(sobs, 1), ocr.generalized_cigar([b for _, b in items]))
        n_nonempty += hi > lo
    assert n_nonempty > 50 and any(o == "." for o, _ in got)
