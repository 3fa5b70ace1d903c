"""Host mirror of the observation wire format entry points (vlr_obs_decode / vlr_obs_encode):
the INFO fields `varlociraptor preprocess` writes and `call` reads
(src/calling/variants/preprocessing/mod.rs:452-598).  No GPU needed."""
import ctypes as C

import numpy as np

from . import _ffi

FIELDS = ["PROB_MAPPING", "PROB_REF", "PROB_ALT", "PROB_MISSED_ALLELE", "PROB_SAMPLE_ALT",
          "PROB_DOUBLE_OVERLAP", "PROB_HIT_BASE", "STRAND", "READ_ORIENTATION", "READ_POSITION",
          "SOFTCLIPPED", "INDEL_OPERATIONS", "PAIRED"]
COLUMNS = ["prob_mapping", "prob_mismapping", "prob_alt", "prob_ref", "prob_missed_allele",
           "prob_sample_alt", "prob_double_overlap", "prob_single_overlap", "prob_hit_base"]


def decode_record(fields):
    """fields: dict tag -> sequence of INFO ints.  Returns (cols dict of fp64 arrays, cat uint16)."""
    lib = _ffi.load()
    arrs = [np.ascontiguousarray(fields[t], dtype=np.int32) if fields.get(t) is not None and len(fields[t])
            else None for t in FIELDS]
    ptrs = (C.c_void_p * 13)(*[a.ctypes.data if a is not None else None for a in arrs])
    lens = np.array([len(a) if a is not None else 0 for a in arrs], np.int32)
    n = C.c_int64()
    rc = lib.vlr_obs_decode(ptrs, lens.ctypes.data, 0, C.byref(n), None, None)
    if rc != 0:
        raise _ffi.VlrError(rc, "vlr_obs_decode: malformed observation record")
    cols = {c: np.empty(n.value, np.float64) for c in COLUMNS}
    cat = np.empty(n.value, np.uint16)
    cptr = (C.c_void_p * 9)(*[cols[c].ctypes.data for c in COLUMNS])
    rc = lib.vlr_obs_decode(ptrs, lens.ctypes.data, n.value, C.byref(n), cptr, cat.ctypes.data)
    if rc != 0:
        raise _ffi.VlrError(rc, "vlr_obs_decode: malformed observation record")
    return cols, cat


def encode_record(cols, cat):
    """Returns dict tag -> int32 array (format 8)."""
    lib = _ffi.load()
    cat = np.ascontiguousarray(cat, dtype=np.uint16)
    n = len(cat)
    keep = [np.ascontiguousarray(cols[c], dtype=np.float64) if c in cols else None for c in COLUMNS]
    cptr = (C.c_void_p * 9)(*[a.ctypes.data if a is not None else None for a in keep])
    cap = int(lib.vlr_obs_encoded_bound(n))
    outs = [np.empty(cap, np.int32) for _ in FIELDS]
    optr = (C.c_void_p * 13)(*[o.ctypes.data for o in outs])
    lens = np.zeros(13, np.int32)
    rc = lib.vlr_obs_encode(n, cptr, cat.ctypes.data, optr, cap, lens.ctypes.data)
    if rc != 0:
        raise _ffi.VlrError(rc, "vlr_obs_encode failed")
    return {t: outs[k][:lens[k]].copy() for k, t in enumerate(FIELDS)}


class WireBatch:
    """A batch of stored observation records (vlr_obs_wire): per field one uint16 array with the words
    of all records back to back and one int64 offset array.  `words[f] is None`: field absent
    (INDEL_OPERATIONS of format 7)."""

    def __init__(self, words, word_off, pin=False):
        self.words = [None if w is None else np.ascontiguousarray(w, dtype=np.uint16) for w in words]
        self.word_off = [None if o is None else np.ascontiguousarray(o, dtype=np.int64) for o in word_off]
        self.n_records = len(next(o for o in self.word_off if o is not None)) - 1
        self._pinned = None
        if pin:  # page-locked copies (bench: asynchronous H2D)
            import torch
            self._pinned = [None if w is None else torch.from_numpy(w.view(np.int16)).pin_memory() for w in self.words]
            self.words = [None if t is None else t.numpy().view(np.uint16) for t in self._pinned]
            self._pinned_off = [None if o is None else torch.from_numpy(o).pin_memory() for o in self.word_off]
            self.word_off = [None if t is None else t.numpy() for t in self._pinned_off]

    @property
    def nbytes(self):
        return sum(w.nbytes for w in self.words if w is not None) + sum(o.nbytes for o in self.word_off if o is not None)

    def c_struct(self):
        w = _ffi.ObsWire()
        for f in range(13):
            w.words[f] = None if self.words[f] is None else self.words[f].ctypes.data
            w.word_off[f] = None if self.word_off[f] is None else self.word_off[f].ctypes.data
        return w

    @staticmethod
    def from_records(records):
        """records: list of dict tag -> INFO ints (a missing / empty INDEL_OPERATIONS everywhere = format 7)."""
        words, offs = [], []
        for t in FIELDS:
            vals = [np.asarray(r.get(t) if r.get(t) is not None else [], dtype=np.int64) for r in records]
            if t == "INDEL_OPERATIONS" and all(len(v) == 0 for v in vals):
                words.append(None)
                offs.append(None)
                continue
            offs.append(np.concatenate([[0], np.cumsum([len(v) for v in vals])]).astype(np.int64))
            words.append((np.concatenate(vals) if vals else np.zeros(0, np.int64)).astype(np.uint16))
        return WireBatch(words, offs)


def encode_batch(cols, cat, obs_off, pin=False):
    """write_observations for a batch: struct-of-arrays columns -> WireBatch (format 8)."""
    lib = _ffi.load()
    obs_off = np.ascontiguousarray(obs_off, dtype=np.int64)
    cat = np.ascontiguousarray(cat, dtype=np.uint16)
    n_rec = len(obs_off) - 1
    keep = [np.ascontiguousarray(cols[c], dtype=np.float64) if c in cols else None for c in COLUMNS]
    cptr = (C.c_void_p * 9)(*[a.ctypes.data if a is not None else None for a in keep])
    n = np.diff(obs_off)
    cap = int(np.sum((8 + n * 8 + 1) // 2 + 1)) + 16
    words = [np.zeros(cap, np.uint16) for _ in FIELDS]
    offs = [np.zeros(n_rec + 1, np.int64) for _ in FIELDS]
    wptr = (C.c_void_p * 13)(*[w.ctypes.data for w in words])
    optr = (C.c_void_p * 13)(*[o.ctypes.data for o in offs])
    rc = lib.vlr_obs_encode_batch(n_rec, cptr, cat.ctypes.data, obs_off.ctypes.data, wptr, optr, cap)
    if rc != 0:
        raise _ffi.VlrError(rc, "vlr_obs_encode_batch failed")
    return WireBatch([w[:o[-1]].copy() for w, o in zip(words, offs)], offs, pin=pin)


def decode_batch(ctx, wire):
    """read_observations on the device (vlr_obs_decode_batch): returns (cols dict, cat, obs_off)."""
    lib = _ffi.load()
    ws = wire.c_struct()
    off = np.zeros(wire.n_records + 1, np.int64)
    ctx._check(lib.vlr_obs_decode_batch(ctx.h, wire.n_records, C.byref(ws), 0, off.ctypes.data, None, None))
    n = int(off[-1])
    cols = {c: np.empty(n, np.float64) for c in COLUMNS}
    cat = np.empty(n, np.uint16)
    cptr = (C.c_void_p * 9)(*[cols[c].ctypes.data for c in COLUMNS])
    ctx._check(lib.vlr_obs_decode_batch(ctx.h, wire.n_records, C.byref(ws), n, off.ctypes.data, cptr, cat.ctypes.data))
    return cols, cat, off
