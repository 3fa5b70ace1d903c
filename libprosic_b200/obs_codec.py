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
