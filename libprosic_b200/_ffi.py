"""ctypes loader for libvlr_b200.so (the C ABI declared in include/vlr_gpu.h).

The library is the product; this module only binds it.  There is no fallback: if the shared
library is missing or the device is not a CUDA GPU, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvlr_b200.so")

VLR_OK = 0
HMM_SUM3_APPROX = 1   # two-way ln_sum3_exp_approx
HMM_PATH_CLOSE = 2
REALIGN_FAST = 4  # vlr_allele_support_batch: --pairhmm-mode fast (PathHMMRealigner)
HMM_SUM3_APPROX3 = 8  # three-way ln_sum3_exp_approx (default)
HMM_LITERAL = 16      # vlr_pairhmm_batch: literal log-space evaluation, deterministic libm
HMM_DEFAULT = HMM_SUM3_APPROX3
MAX_INDEL_OPS = 32

NODE_SET, NODE_RANGE, NODE_FALSE, NODE_SKIP = 0, 1, 2, 3


class VlrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libvlr_b200 error %d: %s" % (code, msg))
        self.code = code


class GapParams(C.Structure):
    """realignment/pairhmm.rs:73-79 (ln probabilities)."""
    _fields_ = [("prob_gap_x", C.c_double), ("prob_gap_y", C.c_double),
                ("prob_gap_x_extend", C.c_double), ("prob_gap_y_extend", C.c_double)]


_DP = C.POINTER(C.c_double)
_U16P = C.POINTER(C.c_uint16)


class ObsColumns(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "prob_mapping", "prob_mismapping", "prob_alt", "prob_ref", "prob_missed_allele",
        "prob_sample_alt", "prob_double_overlap", "prob_single_overlap", "prob_hit_base", "cat")]


class Biases(C.Structure):
    _fields_ = [("strand", C.c_int32), ("orientation", C.c_int32), ("read_position", C.c_int32),
                ("softclip", C.c_int32), ("divindel", C.c_int32), ("_pad", C.c_int32),
                ("min_other_rate", C.c_double)]


class LhItem(C.Structure):
    _fields_ = [("site", C.c_int32), ("sample", C.c_int32), ("bias", C.c_int32),
                ("contaminated", C.c_int32), ("af", C.c_double), ("af_secondary", C.c_double),
                ("purity", C.c_double), ("other_rate", C.c_double)]


class Node(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sample", C.c_int32), ("child_off", C.c_int32),
                ("n_children", C.c_int32), ("vaf_off", C.c_int32), ("n_vafs", C.c_int32)]


class Event(C.Structure):
    _fields_ = [("root_off", C.c_int32), ("n_roots", C.c_int32), ("bias_off", C.c_int32),
                ("n_biases", C.c_int32)]


class ObsColumnsF32(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("prob_mapping", "prob_alt", "prob_ref", "prob_missed_allele",
                                          "prob_sample_alt", "prob_double_overlap", "prob_hit_base", "cat")]


class PriorProg(C.Structure):
    """vlr_prior_prog (include/vlr_gpu.h)."""
    _fields_ = [("kind", C.c_int32 * 8), ("ploidy", C.c_int32 * 8), ("som_kind", C.c_int32 * 8),
                ("parent", C.c_int32 * 8), ("ln_rate", C.c_double * 8), ("som_zero", C.c_double * 8),
                ("ln_genome_size", C.c_double), ("uni_off", C.c_int32 * 9), ("_pad", C.c_int32),
                ("uni_spec", C.c_void_p), ("germ_ln", C.c_void_p), ("n_germ", C.c_int64)]


class Model(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_events", C.c_int32), ("n_nodes", C.c_int32),
                ("n_biases", C.c_int32), ("resolution", C.c_void_p), ("contaminated", C.c_void_p),
                ("by", C.c_void_p), ("purity", C.c_void_p), ("nodes", C.c_void_p),
                ("children", C.c_void_p), ("vafs", C.c_void_p), ("roots", C.c_void_p),
                ("events", C.c_void_p), ("biases", C.c_void_p), ("prior_ln", C.c_void_p),
                ("n_prior", C.c_int64), ("prior_prog", C.c_void_p)]


EXPORTS = [
    "vlr_version", "vlr_build_stamp", "vlr_ctx_create", "vlr_ctx_destroy", "vlr_last_error", "vlr_ctx_set_stream",
    "vlr_ctx_synchronize", "vlr_ctx_launch_count", "vlr_ref_math", "vlr_pairhmm_batch",
    "vlr_pairhmm_workspace_bytes", "vlr_pairhmm_batch_dev", "vlr_measure_fp64_peak",
    "vlr_besthit_batch", "vlr_pathhmm_batch", "vlr_allele_support_batch", "vlr_likelihood_batch", "vlr_call_records",
    "vlr_obs_codes", "vlr_obs_cigar", "vlr_bias_chars", "vlr_model_leaves", "vlr_obs_decode", "vlr_obs_encode", "vlr_obs_encoded_bound", "vlr_obs_decode_batch", "vlr_obs_encode_batch", "vlr_posterior_batch_wire", "vlr_posterior_batch", "vlr_posterior_batch_f32", "vlr_posterior_batch_dev",
]

class ObsWire(C.Structure):
    """vlr_obs_wire: per field the u16 words of all records and their offsets."""
    _fields_ = [("words", C.c_void_p * 13), ("word_off", C.c_void_p * 13)]


_lib = None


def load():
    """Load the shared library (building it first when nvcc is available and sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    _build.build()
    if not os.path.exists(SO_PATH):
        raise ImportError("libvlr_b200.so is missing: run `python -m libprosic_b200.build`")
    L = C.CDLL(SO_PATH)
    vp, i32, i64, u32, d = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_double
    L.vlr_version.restype = C.c_int
    L.vlr_build_stamp.restype = C.c_char_p
    # a binary that was not built from the sources next to it is refused (a GPU box has no nvcc: the
    # library travels prebuilt, and a stale one would otherwise pass silently)
    have, want = L.vlr_build_stamp().decode(), _build.source_stamp()
    if have != want:
        raise ImportError("libvlr_b200.so is stale: built from sources %s, tree has %s -- run "
                          "`python -m libprosic_b200.build -f`" % (have[:12], want[:12]))
    L.vlr_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.vlr_ctx_destroy.argtypes = [vp]
    L.vlr_ctx_destroy.restype = None
    L.vlr_last_error.argtypes = [vp]
    L.vlr_last_error.restype = C.c_char_p
    L.vlr_ctx_set_stream.argtypes = [vp, vp]
    L.vlr_ctx_synchronize.argtypes = [vp]
    L.vlr_ctx_launch_count.argtypes = [vp]
    L.vlr_ctx_launch_count.restype = i64
    L.vlr_ref_math.argtypes = [vp, C.c_int, i64, vp, vp]
    L.vlr_pairhmm_batch.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp,
                                    C.POINTER(GapParams), u32, vp]
    L.vlr_pairhmm_workspace_bytes.argtypes = [i64, i64, i64]
    L.vlr_pairhmm_workspace_bytes.restype = C.c_size_t
    L.vlr_pairhmm_batch_dev.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp,
                                        C.POINTER(GapParams), u32, vp, C.c_size_t, vp]
    L.vlr_measure_fp64_peak.argtypes = [vp, _DP]
    L.vlr_besthit_batch.argtypes = [vp, i64, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.vlr_pathhmm_batch.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    L.vlr_allele_support_batch.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64,
                                           C.POINTER(GapParams), u32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.vlr_likelihood_batch.argtypes = [vp, i64, i32, C.POINTER(ObsColumns), vp, vp, i32, vp, i64, vp]
    L.vlr_call_records.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.vlr_obs_codes.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.vlr_obs_cigar.argtypes = [vp, i64, C.c_int, vp, i64]
    L.vlr_obs_cigar.restype = i64
    L.vlr_bias_chars.argtypes = [vp, vp]
    L.vlr_obs_decode.argtypes = [vp, vp, i64, vp, vp, vp]
    L.vlr_obs_encode.argtypes = [i64, vp, vp, vp, i64, vp]
    L.vlr_obs_encoded_bound.argtypes = [i64]
    L.vlr_obs_encoded_bound.restype = i64
    L.vlr_obs_decode_batch.argtypes = [vp, i64, C.POINTER(ObsWire), i64, vp, vp, vp]
    L.vlr_obs_encode_batch.argtypes = [i64, vp, vp, vp, vp, vp, i64]
    L.vlr_posterior_batch_wire.argtypes = [vp, i64, vp, C.POINTER(ObsWire), vp, vp, vp, vp]
    L.vlr_model_leaves.argtypes = [vp, C.POINTER(Model), i64, vp, vp, vp]
    L.vlr_posterior_batch.argtypes = [vp, i64, C.POINTER(Model), C.POINTER(ObsColumns), vp, vp, vp,
                                      vp, vp]
    L.vlr_posterior_batch_f32.argtypes = [vp, i64, C.POINTER(Model), C.POINTER(ObsColumnsF32), vp, vp, vp,
                                          vp, vp]
    L.vlr_posterior_batch_dev.argtypes = [vp, i64, C.POINTER(Model), C.POINTER(ObsColumns), vp, i32,
                                          vp, vp, vp, vp]
    _lib = L
    return L
