"""ctypes binding of the CPU oracle (oracle/vlr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  The product package ``libprosic_b200`` never imports it.
PARITY UNPINNED against the real Rust binary (see vlr_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvlr_oracle.so")

HMM_SUM3_APPROX = 1      # two-way ln_sum3_exp_approx (round-1 recollection)
HMM_PATH_CLOSE = 2
HMM_SUM3_APPROX3 = 8     # three-way ln_sum3_exp_approx (default)
HMM_RESET_FM_ONLY = 16   # oracle only: reset fm[curr] alone after a column swap
HMM_DEFAULT = HMM_SUM3_APPROX3

NODE_SET, NODE_RANGE, NODE_FALSE, NODE_SKIP = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "vlr_oracle.c")
    hdr = os.path.join(_HERE, "vlr_oracle.h")
    dmh = os.path.join(_HERE, "detmath.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(dmh))):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B", "libvlr_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _SO


class Hit(C.Structure):
    _fields_ = [("dist", C.c_int32), ("start", C.c_int32), ("end", C.c_int32),
                ("n_best", C.c_int32), ("n_indel_ops", C.c_int32)]


class Support(C.Structure):
    _fields_ = [("prob_ref", C.c_double), ("prob_alt", C.c_double), ("raw_ref", C.c_double),
                ("raw_alt", C.c_double), ("informative", C.c_int32), ("alt_dist", C.c_int32),
                ("ref_dist", C.c_int32), ("n_indel_ops", C.c_int32)]


_DP = C.POINTER(C.c_double)
_U16P = C.POINTER(C.c_uint16)
_I32P = C.POINTER(C.c_int32)
_U8P = C.POINTER(C.c_uint8)


class Pileup(C.Structure):
    _fields_ = [(n, _DP) for n in ("prob_mapping", "prob_mismapping", "prob_alt", "prob_ref",
                                   "prob_missed_allele", "prob_sample_alt", "prob_double_overlap",
                                   "prob_single_overlap", "prob_hit_base")] + \
               [("cat", _U16P), ("n_obs", C.c_int32)]


class Biases(C.Structure):
    _fields_ = [("strand", C.c_int32), ("orientation", C.c_int32), ("read_position", C.c_int32),
                ("softclip", C.c_int32), ("divindel", C.c_int32), ("other_rate", C.c_double),
                ("min_other_rate", C.c_double)]


class Node(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sample", C.c_int32), ("child_off", C.c_int32),
                ("n_children", C.c_int32), ("vaf_off", C.c_int32), ("n_vafs", C.c_int32)]


class Event(C.Structure):
    _fields_ = [("root_off", C.c_int32), ("n_roots", C.c_int32), ("bias_off", C.c_int32),
                ("n_biases", C.c_int32)]


class Model(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("resolution", _I32P), ("contaminated", _I32P),
                ("by", _I32P), ("purity", _DP), ("nodes", C.POINTER(Node)), ("children", _I32P),
                ("vafs", _DP), ("roots", _I32P), ("events", C.POINTER(Event)),
                ("n_events", C.c_int32), ("biases", C.POINTER(Biases)),
                ("n_biases_total", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    d = C.c_double
    L.vlro_ref_math.restype = None
    L.vlro_ref_math.argtypes = [C.c_int, C.c_long, _DP, _DP]
    L.vlro_ln_add_exp.restype = d
    L.vlro_ln_add_exp.argtypes = [d, d]
    L.vlro_ln_sum_exp.restype = d
    L.vlro_ln_sum_exp.argtypes = [_DP, C.c_size_t]
    L.vlro_ln_one_minus_exp.restype = d
    L.vlro_ln_one_minus_exp.argtypes = [d]
    L.vlro_ln_prob_miscall.restype = d
    L.vlro_ln_prob_miscall.argtypes = [C.c_uint8]
    L.vlro_ln_prob_call.restype = d
    L.vlro_ln_prob_call.argtypes = [C.c_uint8]
    L.vlro_minilogprob_roundtrip.restype = d
    L.vlro_minilogprob_roundtrip.argtypes = [d]
    L.vlro_f64_to_f16_bits.restype = C.c_uint16
    L.vlro_f64_to_f16_bits.argtypes = [d]
    L.vlro_f16_bits_to_f64.restype = d
    L.vlro_f16_bits_to_f64.argtypes = [C.c_uint16]
    L.vlro_pairhmm_prob_related.restype = d
    L.vlro_pairhmm_prob_related.argtypes = [_U8P, C.c_int, _U8P, _U8P, C.c_int, _DP, C.c_int,
                                            C.c_uint]
    L.vlro_pairhmm_bruteforce.restype = d
    L.vlro_pairhmm_bruteforce.argtypes = [_U8P, C.c_int, _U8P, _U8P, C.c_int, _DP, C.c_uint]
    L.vlro_best_hit.restype = C.c_int
    L.vlro_best_hit.argtypes = [_U8P, C.c_int, _U8P, C.c_int, C.POINTER(Hit), _U8P, C.c_int,
                                _I32P, C.c_int, _U8P, _I32P, C.c_int]
    L.vlro_pathhmm_prob.restype = d
    L.vlro_pathhmm_prob.argtypes = [_U8P, C.c_int, _U8P, _U8P, C.c_int, _DP, C.c_int, _I32P,
                                    _U8P, _I32P]
    L.vlro_allele_support_region.restype = C.c_int
    L.vlro_allele_support_region.argtypes = [_U8P, _I32P, _I32P, _I32P, _U8P, _U8P, C.c_int, _DP,
                                             C.c_uint, C.c_int, C.POINTER(Support), _U8P, C.c_int]
    L.vlro_biases_prob.restype = d
    L.vlro_biases_prob.argtypes = [C.POINTER(Biases), C.POINTER(Pileup), C.c_int]
    L.vlro_biases_prob_any.restype = d
    L.vlro_biases_prob_any.argtypes = [C.POINTER(Biases), C.POINTER(Pileup), C.c_int]
    L.vlro_biases_gate.restype = C.c_int
    L.vlro_biases_gate.argtypes = [C.POINTER(Biases), C.POINTER(Pileup), C.c_int]
    L.vlro_biases_learn.restype = None
    L.vlro_biases_learn.argtypes = [C.POINTER(Biases), C.POINTER(Pileup), C.c_int]
    L.vlro_is_strong_obs.restype = C.c_int
    L.vlro_is_strong_obs.argtypes = [C.POINTER(Pileup), C.c_int]
    L.vlro_likelihood_single.restype = d
    L.vlro_likelihood_single.argtypes = [C.POINTER(Pileup), d, C.POINTER(Biases)]
    L.vlro_likelihood_contaminated.restype = d
    L.vlro_likelihood_contaminated.argtypes = [C.POINTER(Pileup), d, d, C.POINTER(Biases), d,
                                               C.POINTER(Biases)]
    L.vlro_observable_min.restype = d
    L.vlro_observable_min.argtypes = [d, d, C.c_int, C.c_int, C.c_int]
    L.vlro_observable_max.restype = d
    L.vlro_observable_max.argtypes = [d, d, C.c_int, C.c_int, C.c_int]
    L.vlro_grid_points.restype = C.c_int
    L.vlro_grid_points.argtypes = [C.c_int, C.c_int]
    L.vlro_model_compute.restype = C.c_int
    L.vlro_model_compute.argtypes = [C.POINTER(Model), C.POINTER(Pileup), _DP, _DP, _DP, _I32P,
                                     C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_U8P)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_DP)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_I32P)


def ref_math(fn, x):
    """The oracle's libm (oracle/detmath.h): fn 0 exp, 1 expm1, 2 log, 3 log1p."""
    a, ap = _f64(x)
    out = np.empty_like(a)
    lib().vlro_ref_math(int(fn), a.size, ap, out.ctypes.data_as(_DP))
    return out


def pairhmm_prob_related(hap, read, qual, gap_ln, max_edit_dist=-1, flags=HMM_DEFAULT):
    h, hp = _u8(hap)
    r, rp = _u8(read)
    q, qp = _u8(qual)
    g, gp = _f64(gap_ln)
    return lib().vlro_pairhmm_prob_related(hp, len(h), rp, qp, len(r), gp, int(max_edit_dist),
                                           int(flags))


def pairhmm_bruteforce(hap, read, qual, gap_ln, flags=0):
    h, hp = _u8(hap)
    r, rp = _u8(read)
    q, qp = _u8(qual)
    g, gp = _f64(gap_ln)
    return lib().vlro_pairhmm_bruteforce(hp, len(h), rp, qp, len(r), gp, int(flags))


def pairhmm_batch(haps, hap_off, reads, quals, read_off, gap_ln, max_edit_dist=None,
                  flags=HMM_DEFAULT):
    """Loop the scalar oracle over a CSR batch (same layout as vlr_pairhmm_batch)."""
    n = len(hap_off) - 1
    out = np.empty(n, dtype=np.float64)
    for k in range(n):
        med = -1 if max_edit_dist is None else int(max_edit_dist[k])
        out[k] = pairhmm_prob_related(haps[hap_off[k]:hap_off[k + 1]],
                                      reads[read_off[k]:read_off[k + 1]],
                                      quals[read_off[k]:read_off[k + 1]], gap_ln, med, flags)
    return out


def best_hit(hap, read, want_alignments=False):
    h, hp = _u8(hap)
    r, rp = _u8(read)
    hit = Hit()
    cap = len(h) + 1
    ops_cap = 512
    ops = np.zeros(ops_cap, dtype=np.uint8)
    starts = np.zeros(cap, dtype=np.int32)
    aops_cap = cap * (len(h) + len(r) + 2) if want_alignments else 1
    aops = np.zeros(aops_cap, dtype=np.uint8)
    aoff = np.zeros(cap + 1, dtype=np.int32)
    ok = lib().vlro_best_hit(hp, len(h), rp, len(r), C.byref(hit), ops.ctypes.data_as(_U8P),
                             ops_cap, starts.ctypes.data_as(_I32P), cap,
                             aops.ctypes.data_as(_U8P) if want_alignments else None,
                             aoff.ctypes.data_as(_I32P) if want_alignments else None, aops_cap)
    if not ok:
        return None
    res = dict(dist=hit.dist, start=hit.start, end=hit.end, n_best=hit.n_best,
               indel_ops=ops[:hit.n_indel_ops].copy(), starts=starts[:hit.n_best].copy())
    if want_alignments:
        res["alignments"] = [aops[aoff[k]:aoff[k + 1]].copy() for k in range(hit.n_best)]
    return res


def pathhmm_prob(hap, read, qual, gap_ln):
    """PathHMMRealigner::calculate_prob_allele on the best hit of (read, hap); None without a hit."""
    hit = best_hit(hap, read, want_alignments=True)
    if hit is None:
        return None
    h, hp = _u8(hap)
    r, rp = _u8(read)
    q, qp = _u8(qual)
    g, gp = _f64(gap_ln)
    starts = np.ascontiguousarray(hit["starts"], dtype=np.int32)
    ops = np.concatenate(hit["alignments"]).astype(np.uint8) if hit["alignments"] else np.zeros(1, np.uint8)
    off = np.concatenate([[0], np.cumsum([len(a) for a in hit["alignments"]])]).astype(np.int32)
    fn = lib().vlro_pathhmm_prob
    fn.restype = C.c_double
    return float(fn(hp, len(h), rp, qp, len(r), gp, len(starts), starts.ctypes.data_as(_I32P),
                    ops.ctypes.data_as(_U8P), off.ctypes.data_as(_I32P)))


def allele_support_region(ref_haps, alt_haps, read, qual, gap_ln, shrink_extra=None,
                          flags=HMM_DEFAULT, fast_mode=False):
    """ref_haps / alt_haps: lists of uint8 arrays.  shrink_extra: list (ref haps then alt haps)."""
    allh = list(ref_haps) + list(alt_haps)
    off = np.zeros(len(allh) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(h) for h in allh])
    flat = np.concatenate([np.asarray(h, dtype=np.uint8) for h in allh])
    if shrink_extra is None:
        shrink_extra = [0] * len(allh)
    f, fp = _u8(flat)
    o, op = _i32(off)
    nh, nhp = _i32([len(ref_haps), len(alt_haps)])
    se, sep = _i32(shrink_extra)
    r, rp = _u8(read)
    q, qp = _u8(qual)
    g, gp = _f64(gap_ln)
    out = Support()
    ops = np.zeros(512, dtype=np.uint8)
    lib().vlro_allele_support_region(fp, op, nhp, sep, rp, qp, len(r), gp, int(flags),
                                     int(fast_mode), C.byref(out), ops.ctypes.data_as(_U8P), 512)
    return dict(prob_ref=out.prob_ref, prob_alt=out.prob_alt, raw_ref=out.raw_ref,
                raw_alt=out.raw_alt, informative=bool(out.informative), ref_dist=out.ref_dist,
                alt_dist=out.alt_dist, indel_ops=ops[:out.n_indel_ops].copy())


OBS_COLUMNS = ("prob_mapping", "prob_mismapping", "prob_alt", "prob_ref", "prob_missed_allele",
               "prob_sample_alt", "prob_double_overlap", "prob_single_overlap", "prob_hit_base")


def pack_cat(strand, orientation, read_position, softclipped, indel_ops, paired):
    """cat bit layout, see VLRO_CAT_* in vlr_oracle.h."""
    return (np.asarray(strand, np.uint16) | (np.asarray(orientation, np.uint16) << 2)
            | (np.asarray(read_position, np.uint16) << 6)
            | (np.asarray(softclipped, np.uint16) << 7) | (np.asarray(indel_ops, np.uint16) << 8)
            | (np.asarray(paired, np.uint16) << 10)).astype(np.uint16)


class PileupView:
    """Keeps numpy column arrays alive and exposes a ctypes Pileup over rows [lo, hi)."""

    def __init__(self, cols, cat, lo=0, hi=None):
        hi = len(cat) if hi is None else hi
        self.keep = [np.ascontiguousarray(cols[n][lo:hi], dtype=np.float64) for n in OBS_COLUMNS]
        self.cat = np.ascontiguousarray(cat[lo:hi], dtype=np.uint16)
        self.c = Pileup(*[a.ctypes.data_as(_DP) for a in self.keep],
                        self.cat.ctypes.data_as(_U16P), hi - lo)


def make_biases(strand=0, orientation=0, read_position=0, softclip=0, divindel=0, other_rate=0.0,
                min_other_rate=0.0):
    return Biases(strand, orientation, read_position, softclip, divindel, other_rate,
                  min_other_rate)


def likelihood_single(pv, af, biases):
    return lib().vlro_likelihood_single(C.byref(pv.c), float(af), C.byref(biases))


def likelihood_contaminated(pv, purity, af_p, b_p, af_s, b_s):
    return lib().vlro_likelihood_contaminated(C.byref(pv.c), float(purity), float(af_p),
                                              C.byref(b_p), float(af_s), C.byref(b_s))


class ModelSpec:
    """Flattened scenario (events, VAF trees, bias configs) shared by oracle and GPU tests.

    nodes: list of dicts {kind, sample, children:[idx], vafs:[...]} ; events: list of dicts
    {roots:[node idx], biases:[bias idx]}; biases: list of Biases.
    """

    def __init__(self, n_samples, resolution, contaminated, by, purity, nodes, events, biases):
        self.n_samples = n_samples
        self.resolution = np.asarray(resolution, np.int32)
        self.contaminated = np.asarray(contaminated, np.int32)
        self.by = np.asarray(by, np.int32)
        self.purity = np.asarray(purity, np.float64)
        self.nodes_py, self.events_py, self.biases_py = nodes, events, biases
        children, vafs = [], []
        self.nodes = (Node * max(len(nodes), 1))()
        for k, nd in enumerate(nodes):
            self.nodes[k] = Node(nd["kind"], nd.get("sample", 0), len(children),
                                 len(nd.get("children", [])), len(vafs), len(nd.get("vafs", [])))
            children.extend(nd.get("children", []))
            vafs.extend(nd.get("vafs", []))
        self.children = np.asarray(children if children else [0], np.int32)
        self.vafs = np.asarray(vafs if vafs else [0.0], np.float64)
        roots = []
        boff = 0
        self.events = (Event * max(len(events), 1))()
        flat_biases = []
        for k, ev in enumerate(events):
            self.events[k] = Event(len(roots), len(ev["roots"]), len(flat_biases),
                                   len(ev["biases"]))
            roots.extend(ev["roots"])
            flat_biases.extend(ev["biases"])
        self.roots = np.asarray(roots if roots else [0], np.int32)
        self.flat_bias_idx = flat_biases  # event-local bias -> index into biases list
        self.biases = (Biases * max(len(flat_biases), 1))()
        for k, bi in enumerate(flat_biases):
            b = biases[bi]
            self.biases[k] = Biases(b.strand, b.orientation, b.read_position, b.softclip,
                                    b.divindel, b.other_rate, b.min_other_rate)
        self.c = Model(n_samples, self.resolution.ctypes.data_as(_I32P),
                       self.contaminated.ctypes.data_as(_I32P), self.by.ctypes.data_as(_I32P),
                       self.purity.ctypes.data_as(_DP), self.nodes,
                       self.children.ctypes.data_as(_I32P), self.vafs.ctypes.data_as(_DP),
                       self.roots.ctypes.data_as(_I32P), self.events, len(events), self.biases,
                       len(flat_biases))


PRIOR_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.c_void_p)


def model_compute_batch(spec, cols, cat, obs_off, n_sites, n_threads=1):
    """Model::compute over the first n_sites sites of a struct-of-arrays batch on n_threads host
    threads (flat prior).  Returns dict(event_ln [n, E], marginal, map_af [n, S], map_bias, n_evals)."""
    L = lib()
    ns, ne = spec.n_samples, len(spec.events_py)
    keep = [np.ascontiguousarray(cols[n], dtype=np.float64) for n in OBS_COLUMNS]
    catc = np.ascontiguousarray(cat, dtype=np.uint16)
    off = np.ascontiguousarray(obs_off, dtype=np.int64)
    base = Pileup(*[a.ctypes.data_as(_DP) for a in keep], catc.ctypes.data_as(_U16P), 0)
    ev = np.zeros((n_sites, ne), np.float64)
    marg = np.zeros(n_sites, np.float64)
    maf = np.zeros((n_sites, ns), np.float64)
    mb = np.zeros((n_sites, ns), np.int32)
    nev = np.zeros(n_sites, np.int64)
    L.vlro_model_compute_batch.restype = C.c_int
    L.vlro_model_compute_batch.argtypes = [C.POINTER(Model), C.POINTER(Pileup), C.POINTER(C.c_int64), C.c_int64,
                                           C.c_int, _DP, _DP, _DP, _I32P, C.POINTER(C.c_int64)]
    st = L.vlro_model_compute_batch(C.byref(spec.c), C.byref(base), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                    n_sites, int(n_threads), ev.ctypes.data_as(_DP), marg.ctypes.data_as(_DP),
                                    maf.ctypes.data_as(_DP), mb.ctypes.data_as(_I32P),
                                    nev.ctypes.data_as(C.POINTER(C.c_int64)))
    return dict(event_ln=ev, marginal=marg, map_af=maf, map_bias=mb, n_evals=nev, status=st)


def model_compute(spec, pileup_views, prior=None):
    """Returns dict(event_ln, marginal, map_af, map_bias, n_evals, status).
    prior: optional callable(list of allele frequencies) -> ln prior (Prior::compute)."""
    ns = spec.n_samples
    cb = PRIOR_FN(lambda afs, n, _u: float(prior([afs[k] for k in range(n)]))) if prior else None
    arr = (Pileup * ns)(*[pv.c for pv in pileup_views])
    ev = np.zeros(len(spec.events_py), np.float64)
    marg = C.c_double()
    maf = np.zeros(ns, np.float64)
    mb = np.zeros(ns, np.int32)
    ne = C.c_int64()
    st = lib().vlro_model_compute(C.byref(spec.c), arr, ev.ctypes.data_as(_DP), C.byref(marg),
                                  maf.ctypes.data_as(_DP), mb.ctypes.data_as(_I32P), C.byref(ne),
                                  C.cast(cb, C.c_void_p) if cb else None, None)
    return dict(event_ln=ev, marginal=marg.value, map_af=maf, map_bias=mb, n_evals=ne.value,
                status=st)
