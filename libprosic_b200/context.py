"""Host-side handle on one GPU context (vlr_ctx) with numpy / device-pointer entry points.

Mirrors the reference seams:
  * ``Context.pairhmm_batch``  <-> PairHMM::prob_related at realignment/mod.rs:439-443
  * ``Context.likelihood_batch`` / ``posterior_batch`` <-> Likelihood::compute / Model::compute
"""
import ctypes as C
import math

import numpy as np

from . import _ffi
from ._ffi import GapParams as _GapParams
from ._ffi import VlrError


class GapParams:
    """GapParams of realignment/pairhmm.rs:73-79, built from linear artifact rates like
    cli.rs:672-677 does (``LogProb::from(Prob(rate))``)."""

    def __init__(self, prob_insertion_artifact=2.8e-6, prob_deletion_artifact=5.1e-6,
                 prob_insertion_extend_artifact=0.0, prob_deletion_extend_artifact=0.0):
        def ln(p):
            return math.log(p) if p > 0.0 else -math.inf
        self.c = _GapParams(ln(prob_insertion_artifact), ln(prob_deletion_artifact),
                            ln(prob_insertion_extend_artifact), ln(prob_deletion_extend_artifact))

    def as_ln_array(self):
        return np.array([self.c.prob_gap_x, self.c.prob_gap_y, self.c.prob_gap_x_extend,
                         self.c.prob_gap_y_extend], dtype=np.float64)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def expand_indel_runs(runs, n_ops):
    """Decode the run-length encoded best_indel_operations of vlr_besthit_batch /
    vlr_allele_support_batch (one byte per run: bit 7 = Ins, bits 0-6 = length; include/vlr_gpu.h)
    into a zero padded matrix [n, max(32, max n_ops)] of operations (2 = Del, 3 = Ins).  Returns
    (ops, overflow): overflow[k] marks records whose alignment had more than 32 runs (their decoded
    prefix is shorter than n_ops[k])."""
    runs = np.asarray(runs, np.uint8).reshape(len(n_ops), -1)
    lens = (runs & 0x7f).astype(np.int64)
    tot = lens.sum(axis=1)
    n_ops = np.asarray(n_ops)
    width = max(32, int(tot.max()) if len(tot) else 0)
    out = np.zeros((len(n_ops), width), np.uint8)
    for k in np.nonzero(tot)[0]:
        out[k, :tot[k]] = np.repeat(np.where(runs[k] & 0x80, 3, 2).astype(np.uint8), lens[k])
    return out, tot != n_ops


class Context:
    def __init__(self, device=0):
        self.lib = _ffi.load()
        h = C.c_void_p()
        rc = self.lib.vlr_ctx_create(int(device), C.byref(h))
        if rc != _ffi.VLR_OK:
            raise VlrError(rc, "vlr_ctx_create(device=%d) failed: no CUDA device? "
                               "(there is no CPU fallback)" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.vlr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _ffi.VLR_OK:
            raise VlrError(rc, self.lib.vlr_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.vlr_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        self._check(self.lib.vlr_ctx_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.vlr_ctx_launch_count(self.h))

    def measure_fp64_peak(self):
        v = C.c_double()
        self._check(self.lib.vlr_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def ref_math(self, fn, x, on_device=True):
        """The library's deterministic libm (csrc/detmath.cuh): fn 0 exp, 1 expm1, 2 log, 3 log1p."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        self._check(self.lib.vlr_ref_math(self.h if on_device else None, int(fn), x.size, _ptr(x), _ptr(out)))
        return out

    # ---------------------------------------------------------------- pair-HMM
    def pairhmm_batch(self, hap_bytes, hap_off, read_bases, read_quals, read_off, gap,
                      pair_hap=None, pair_read=None, max_edit_dist=None, flags=_ffi.HMM_DEFAULT,
                      out=None):
        """ln P(read | haplotype) for every pair (host buffers; copies inside)."""
        hap_bytes = _arr(hap_bytes, np.uint8)
        hap_off = _arr(hap_off, np.int64)
        read_bases = _arr(read_bases, np.uint8)
        read_quals = _arr(read_quals, np.uint8)
        read_off = _arr(read_off, np.int64)
        pair_hap = _arr(pair_hap, np.int32)
        pair_read = _arr(pair_read, np.int32)
        max_edit_dist = _arr(max_edit_dist, np.int32)
        n_haps, n_reads = len(hap_off) - 1, len(read_off) - 1
        if pair_hap is not None:
            n_pairs = len(pair_hap)
        elif pair_read is not None:
            n_pairs = len(pair_read)
        else:
            n_pairs = min(n_haps, n_reads)
        if out is None:
            out = np.empty(n_pairs, dtype=np.float64)
        self._check(self.lib.vlr_pairhmm_batch(
            self.h, n_pairs, _ptr(hap_bytes), _ptr(hap_off), n_haps, _ptr(read_bases),
            _ptr(read_quals), _ptr(read_off), n_reads, _ptr(pair_hap), _ptr(pair_read),
            _ptr(max_edit_dist), C.byref(gap.c), int(flags), _ptr(out)))
        return out

    def besthit_batch(self, hap_bytes, hap_off, read_bases, read_off, pair_hap=None, pair_read=None):
        """calc_best_hit for every pair: dict(dist, start, end, n_best, n_indel_ops, indel_ops)."""
        hap_bytes = _arr(hap_bytes, np.uint8)
        hap_off = _arr(hap_off, np.int64)
        read_bases = _arr(read_bases, np.uint8)
        read_off = _arr(read_off, np.int64)
        pair_hap = _arr(pair_hap, np.int32)
        pair_read = _arr(pair_read, np.int32)
        n_haps, n_reads = len(hap_off) - 1, len(read_off) - 1
        n = len(pair_hap) if pair_hap is not None else (len(pair_read) if pair_read is not None
                                                        else min(n_haps, n_reads))
        out = {k: np.empty(n, np.int32) for k in ("dist", "start", "end", "n_best", "n_indel_ops")}
        ops = np.zeros((n, _ffi.MAX_INDEL_OPS), np.uint8)
        self._check(self.lib.vlr_besthit_batch(
            self.h, n, _ptr(hap_bytes), _ptr(hap_off), n_haps, _ptr(read_bases), _ptr(read_off),
            n_reads, _ptr(pair_hap), _ptr(pair_read), _ptr(out["dist"]), _ptr(out["start"]),
            _ptr(out["end"]), _ptr(out["n_best"]), _ptr(out["n_indel_ops"]), _ptr(ops)))
        out["indel_runs"] = ops
        out["indel_ops"], out["indel_overflow"] = expand_indel_runs(ops, out["n_indel_ops"])
        return out

    def pathhmm_batch(self, hap_bytes, hap_off, read_bases, read_quals, read_off, gap,
                      pair_hap=None, pair_read=None):
        """PathHMMRealigner::calculate_prob_allele (fast mode): (ln prob of the best path, dist)."""
        hap_bytes = _arr(hap_bytes, np.uint8)
        hap_off = _arr(hap_off, np.int64)
        read_bases = _arr(read_bases, np.uint8)
        read_quals = _arr(read_quals, np.uint8)
        read_off = _arr(read_off, np.int64)
        pair_hap = _arr(pair_hap, np.int32)
        pair_read = _arr(pair_read, np.int32)
        n_haps, n_reads = len(hap_off) - 1, len(read_off) - 1
        n = len(pair_hap) if pair_hap is not None else (len(pair_read) if pair_read is not None
                                                        else min(n_haps, n_reads))
        out, dist = np.empty(n, np.float64), np.empty(n, np.int32)
        self._check(self.lib.vlr_pathhmm_batch(
            self.h, n, _ptr(hap_bytes), _ptr(hap_off), n_haps, _ptr(read_bases), _ptr(read_quals),
            _ptr(read_off), n_reads, _ptr(pair_hap), _ptr(pair_read), C.byref(gap.c), _ptr(out),
            _ptr(dist)))
        return out, dist

    def allele_support_batch(self, region_read, region_hap_off, region_haps, region_n_ref, hap_bytes,
                             hap_off, read_bases, read_quals, read_off, gap, shrink_extra=None,
                             flags=_ffi.HMM_DEFAULT, out=None):
        """Fused Realigner::allele_support per merged region (see include/vlr_gpu.h).
        out: optional dict of preallocated (e.g. pinned) output arrays with the keys returned."""
        region_read = _arr(region_read, np.int32)
        region_hap_off = _arr(region_hap_off, np.int32)
        region_haps = _arr(region_haps, np.int32)
        region_n_ref = _arr(region_n_ref, np.int32)
        hap_bytes = _arr(hap_bytes, np.uint8)
        hap_off = _arr(hap_off, np.int64)
        read_bases = _arr(read_bases, np.uint8)
        read_quals = _arr(read_quals, np.uint8)
        read_off = _arr(read_off, np.int64)
        shrink_extra = _arr(shrink_extra, np.int32)
        n = len(region_read)
        decode = out is None
        if out is None:
            out = dict(prob_ref=np.empty(n), prob_alt=np.empty(n), raw_ref=np.empty(n), raw_alt=np.empty(n),
                       ref_dist=np.empty(n, np.int32), alt_dist=np.empty(n, np.int32),
                       n_indel_ops=np.empty(n, np.int32),
                       indel_ops=np.zeros((n, _ffi.MAX_INDEL_OPS), np.uint8))
        self._check(self.lib.vlr_allele_support_batch(
            self.h, n, _ptr(region_read), _ptr(region_hap_off), _ptr(region_haps), _ptr(region_n_ref),
            _ptr(hap_bytes), _ptr(hap_off), len(hap_off) - 1, _ptr(shrink_extra), _ptr(read_bases),
            _ptr(read_quals), _ptr(read_off), len(read_off) - 1, C.byref(gap.c), int(flags),
            _ptr(out["prob_ref"]), _ptr(out["prob_alt"]), _ptr(out["raw_ref"]), _ptr(out["raw_alt"]),
            _ptr(out["ref_dist"]), _ptr(out["alt_dist"]), _ptr(out["n_indel_ops"]),
            _ptr(out["indel_ops"])))
        if decode:  # caller-provided (pinned) buffers keep the run-length encoded records
            out["indel_runs"] = out["indel_ops"]
            out["indel_ops"], out["indel_overflow"] = expand_indel_runs(out["indel_runs"], out["n_indel_ops"])
        return out

    def pairhmm_workspace_bytes(self, n_pairs, max_len_x=0, max_len_y=0):
        return int(self.lib.vlr_pairhmm_workspace_bytes(n_pairs, max_len_x, max_len_y))

    def pairhmm_batch_dev(self, n_pairs, d_hap, d_hap_off, n_haps, d_rb, d_rq, d_read_off, n_reads,
                          d_pair_hap, d_pair_read, d_med, gap, flags, d_ws, ws_bytes, d_out):
        """All arguments are raw device pointers (ints); enqueues on the bound stream."""
        vp = C.c_void_p
        self._check(self.lib.vlr_pairhmm_batch_dev(
            self.h, n_pairs, vp(d_hap), vp(d_hap_off), n_haps, vp(d_rb), vp(d_rq), vp(d_read_off),
            n_reads, vp(d_pair_hap or 0), vp(d_pair_read or 0), vp(d_med or 0), C.byref(gap.c),
            int(flags), vp(d_ws), ws_bytes, vp(d_out)))

    # ---------------------------------------------------------------- likelihood + posterior
    @staticmethod
    def _obs_columns(cols, cat, keep):
        from ._ffi import ObsColumns
        names = ("prob_mapping", "prob_mismapping", "prob_alt", "prob_ref", "prob_missed_allele",
                 "prob_sample_alt", "prob_double_overlap", "prob_single_overlap", "prob_hit_base")
        arrs = [np.ascontiguousarray(cols[n], dtype=np.float64) for n in names]
        catc = np.ascontiguousarray(cat, dtype=np.uint16)
        keep.extend(arrs + [catc])
        return ObsColumns(*[a.ctypes.data for a in arrs], catc.ctypes.data)

    @staticmethod
    def build_model(scn, keep):
        """dict scenario (libprosic_b200.scenario) -> ctypes vlr_model."""
        from ._ffi import Biases, Event, Model, Node
        nodes = (Node * max(1, len(scn["nodes"])))()
        children, vafs = [], []
        for k, nd in enumerate(scn["nodes"]):
            nodes[k] = Node(nd["kind"], nd.get("sample", 0), len(children), len(nd["children"]),
                            len(vafs), len(nd["vafs"]))
            children.extend(nd["children"])
            vafs.extend(nd["vafs"])
        roots, flat = [], []
        events = (Event * len(scn["events"]))()
        for k, ev in enumerate(scn["events"]):
            events[k] = Event(len(roots), len(ev["roots"]), len(flat), len(ev["biases"]))
            roots.extend(ev["roots"])
            flat.extend(ev["biases"])
        biases = (Biases * max(1, len(flat)))()
        for k, bi in enumerate(flat):
            b = scn["biases"][bi]
            biases[k] = Biases(b["strand"], b["orientation"], b["read_position"], b["softclip"],
                               b["divindel"], 0, b["min_other_rate"])
        res = np.asarray(scn["resolution"], np.int32)
        cont = np.asarray(scn["contaminated"], np.int32)
        by = np.asarray(scn["by"], np.int32)
        pur = np.asarray(scn["purity"], np.float64)
        ch = np.asarray(children if children else [0], np.int32)
        vf = np.asarray(vafs if vafs else [0.0], np.float64)
        rt = np.asarray(roots, np.int32)
        keep.extend([nodes, events, biases, res, cont, by, pur, ch, vf, rt])
        prior = scn.get("prior_ln")  # ln prior per leaf in vlr_model_leaves order (None = flat)
        if prior is not None:
            prior = np.ascontiguousarray(prior, dtype=np.float64)
            keep.append(prior)
        prog_ptr = None
        pp = scn.get("prior_prog")  # dict form of vlr_prior_prog (libprosic_b200.prior.make_prior)
        if pp is not None:
            from ._ffi import PriorProg
            ns = scn["n_samples"]
            prog = PriorProg()
            for k in range(ns):
                prog.kind[k], prog.ploidy[k] = pp["kind"][k], pp["ploidy"][k]
                prog.som_kind[k], prog.parent[k] = pp["som_kind"][k], pp["parent"][k]
                prog.ln_rate[k], prog.som_zero[k] = pp["ln_rate"][k], pp["som_zero"][k]
            prog.ln_genome_size = pp["ln_genome_size"]
            uni, uoff = [], [0]
            for k in range(ns):
                uni.extend(pp["uni"][k])
                uoff.append(len(uni))
            for k in range(9):
                prog.uni_off[k] = uoff[min(k, ns)]
            us = np.asarray(uni if uni else [(0.0,) * 5], np.float64).reshape(-1, 5)
            gl = np.asarray(pp["germ_ln"], np.float64)
            prog.uni_spec, prog.germ_ln, prog.n_germ = us.ctypes.data, gl.ctypes.data, len(gl)
            keep.extend([prog, us, gl])
            prog_ptr = C.addressof(prog)
        return Model(scn["n_samples"], len(scn["events"]), len(scn["nodes"]), len(flat),
                     res.ctypes.data, cont.ctypes.data, by.ctypes.data, pur.ctypes.data,
                     C.addressof(nodes), ch.ctypes.data, vf.ctypes.data, rt.ctypes.data,
                     C.addressof(events), C.addressof(biases),
                     prior.ctypes.data if prior is not None else None,
                     len(prior) if prior is not None else 0, prog_ptr)

    def call_records(self, scn, post, prob_mapping, obs_off):
        """BCF values of the call records: dict(prob[n_sites, n_named + 1] f32 PHRED, af f32, dp i32,
        names).  post = the dict posterior_batch returned."""
        is_art = np.array([any(scn["biases"][b][k] for k in ("strand", "orientation", "read_position",
                                                             "softclip", "divindel"))
                           for b in (ev["biases"][0] for ev in scn["events"])], np.uint8)
        ev = np.ascontiguousarray(post["event_ln"], np.float64)
        n_sites, ne = ev.shape
        ns = scn["n_samples"]
        marg = np.ascontiguousarray(post["marginal"], np.float64)
        maf = np.ascontiguousarray(post["map_af"], np.float64)
        pm = np.ascontiguousarray(prob_mapping, np.float64)
        off = np.ascontiguousarray(obs_off, np.int64)
        n_named = int((is_art == 0).sum())
        prob = np.empty((n_sites, n_named + 1), np.float32)
        af = np.empty((n_sites, ns), np.float32)
        dp = np.empty((n_sites, ns), np.int32)
        self._check(self.lib.vlr_call_records(self.h, n_sites, ne, ns, _ptr(is_art), _ptr(ev), _ptr(marg),
                                              _ptr(maf), _ptr(pm), _ptr(off), _ptr(prob), _ptr(af), _ptr(dp)))
        names = [e["name"] for e, a in zip(scn["events"], is_art) if not a] + ["artifact"]
        return dict(prob=prob, af=af, dp=dp, names=names)

    def call_obs_strings(self, cols, cat, obs_off):
        """OBS and SOBS strings of every (site, sample) pileup: list of (obs, sobs) in obs_off order."""
        pa = np.ascontiguousarray(cols["prob_alt"], np.float64)
        pr = np.ascontiguousarray(cols["prob_ref"], np.float64)
        pm = np.ascontiguousarray(cols["prob_mapping"], np.float64)
        ct = np.ascontiguousarray(cat, np.uint16)
        codes = np.empty(len(ct), np.uint16)
        self._check(self.lib.vlr_obs_codes(self.h, len(ct), _ptr(pa), _ptr(pr), _ptr(pm), _ptr(ct), _ptr(codes)))
        out = []
        buf = C.create_string_buffer(1 << 16)
        for k in range(len(obs_off) - 1):
            seg = np.ascontiguousarray(codes[int(obs_off[k]):int(obs_off[k + 1])])
            pair = []
            for simple in (0, 1):
                n = self.lib.vlr_obs_cigar(_ptr(seg) if len(seg) else None, len(seg), simple, buf, len(buf))
                if n < 0:
                    raise _ffi.VlrError(int(n), "vlr_obs_cigar")
                pair.append(buf.value.decode())
            out.append(tuple(pair))
        return out

    def bias_chars(self, b):
        """SB ROB RPB SCB DIB flag characters of a scenario bias dict."""
        from ._ffi import Biases
        bb = Biases(b["strand"], b["orientation"], b["read_position"], b["softclip"], b["divindel"], 0, b["min_other_rate"])
        out = C.create_string_buffer(5)
        self._check(self.lib.vlr_bias_chars(C.byref(bb), out))
        return out.raw.decode()

    def model_leaves(self, scn):
        """Leaf AF vectors of a set-only scenario: (event index per leaf, afs[n_leaves, n_samples])."""
        keep = []
        scn = dict(scn)
        scn.pop("prior_ln", None)
        model = self.build_model(scn, keep)
        n = C.c_int64()
        self._check(self.lib.vlr_model_leaves(self.h, C.byref(model), 0, C.byref(n), None, None))
        ev = np.empty(n.value, np.int32)
        afs = np.empty((n.value, scn["n_samples"]), np.float64)
        self._check(self.lib.vlr_model_leaves(self.h, C.byref(model), n.value, C.byref(n), _ptr(ev), _ptr(afs)))
        return ev, afs

    @staticmethod
    def _posterior_out(out, n_sites, ne, ns):
        """Output arrays of the posterior entries: the caller's (pinned) buffers or fresh ones."""
        if out is None:
            out = dict(event_ln=np.empty((n_sites, ne), np.float64), marginal=np.empty(n_sites, np.float64),
                       map_af=np.empty((n_sites, ns), np.float64), map_bias=np.empty((n_sites, ns), np.int32))
        for k, shape, dt in (("event_ln", (n_sites, ne), np.float64), ("marginal", (n_sites,), np.float64),
                             ("map_af", (n_sites, ns), np.float64), ("map_bias", (n_sites, ns), np.int32)):
            a = out[k]
            if a.dtype != dt or a.size != int(np.prod(shape)) or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("out[%r] must be a C-contiguous %s array of %d elements" % (k, np.dtype(dt).name, int(np.prod(shape))))
        return out

    def posterior_batch(self, scn, cols, cat, obs_off, out=None):
        """Model::compute for every site: returns dict(event_ln[n_sites, n_events], marginal,
        map_af[n_sites, n_samples], map_bias).  `out`: optional dict of preallocated (pinned) arrays."""
        keep = []
        model = self._cached_model(scn)
        oc = self._obs_columns(cols, cat, keep)
        obs_off = np.ascontiguousarray(obs_off, dtype=np.int64)
        ns, ne = scn["n_samples"], len(scn["events"])
        n_sites = (len(obs_off) - 1) // ns
        out = self._posterior_out(out, n_sites, ne, ns)
        self._check(self.lib.vlr_posterior_batch(self.h, n_sites, C.byref(model), C.byref(oc),
                                                 _ptr(obs_off), _ptr(out["event_ln"]), _ptr(out["marginal"]),
                                                 _ptr(out["map_af"]), _ptr(out["map_bias"])))
        return out

    def posterior_batch_f32(self, scn, cols32, cat, obs_off, out=None):
        """Compact variant: the seven stored fields as float32 arrays (dict), derived columns on the
        device.  Same outputs as posterior_batch."""
        from ._ffi import ObsColumnsF32
        model = self._cached_model(scn)
        names = ("prob_mapping", "prob_alt", "prob_ref", "prob_missed_allele", "prob_sample_alt",
                 "prob_double_overlap", "prob_hit_base")
        arrs = [np.ascontiguousarray(cols32[n], dtype=np.float32) for n in names]
        catc = np.ascontiguousarray(cat, dtype=np.uint16)
        oc = ObsColumnsF32(*[a.ctypes.data for a in arrs], catc.ctypes.data)
        obs_off = np.ascontiguousarray(obs_off, dtype=np.int64)
        ns, ne = scn["n_samples"], len(scn["events"])
        n_sites = (len(obs_off) - 1) // ns
        out = self._posterior_out(out, n_sites, ne, ns)
        self._check(self.lib.vlr_posterior_batch_f32(self.h, n_sites, C.byref(model), C.byref(oc),
                                                     _ptr(obs_off), _ptr(out["event_ln"]), _ptr(out["marginal"]),
                                                     _ptr(out["map_af"]), _ptr(out["map_bias"])))
        return out

    def posterior_batch_wire(self, scn, wire, out=None):
        """Part 2 straight from stored observation records (obs_codec.WireBatch): upload, device
        decode, posterior.  Same outputs as posterior_batch.  `out`: optional dict of preallocated
        (pinned) arrays event_ln / marginal / map_af / map_bias."""
        model = self._cached_model(scn)
        ns, ne = scn["n_samples"], len(scn["events"])
        n_sites = wire.n_records // ns
        out = self._posterior_out(out, n_sites, ne, ns)
        ws = wire.c_struct()
        self._check(self.lib.vlr_posterior_batch_wire(self.h, n_sites, C.byref(model), C.byref(ws), _ptr(out["event_ln"]),
                                                      _ptr(out["marginal"]), _ptr(out["map_af"]), _ptr(out["map_bias"])))
        return out

    def _cached_model(self, scn):
        """vlr_model of a scenario dict, rebuilt when the dict changed.  The cache belongs to this
        context, holds the dict itself (its id cannot be reused while cached) and compares a
        fingerprint of the fields callers mutate in place (prior, resolutions, purities)."""
        import json
        cache = self.__dict__.setdefault("_model_cache", {})
        fp = json.dumps([scn["n_samples"], list(map(int, scn["resolution"])), list(map(float, scn["purity"])),
                         list(map(int, scn["contaminated"])), list(map(int, scn["by"])), scn["nodes"], scn["events"],
                         scn["biases"]], sort_keys=True, default=str) + \
            "|%d|%d" % (id(scn.get("prior_ln")), id(scn.get("prior_prog")))
        ent = cache.get(id(scn))
        if ent is None or ent[0] is not scn or ent[1] != fp:
            keep = []
            ent = (scn, fp, self.build_model(scn, keep), keep, scn.get("prior_ln"), scn.get("prior_prog"))
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[id(scn)] = ent
        return ent[2]

    def posterior_batch_dev(self, scn, n_sites, d_cols, d_cat, d_obs_off, max_obs, d_ev, d_marg, d_maf, d_mb):
        """Device-pointer variant (ints).  d_cols: dict name -> device pointer."""
        from ._ffi import ObsColumns
        model = self._cached_model(scn)
        names = ("prob_mapping", "prob_mismapping", "prob_alt", "prob_ref", "prob_missed_allele",
                 "prob_sample_alt", "prob_double_overlap", "prob_single_overlap", "prob_hit_base")
        oc = ObsColumns(*[d_cols[n] for n in names], d_cat)
        vp = C.c_void_p
        self._check(self.lib.vlr_posterior_batch_dev(self.h, n_sites, C.byref(model), C.byref(oc),
                                                     vp(d_obs_off), int(max_obs), vp(d_ev), vp(d_marg),
                                                     vp(d_maf), vp(d_mb)))

    def likelihood_batch(self, n_sites, n_samples, cols, cat, obs_off, biases, items):
        """Likelihood::compute per work item.  biases: list of scenario bias dicts; items: list of
        dict(site, sample, bias, contaminated, af, af_secondary, purity, other_rate)."""
        from ._ffi import Biases, LhItem
        keep = []
        oc = self._obs_columns(cols, cat, keep)
        obs_off = np.ascontiguousarray(obs_off, dtype=np.int64)
        barr = (Biases * len(biases))()
        for k, b in enumerate(biases):
            barr[k] = Biases(b["strand"], b["orientation"], b["read_position"], b["softclip"],
                             b["divindel"], 0, b["min_other_rate"])
        iarr = (LhItem * max(1, len(items)))()
        for k, it in enumerate(items):
            iarr[k] = LhItem(it["site"], it["sample"], it["bias"], int(it.get("contaminated", 0)),
                             it["af"], it.get("af_secondary", 0.0), it.get("purity", 1.0),
                             it.get("other_rate", 0.0))
        out = np.empty(len(items), np.float64)
        self._check(self.lib.vlr_likelihood_batch(self.h, n_sites, n_samples, C.byref(oc),
                                                  _ptr(obs_off), C.addressof(barr), len(biases),
                                                  C.addressof(iarr), len(items), _ptr(out)))
        return out
