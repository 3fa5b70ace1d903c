"""Host glue of `varlociraptor preprocess` + `call` for the reference's bundled tumor/normal indel
testcases: everything AROUND the hot path, so that the GPU kernels
can be run on the testcases end to end and checked against the `expected:` blocks the reference's
own integration tests assert (tests/lib.rs, tests/common/mod.rs:289-343).

Restated from the reference (paths relative to /root/reference/src):
  * record selection and pairing      variants/types/mod.rs:196-312, variants/sample.rs:211-216
  * evidence validity                 variants/types/deletion.rs:163-203, insertion.rs:105-122
  * candidate regions / windows       variants/evidence/realignment/mod.rs:61-162
  * haplotypes                        realignment/pairhmm.rs:213-237, deletion.rs:129-152,300-312,
                                      insertion.rs:47-74,200-222
  * allele support per record / pair  realignment/mod.rs:164-335, types/mod.rs:41-83,
                                      deletion.rs:63-96,216-256 (insert-size evidence)
  * sampling bias                     variants/sampling_bias/{reads,fragments}.rs
  * observation assembly              variants/evidence/observation.rs:229-275,379-418
  * tumor/normal scenario             cli.rs:918-940; bias sets calling.rs:398-412
BAM decoding follows the SAM/BAM specification; CigarStringView::read_pos and
Record::read_pair_orientation are recollections of rust-htslib 0.36.

The two compute stages are injected as callables, so the same glue runs on the GPU context
(`GpuBackend`) and, in the CPU test-suite, on the oracle; this module never imports the oracle.
Breakend groups (replacements, duplications, inversions, BND events) are restated in
libprosic_b200/breakends.py.  Not restated: aux-tag strand / orientation info, events across contigs.
"""
import gzip
import json
import math
import os
import struct
from collections import Counter

import numpy as np
import yaml

from . import breakends as B, obs_codec, scenario

SEQ_CODE = "=ACMGRSVTWYHKDBN"
M, I, D, N, S, H, P, EQ, X = range(9)
LN2 = math.log(2.0)


# ------------------------------------------------------------------------------- BAM records
def read_bam(path):
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", data, 4)[0]
    pos = 8 + l_text
    n_ref = struct.unpack_from("<i", data, pos)[0]
    pos += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, pos)[0]
        refs.append(data[pos + 4:pos + 4 + l_name - 1].decode())
        pos += 4 + l_name + 4
    recs = []
    while pos < len(data):
        bs = struct.unpack_from("<i", data, pos)[0]
        (ref_id, p, l_rn, mapq, _bin, n_cig, flag, l_seq, nref, npos, _tl) = struct.unpack_from("<iiBBHHHiiii", data, pos + 4)
        q = pos + 36
        name = data[q:q + l_rn - 1]
        q += l_rn
        cigar = [(v & 15, v >> 4) for v in struct.unpack_from("<%dI" % n_cig, data, q)]
        q += 4 * n_cig
        raw = data[q:q + (l_seq + 1) // 2]
        seq = "".join(SEQ_CODE[(raw[k >> 1] >> (4 if k % 2 == 0 else 0)) & 15] for k in range(l_seq)).encode()
        q += (l_seq + 1) // 2
        qual = bytes(data[q:q + l_seq])
        q += l_seq
        # aux fields: per-base strand info "SI" and read orientation "RO" of consensus reads
        # (utils/mod.rs:42-48, observation.rs:142-170); everything else is skipped
        aux, end = {}, pos + 4 + bs
        size = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
        while q + 3 <= end:
            tag, ty = data[q:q + 2].decode("latin-1"), chr(data[q + 2])
            q += 3
            if ty in size:
                q += size[ty]
            elif ty in ("Z", "H"):
                z = data.index(b"\0", q)
                if ty == "Z" and tag in ("SI", "RO"):
                    aux[tag] = bytes(data[q:z])
                q = z + 1
            elif ty == "B":
                sub, cnt = chr(data[q]), struct.unpack_from("<i", data, q + 1)[0]
                q += 5 + cnt * size[sub]
            else:
                break
        recs.append(dict(name=name, tid=ref_id, contig=refs[ref_id] if ref_id >= 0 else None, pos=p, mapq=mapq,
                         flag=flag, cigar=cigar, seq=seq, qual=qual, mtid=nref, mpos=npos,
                         si=aux.get("SI"), ro=aux.get("RO")))
        pos += 4 + bs
    return recs


def end_pos(r):
    return r["pos"] + sum(l for op, l in r["cigar"] if op in (M, D, N, EQ, X))


def _clips(r, leading, kinds):
    ops = r["cigar"] if leading else r["cigar"][::-1]
    n = 0
    for op, l in ops:
        if op in kinds:
            n += l
        elif op not in (S, H):
            break
    return n


def leading_softclips(r):
    return _clips(r, True, (S,))


def trailing_softclips(r):
    return _clips(r, False, (S,))


def is_valid_record(r):  # sample.rs:211-216
    return not (r["flag"] & (0x100 | 0x400 | 0x4 | 0x200)) and bool(r["cigar"])


def read_pos(rec, ref_pos, include_softclips, include_dels):
    """CigarStringView::read_pos (recollected, appendix A.5)."""
    rpos, qpos, j = rec["pos"], 0, 0
    cig = rec["cigar"]
    for i, (op, l) in enumerate(cig):
        if op in (M, X, EQ, I):
            j = i
            break
        if op == S:
            j = i
            if include_softclips:
                rpos = max(0, rpos - l)
            break
        if op in (D, N):
            return None
        if op in (P, H) and i == len(cig) - 1:
            return None
    while rpos <= ref_pos and j < len(cig):
        op, l = cig[j]
        inside = rpos <= ref_pos < rpos + l
        if op in (M, X, EQ) and inside:
            return qpos + ref_pos - rpos
        if op == S and include_softclips and inside:
            return qpos + ref_pos - rpos
        if op == D and include_dels and inside:
            return qpos
        if op in (M, X, EQ):
            rpos += l
            qpos += l
        elif op == S:
            qpos += l
            if include_softclips:
                rpos += l
        elif op == I:
            qpos += l
        elif op in (D, N):
            rpos += l
        elif op == H and j == len(cig) - 1:
            return None
        j += 1
    return None


def overlap(locus, r, consider_clips):
    """SingleLocus::overlap (types/mod.rs:333-361): 'enclosing', 'left', 'right', 'enclosed', None."""
    pos, end = r["pos"], end_pos(r)
    if consider_clips:
        pos = max(0, pos - leading_softclips(r))
        end += trailing_softclips(r)
    l0, l1 = locus
    if pos <= l0:
        if end >= l1:
            return "enclosing"
        if end > l0:
            return "left"
    elif end >= l1 and pos < l1:
        return "right"
    elif pos >= l0 and end <= l1:
        return "enclosed"
    return None


_RO_CODES = {b"F1R2": 0, b"F2R1": 1, b"F1F2": 2, b"F2F1": 3, b"R1R2": 4, b"R2R1": 5, b"R1F2": 6, b"R2F1": 7, b"None": 8}


def strand_or(a, b):
    """Strand |= (observation.rs:105-118): 0 Forward, 1 Reverse, 2 Both, 3 None."""
    if a == 3:
        return b
    if b == 3 or a == b:
        return a
    return 2


def strand_from_aux(si):
    """Strand::from_aux (observation.rs:72-92) over a slice of the SI tag."""
    st = 3
    for c in si:
        st = strand_or(st, {43: 0, 45: 1, 42: 2}[c])
    return st


def read_orientation(r):
    """observation.rs:142-170: the RO aux tag of consensus reads, else
    Record::read_pair_orientation (rust-htslib, recollected) -> bio-types variant index."""
    if r.get("ro") is not None:
        parts = r["ro"].split(b",")
        return 8 if len(parts) != 1 else _RO_CODES[parts[0]]
    f = r["flag"]
    if (f & 1) and not (f & 4) and not (f & 8) and r["tid"] == r["mtid"]:
        if r["pos"] == r["mpos"]:
            return 8
        rev, mrev = bool(f & 0x10), bool(f & 0x20)
        if f & 0x40:
            p1, p2, f1, f2 = r["pos"], r["mpos"], not rev, not mrev
        else:
            p1, p2, f1, f2 = r["mpos"], r["pos"], not mrev, not rev
        if p1 < p2:  # F1R2 F2R1 R1F2 R2F1 F1F2 R1R2 F2F1 R2R1 None
            return {(True, True): 4, (True, False): 0, (False, True): 2, (False, False): 5}[(f1, f2)]
        return {(True, True): 6, (True, False): 1, (False, True): 3, (False, False): 7}[(f2, f1)]
    return 8


# ------------------------------------------------------------------------------- testcases
def read_first_bcf_record(raw):
    """First record of a BCF2 file (some testcases ship `candidates.vcf` as BCF) as VCF columns
    [CHROM, POS, ID, REF, ALT, QUAL, FILTER, INFO]; integer / float / string INFO values only."""
    data = raw if raw[:3] == b"BCF" else gzip.decompress(raw)
    assert data[:5] == b"BCF\x02\x02", "not a BCF2 file"
    l_text = struct.unpack_from("<I", data, 5)[0]
    text = data[9:9 + l_text].split(b"\0")[0].decode("latin-1")
    contigs, dictionary = [], ["PASS"]
    for line in text.split("\n"):
        if line.startswith("##contig=<"):
            contigs.append(line.split("ID=")[1].split(",")[0].rstrip(">"))
        elif line.startswith(("##INFO=<", "##FILTER=<", "##FORMAT=<")):
            name = line.split("ID=")[1].split(",")[0].rstrip(">")
            idx = int(line.split("IDX=")[1].split(",")[0].rstrip(">")) if "IDX=" in line else None
            if name == "PASS" and idx in (None, 0):
                continue
            if idx is not None:
                dictionary += [None] * (idx + 1 - len(dictionary))
                dictionary[idx] = name
            elif name not in dictionary:
                dictionary.append(name)
    pos = 9 + l_text
    l_shared, _l_indiv = struct.unpack_from("<II", data, pos)
    p = pos + 8
    chrom, vpos, _rlen, _qual, n_allele_info, _n_fmt_sample = struct.unpack_from("<iiifII", data, p)
    p += 24
    n_allele, n_info = n_allele_info >> 16, n_allele_info & 0xffff

    def typed(p):
        b = data[p]
        p += 1
        n, t = b >> 4, b & 15
        if n == 15:
            (n,), p = typed(p)
        fmt = {1: "b", 2: "h", 3: "i", 5: "f", 7: "c"}.get(t)
        if t == 0 or fmt is None:
            return [], p
        size = struct.calcsize(fmt)
        if t == 7:
            return [data[p:p + n].split(b"\0")[0].decode("latin-1")], p + n
        return list(struct.unpack_from("<%d%s" % (n, fmt), data, p)), p + n * size
    ident, p = typed(p)
    alleles = []
    for _ in range(n_allele):
        a, p = typed(p)
        alleles.append(a[0] if a else "")
    _filt, p = typed(p)
    info = []
    for _ in range(n_info):
        key, p = typed(p)
        val, p = typed(p)
        name = dictionary[key[0]]
        info.append(name if not val else "%s=%s" % (name, ",".join(str(v) for v in val)))
    return [contigs[chrom], str(vpos + 1), ident[0] if ident else ".", alleles[0], ",".join(alleles[1:]), ".", ".",
            ";".join(info)]


def load_testcase(path):
    y = yaml.safe_load(open(os.path.join(path, "testcase.yaml")))
    opts = json.loads(y["options"])["Call"]["kind"]["Variants"] if "options" in y else {}
    cand = open(os.path.join(path, y.get("candidate", "candidates.vcf")), "rb").read()
    rec = None
    if cand[:3] == b"BCF":
        rec = read_first_bcf_record(cand)
    elif cand[:2] == b"\x1f\x8b":
        try:
            rec = read_first_bcf_record(cand)
        except AssertionError:   # bgzipped VCF text
            cand = gzip.decompress(cand)
    all_recs = [rec] if rec is not None else None
    if rec is None:
        all_recs = [l.split("\t") for l in cand.decode("latin-1").split("\n") if l and not l.startswith("#")]
        rec = all_recs[0]
    info = dict(kv.split("=", 1) for kv in rec[7].split(";") if "=" in kv)
    # every record of the first record's breakend event (SVTYPE=BND records share an EVENT)
    event_records = []
    if info.get("SVTYPE") == "BND" and "EVENT" in info:
        for r_ in all_recs:
            i_ = dict(kv.split("=", 1) for kv in r_[7].split(";") if "=" in kv)
            if i_.get("SVTYPE") == "BND" and i_.get("EVENT") == info["EVENT"]:
                event_records.append([int(r_[1]) - 1, r_[2], r_[3], r_[4].split(",")[0]])
    samples = {}
    for name, s in y["samples"].items():
        props = json.loads(s["properties"])
        props.setdefault("max_read_len", 100)  # version-0 testcases (tests/common/version0.rs:45-56)
        sopt = json.loads(s["options"])["Preprocess"]["kind"]["Variants"] if s.get("options") else opts
        samples[name] = dict(bam=os.path.join(path, s["path"]), props=props, opts=sopt)
    purity = y.get("purity")
    if purity is None and opts and "TumorNormal" in opts.get("mode", {}):
        purity = opts["mode"]["TumorNormal"]["purity"]
    if "seq" in y["reference"]:
        ref_seq, contig = y["reference"]["seq"].encode(), y["reference"]["name"]
    else:  # later testcase versions ship a FASTA file
        fa = open(os.path.join(path, y["reference"]["path"])).read().split(">")[1]
        contig = fa.split("\n", 1)[0].split()[0]
        ref_seq = "".join(fa.split("\n")[1:]).encode()
    scenario_yaml = None
    if y.get("scenario"):
        scenario_yaml = open(os.path.join(path, y["scenario"])).read()
    return dict(name=os.path.basename(path), ref=ref_seq, contig=contig, scenario_yaml=scenario_yaml,
                pos=int(rec[1]) - 1, ref_allele=rec[3], alt_allele=rec[4].split(",")[0], info=info,
                event_records=event_records, samples=samples, purity=purity, expected=y.get("expected", {}),
                omit={k: bool(y.get("omit_" + k, False)) for k in
                      ("strand_bias", "read_orientation_bias", "read_position_bias", "softclip_bias", "divindel_bias")})


def save_fixture(tc, path):
    """Compact, self-contained form of a testcase (reference sequence, candidate, expectations and
    the valid records of every sample) -- what the GPU box gets instead of /root/reference."""
    out = {k: v for k, v in tc.items() if k not in ("ref", "samples")}
    out["ref"] = tc["ref"].decode()
    out["samples"] = {}
    for name, s in tc["samples"].items():
        props = dict(s["props"])
        mw = int(s["opts"].get("realignment_window", s["opts"].get("indel_window", 64)))
        variant = make_variant(tc, int(mw * 1.5))
        if variant["kind"] in ("snv", "mnv"):  # SingleEndEvidence: records enclosing the locus
            recs = [r for r in read_bam(s["bam"]) if r["contig"] == tc["contig"] and is_valid_record(r) and
                    overlap(variant["locus"], r, False) == "enclosing"]
        else:
            evs = select_evidences(tc, name, variant, props)
            recs = [r for ev in evs for r in ev if r is not None]
        rank = {n: k for k, n in enumerate(sorted({r["name"] for r in recs}))}  # keeps the qname order
        recs.sort(key=lambda r: r["pos"])
        out["samples"][name] = dict(props=props, props_final=True, opts=s["opts"], records=[
            ["%05d" % rank[r["name"]], r["flag"], r["pos"], r["mapq"], r["cigar"], r["seq"].decode(),
             bytes(q + 33 for q in r["qual"]).decode(), r["tid"], r["mtid"], r["mpos"]] +
            ([r["si"].decode() if r.get("si") is not None else None, r["ro"].decode() if r.get("ro") is not None else None]
             if (r.get("si") is not None or r.get("ro") is not None) else []) for r in recs])
    with gzip.open(path, "wt") as f:
        json.dump(out, f, separators=(",", ":"))


def load_fixture(path):
    with gzip.open(path, "rt") as f:
        tc = json.load(f)
    tc["ref"] = tc["ref"].encode()
    for s in tc["samples"].values():
        s["records"] = [dict(name=r[0].encode(), flag=r[1], pos=r[2], mapq=r[3], cigar=[tuple(c) for c in r[4]],
                             seq=r[5].encode(), qual=bytes(ord(c) - 33 for c in r[6]), tid=r[7], mtid=r[8], mpos=r[9],
                             si=r[10].encode() if len(r) > 10 and r[10] is not None else None,
                             ro=r[11].encode() if len(r) > 11 and r[11] is not None else None,
                             contig=tc["contig"]) for r in s["records"]]
    return tc


def make_variant(tc, ref_window):
    """Deletion / Insertion from the candidate record (preprocessing/mod.rs:318-356)."""
    start, r, a, ref = tc["pos"], tc["ref_allele"], tc["alt_allele"], tc["ref"]
    svtype = tc["info"].get("SVTYPE")

    def bnd_variant(group, vtype, length):
        return dict(kind="bnd", vtype=vtype, len=length, group=group, loci=group.loci, fetch_loci=group.loci,
                    alts=group.alt_alleles(ref, ref_window), alt_extra=-1, locus=None)
    # utils/collect_variants.rs:93-214: SVTYPE first, then the allele shapes
    if svtype == "INV":
        length = int(tc["info"]["END"]) - 1 + 1 - start
        return bnd_variant(B.inversion_group(ref, start, length), "inv", length)
    if svtype == "DUP":
        length = int(tc["info"]["END"]) - 1 + 1 - start
        return bnd_variant(B.duplication_group(ref, start, length), "dup", length)
    if svtype == "BND":
        recs = tc.get("event_records") or []
        if not recs:
            raise NotImplementedError("breakend without EVENT")
        bnds = [B.parse_breakend(tc["contig"], p_, r_, a_, id_) for p_, id_, r_, a_ in recs]
        return bnd_variant(B.BreakendGroup([b for b in bnds if b is not None]), "bnd", 1)
    is_del = a == "<DEL>" or (len(r) > len(a) and r[:len(a)].upper() == a.upper())
    is_ins = len(r) < len(a) and a[:len(r)].upper() == r.upper()
    if svtype is None and not a.startswith("<") and len(r) != len(a) and not is_del and not is_ins:
        # neither a deletion nor an insertion (collect_variants.rs:196-214): arbitrary replacement
        return bnd_variant(B.replacement_group(ref, start, r, a), "rep", len(a))
    if not a.startswith("<") and len(r) != len(a) and (min(len(r), len(a)) != 1):
        # deletion / insertion behind a multi-base anchor: the reference keeps the record position
        # ("TODO fix position", collect_variants.rs:199) -- not met in the testcases
        raise NotImplementedError("variant type of %s: multi-base anchor %s>%s" % (tc["name"], r[:10], a[:10]))
    if a == "<DEL>" or (not a.startswith("<") and len(r) > len(a)):
        dl = (int(tc["info"]["END"]) - 1 - start) if a == "<DEL>" and "END" in tc["info"] else len(r) - len(a)
        if a == "<DEL>" and "SVLEN" in tc["info"]:
            dl = abs(int(tc["info"]["SVLEN"]))
        ro, re_ = max(0, start - ref_window), min(start + ref_window, len(ref) - dl)
        alt = bytes(ref[i] if i <= start else ref[i + dl] for i in range(ro, re_))
        center = start + int(math.floor(dl / 2.0 + 0.5))
        return dict(kind="del", len=dl, locus=(start, start + dl), centerpoint=center, alt=alt, alt_extra=0,
                    fetch_loci=[(start, start + 1), (center, center + 1), (start + dl - 1, start + dl)])
    if not a.startswith("<") and len(a) > len(r):
        ins = a[1:].encode()
        il = len(ins)
        ro, re_ = max(0, start - ref_window), min(start + il + ref_window, len(ref))
        alt = bytes(ref[i] if i <= start else (ref[i - il] if i > start + il else ins[i - start - 1])
                    for i in range(ro, re_ + il))
        return dict(kind="ins", len=il, locus=(start, start + 1), alt=alt, alt_extra=il,
                    fetch_loci=[(start, start + 1)])
    if len(r) == 1 and len(a) == 1 and not a.startswith("<"):  # SNV (types/snv.rs:35-73, 175-186)
        ro, re_ = max(0, start - ref_window), min(start + 1 + ref_window, len(ref))
        alt = bytes(ord(a.upper()) if i == start else ref[i] for i in range(ro, re_))
        return dict(kind="snv", len=1, locus=(start, start + 1), alt=alt, alt_extra=0, fetch_loci=[(start, start + 1)],
                    ref_base=r.upper(), alt_base=a.upper())
    if len(r) == len(a) and len(r) > 1 and not a.startswith("<"):  # MNV (types/mnv.rs:30-82, 215-227)
        n = len(a)
        ro, re_ = max(0, start - ref_window), min(start + n + ref_window, len(ref))
        au = a.upper().encode()
        alt = bytes(au[i - start] if start <= i < start + n else ref[i] for i in range(ro, re_))
        return dict(kind="mnv", len=n, locus=(start, start + n), alt=alt, alt_extra=0, fetch_loci=[(start, start + n)],
                    ref_base=r.upper(), alt_base=a.upper())
    raise NotImplementedError("variant type of %s %s>%s" % (tc["name"], r[:10], a[:10]))


# ------------------------------------------------------------------------------- realignment glue
def candidate_region(rec, locus, ref_len, max_window):
    """Realigner::candidate_region (realignment/mod.rs:61-156)."""
    ref_window = int(max_window * 1.5)
    max_pattern = 128
    seq_len = len(rec["seq"])
    l0, l1 = locus

    def ref_iv(bp):
        return max(0, bp - ref_window), min(bp + ref_window, ref_len)
    qs, qe = read_pos(rec, l0, True, True), read_pos(rec, l1, True, True)
    if qs is not None and qe is not None:
        mw = max(0, max_window - (qe - qs) // 2)
        ro, re_ = max(0, qs - mw), min(qe + mw, seq_len)
        exceed = max(0, (re_ - ro) - max_pattern)
        if exceed > 0:
            ro += exceed // 2
            re_ -= (exceed + 1) // 2
        return True, (ro, re_), ref_iv(l0)
    if qs is not None:
        return True, (max(0, qs - max_window), min(qs + max_window, seq_len)), ref_iv(l0)
    if qe is not None:
        return True, (max(0, qe - max_window), min(qe + max_window, seq_len)), ref_iv(l1)
    m = seq_len // 2
    enclosed = rec["pos"] >= l0 and end_pos(rec) <= l1
    return enclosed, (max(0, m - max_window), min(m + max_window - 1, seq_len)), ref_iv(rec["pos"] + m)


class RealignBatch:
    """Collects the realignment jobs of many records (Realigner::allele_support, realignment/mod.rs:
    164-335): per record the candidate regions of every locus, overlapping ones merged, one job per
    merged region = (read window, ref window, alt haplotypes); one backend call for all of them."""

    def __init__(self, tc, variant, max_window):
        self.tc, self.variant, self.max_window = tc, variant, max_window
        self.jobs, self.by_record = [], {}
        if variant["kind"] == "bnd":
            self.loci = variant["loci"]
            self.alts = [(a, -1) for a in variant["alts"]]       # not shrinkable (breakends.rs:649-664)
        else:
            self.loci = [variant["locus"]]
            self.alts = [(variant["alt"], variant["alt_extra"])]

    def add(self, rec):
        key = id(rec)
        if key in self.by_record:
            return
        regions = [candidate_region(rec, l, len(self.tc["ref"]), self.max_window) for l in self.loci]
        regions = [(q, r) for ov, q, r in regions if ov and q[1] > q[0]]
        if not regions:
            self.by_record[key] = None  # no overlap: (ln .5, ln .5, Strand::None), mod.rs:188-199
            return
        merged = []
        for (q0, q1), (r0, r1) in regions:                       # mod.rs:201-223
            if merged and r0 <= merged[-1][1][1]:
                (mq0, mq1), (mr0, _) = merged[-1]
                merged[-1] = ((min(mq0, q0), max(mq1, q1)), (mr0, r1))
            else:
                merged.append(((q0, q1), (r0, r1)))
        self.by_record[key] = []
        for (q0, q1), (r0, r1) in merged:
            self.by_record[key].append(len(self.jobs))
            self.jobs.append(dict(ref=self.tc["ref"][r0:r1], alts=self.alts,
                                  bases=rec["seq"][q0:q1], quals=rec["qual"][q0:q1], read_iv=(q0, q1)))

    def run(self, backend):
        self.results = backend.allele_support(self.jobs) if self.jobs else []

    def support(self, rec):
        """AlleleSupport of one record: dict(prob_ref, prob_alt, strand, indel_ops)."""
        js = self.by_record[id(rec)]
        if js is None:
            return dict(prob_ref=-LN2, prob_alt=-LN2, strand=3, indel_ops=())
        pr = pa = 0.0
        ops, strand = (), 3
        for j in js:
            r = self.results[j]
            if r["prob_ref"] != r["prob_alt"]:   # uninformative regions keep nothing (mod.rs:303-315)
                if rec.get("si") is not None:    # per-base strand info of consensus reads
                    q0, q1 = self.jobs[j]["read_iv"]
                    strand = strand_or(strand, strand_from_aux(rec["si"][q0:q1]))
                ops += tuple(r["indel_ops"])
            pr += r["prob_ref"]
            pa += r["prob_alt"]
        if rec.get("si") is None and pr != pa:
            strand = 1 if rec["flag"] & 0x10 else 0  # Strand::from_record (mod.rs:321-325)
        return dict(prob_ref=pr, prob_alt=pa, strand=strand, indel_ops=ops)


def merge_support(a, b):
    """AlleleSupport::merge (types/mod.rs:67-83)."""
    out = dict(prob_ref=a["prob_ref"] + b["prob_ref"], prob_alt=a["prob_alt"] + b["prob_alt"],
               strand=a["strand"], indel_ops=a["indel_ops"])
    if a["strand"] == 3:
        out["strand"], out["indel_ops"] = b["strand"], b["indel_ops"]
    elif b["strand"] != 3:
        if a["strand"] != b["strand"]:
            out["strand"] = 2
        out["indel_ops"] = a["indel_ops"] + b["indel_ops"]
    return out


# ------------------------------------------------------------------------------- sampling bias / insert size
def _phi(x):  # GSL ugaussian_P
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def _ln(x):
    return math.log(x) if x > 0 else -math.inf


def isize_pmf(value, mean, sd):  # sampling_bias/fragments.rs:175-178
    return _ln(_phi((value + 0.5 - mean) / sd) - _phi((value - 0.5 - mean) / sd))


def ln_one_minus_exp(p):
    if p == -math.inf:
        return 0.0
    if p >= 0.0:
        return -math.inf
    return math.log1p(-math.exp(p)) if p < -0.693 else math.log(-math.expm1(p))


def ln_sum_exp(ps):
    ps = [p for p in ps]
    if not ps:
        return -math.inf
    mx = max(ps)
    if mx == -math.inf:
        return -math.inf
    k = ps.index(mx)
    return mx + math.log1p(sum(math.exp(p - mx) for i, p in enumerate(ps) if i != k and p != -math.inf))


def feasible_bases(variant, read_len, props):  # deletion.rs:104-118, insertion.rs:78-92, breakends.rs:352-380
    if variant["kind"] == "bnd":
        return variant["group"].feasible_bases(read_len, props)
    maxlen = props["max_del_cigar_len"] if variant["kind"] == "del" else props["max_ins_cigar_len"]
    if maxlen is not None and variant["len"] <= maxlen:
        return read_len
    if props["frac_max_softclip"] is not None:
        return int(read_len * props["frac_max_softclip"])
    return None


def prob_sample_alt_read(variant, read_len, props):  # sampling_bias/reads.rs:21-40
    feasible = feasible_bases(variant, read_len, props)
    if feasible is None:
        return 0.0
    if variant["kind"] == "bnd":   # enclosable_len(): None unless the group is a plain deletion / insertion
        el = variant["group"].enclosable_len()
        n_alt = read_len if el is None else min(el, read_len)
    else:
        n_alt = min(variant["len"], read_len)
    return _ln(min(n_alt, feasible)) - math.log(n_alt)


def prob_sample_alt_fragment(variant, ll, rl, props):  # sampling_bias/fragments.rs:44-151
    lf, rf = feasible_bases(variant, ll, props), feasible_bases(variant, rl, props)
    if lf is None or rf is None:
        return -math.inf
    delta_ref, delta_alt = variant["len"], 0
    infeasible_read_pos = max(0, ll - lf) + max(0, rl - rf)  # enclose_only = true
    mean, sd = props["insert_size"]["mean"], props["insert_size"]["sd"]
    m = int(math.floor(mean + 0.5))
    s = int(math.ceil(sd)) * 6
    terms = []
    for x in range(max(0, m - s), m + s):
        internal = max(0, max(0, x - ll) - rl)
        infeasible_pos_alt = infeasible_read_pos + max(0, internal + 1 - delta_alt)
        infeasible_pos_ref = max(0, internal + 1 - delta_ref)
        valid_alt = max(0, max(0, x - delta_alt) - infeasible_pos_alt)
        valid_ref = max(0, x - infeasible_pos_ref)
        if x <= delta_alt or x <= delta_alt + infeasible_pos_alt or x <= infeasible_pos_ref:
            continue
        terms.append(isize_pmf(x, mean, sd) + _ln(valid_alt) - math.log(valid_ref))
    p = ln_sum_exp(terms)
    return 0.0 if 0.0 < p <= 1e-3 else p  # cap_numerical_overshoot


def estimate_insert_size(left, right):  # variants/evidence/insert_size.rs:17-45
    def aln(r):
        return (max(0, r["pos"] - _clips(r, True, (S, H))), end_pos(r) + _clips(r, False, (S, H)))
    (ls, le), (rs, re_) = aln(left), aln(right)
    return (rs - le) + (le - ls) + (re_ - rs)


# ------------------------------------------------------------------------------- subsampling
class Subsampler:
    """SubsampleCandidates (sample.rs:132-160): `StdRng::seed_from_u64(48074578)` of rand 0.7.3
    (Cargo.lock) = ChaCha20 (rand_chacha 0.2: 64-bit block counter, stream 0) keyed by the PCG32
    expansion of the seed (rand_core 0.5 SeedableRng::seed_from_u64); keep() draws
    Uniform::new(0.0, 1.0): (next_u64 >> 12) as the mantissa of a float in [1, 2), minus 1."""

    def __init__(self, max_depth, depth):
        self.needed = depth > max_depth
        self.prob = max_depth / depth if depth else 1.0
        if self.needed:
            state, key = 48074578, []
            for _ in range(8):
                state = (state * 6364136223846793005 + 11634580027462260723) & 0xFFFFFFFFFFFFFFFF
                x = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
                rot = state >> 59
                key.append(((x >> rot) | (x << ((32 - rot) & 31))) & 0xFFFFFFFF)
            self.key, self.counter, self.buf = key, 0, []

    def _block(self):
        c = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + self.key + \
            [self.counter & 0xFFFFFFFF, self.counter >> 32, 0, 0]
        x = list(c)

        def qr(a, b, cc, d):
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & 0xFFFFFFFF
            x[cc] = (x[cc] + x[d]) & 0xFFFFFFFF; x[b] ^= x[cc]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & 0xFFFFFFFF
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & 0xFFFFFFFF
            x[cc] = (x[cc] + x[d]) & 0xFFFFFFFF; x[b] ^= x[cc]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & 0xFFFFFFFF
        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        self.counter += 1
        return [(a + b) & 0xFFFFFFFF for a, b in zip(x, c)]

    def keep(self):
        if not self.needed:
            return True
        if len(self.buf) < 2:
            self.buf += self._block()
        lo, hi = self.buf[0], self.buf[1]
        del self.buf[:2]
        u = ((hi << 32) | lo) >> 12
        value = struct.unpack("<d", struct.pack("<Q", (1023 << 52) | u))[0] - 1.0
        return value <= self.prob


# ------------------------------------------------------------------------------- observations
def select_evidences(tc, sample_name, variant, props):
    """Record fetching, pairing and validity (types/mod.rs:196-312): list of (left, right|None).
    `props` (alignment properties) is updated as update_max_cigar_ops_len does, unless the fixture
    already holds the updated values."""
    smp = tc["samples"][sample_name]
    records = smp.get("records")
    if records is None:
        records = [r for r in read_bam(smp["bam"]) if r["contig"] == tc["contig"]]
    # ---- fetch windows around the loci (sample.rs:40-100): read-pair window on both sides, merged
    w = int(props["insert_size"]["mean"] + props["insert_size"]["sd"] * 6.0) if props["insert_size"] else props["max_read_len"]
    fetches = []
    for l0, l1 in variant["fetch_loci"]:
        if fetches and max(0, l0 - w) <= fetches[-1][1] + w:
            fetches[-1][1] = l1
        else:
            fetches.append([l0, l1])
    cands = {}
    for f0, f1 in fetches:
        lo, hi = max(0, f0 - w), f1 + w
        for r in records:
            if not is_valid_record(r) or r["pos"] >= hi or end_pos(r) < lo:
                continue
            if not smp.get("props_final"):
                for op, l in r["cigar"]:  # AlignmentProperties::update_max_cigar_ops_len
                    if op == S and props["frac_max_softclip"] is not None:
                        props["frac_max_softclip"] = max(props["frac_max_softclip"], l / len(r["seq"]))
                    elif op == D and props["max_del_cigar_len"] is not None:
                        props["max_del_cigar_len"] = max(props["max_del_cigar_len"], l)
                    elif op == I and props["max_ins_cigar_len"] is not None:
                        props["max_ins_cigar_len"] = max(props["max_ins_cigar_len"], l)
            c = cands.get(r["name"])
            if c is None:
                cands[r["name"]] = [r, None]
            elif c[0] is not r and c[1] is not r:
                c[1] = r
    evidences = []
    for name in sorted(cands):  # BTreeMap order
        left, right = cands[name]
        if right is not None:
            if left["mapq"] == 0 or right["mapq"] == 0:
                continue
            ev = (left, right)
        else:
            ev = (left, None)
        if _is_valid_evidence(variant, ev, props):
            evidences.append(ev)
    return evidences


def extract_observations(tc, sample_name, backend, max_window=64, trace=None):
    """Observable<PairedEndEvidence>::extract_observations + Sample::extract_observations for the
    testcase's candidate in one sample.  Returns (cols dict, cat) of the processed pileup,
    quantised through the observation wire format as `call` would read it."""
    smp = tc["samples"][sample_name]
    props = dict(smp["props"])
    max_window = int(smp["opts"].get("realignment_window", smp["opts"].get("indel_window", max_window)))
    variant = make_variant(tc, int(max_window * 1.5))
    evidences = select_evidences(tc, sample_name, variant, props)
    # METHOD (types/mod.rs:289-303): if ALL loci exceed the maximum depth, the evidences are subsampled
    max_depth = int(smp["opts"].get("max_depth", 200))
    if variant["kind"] == "bnd":
        depth = [0] * len(variant["loci"])
        for ev in evidences:
            for i in variant["group"].is_valid_evidence(ev, overlap, end_pos):
                depth[i] += 1
    else:
        depth = [len(evidences)]
    if evidences and all(d > max_depth for d in depth):
        sub = Subsampler(max_depth, len(evidences))
        evidences = [ev for ev in evidences if sub.keep()]
    report_ops = variant["kind"] in ("del", "ins") or variant.get("vtype") == "rep"
    batch = RealignBatch(tc, variant, max_window)
    for left, right in evidences:
        batch.add(left)
        if right is not None:
            batch.add(right)
    batch.run(backend)
    raw = []
    for left, right in evidences:
        sup = batch.support(left)
        if right is not None:
            sup = merge_support(sup, batch.support(right))
            if variant["kind"] == "del" and props["insert_size"] is not None:
                sup = merge_support(sup, _isize_support(variant, left, right, props))
        if trace is not None:  # per evidence: what the realigner said and whether the observation is kept
            trace.append(dict(name=left["name"], kept=sup["strand"] != 3, strand=sup["strand"],
                              indel_ops=sup["indel_ops"], prob_ref=sup["prob_ref"], prob_alt=sup["prob_alt"]))
        if sup["strand"] == 3:
            continue  # evidence_to_observation: unstranded support is dropped (observation.rs:384-388)
        mis = [r["mapq"] * (-math.log(10.0) / 10.0) for r in (left, right) if r is not None]
        if right is None:
            psa = prob_sample_alt_read(variant, len(left["seq"]), props)
            ln_len = len(left["seq"])
            softclipped = leading_softclips(left) > 0 or trailing_softclips(left) > 0
        else:
            ll, rl = len(left["seq"]), len(right["seq"])
            if variant["kind"] == "del" and props["insert_size"] is not None:
                psa = prob_sample_alt_fragment(variant, ll, rl, props)
            else:
                psa = ln_one_minus_exp(ln_one_minus_exp(prob_sample_alt_read(variant, ll, props)) +
                                       ln_one_minus_exp(prob_sample_alt_read(variant, rl, props)))
            ln_len = ll + rl
            softclipped = any(leading_softclips(r) > 0 or trailing_softclips(r) > 0 for r in (left, right))
        raw.append(dict(prob_mapping=ln_one_minus_exp(max(mis)), prob_alt=sup["prob_alt"], prob_ref=sup["prob_ref"],
                        prob_missed_allele=ln_sum_exp([sup["prob_ref"], sup["prob_alt"]]) - LN2,
                        prob_sample_alt=psa, prob_double_overlap=0.0 if sup["strand"] == 2 else -math.inf,
                        prob_hit_base=-math.log(ln_len), strand=sup["strand"], orientation=read_orientation(left),
                        softclipped=softclipped, paired=bool(left["flag"] & 1),
                        # Variant::report_indel_operations (deletion.rs:158, insertion.rs:105, replacement.rs:81)
                        indel_ops=sup["indel_ops"] if report_ops else ()))
    # ---- Observation::process: major indel operations (observation.rs:229-275, 339-359)
    counter = Counter(o["indel_ops"] for o in raw if o["indel_ops"])
    major = counter.most_common(1)[0][0] if counter else None
    n = len(raw)
    cols = {c: np.array([o[c] for o in raw], np.float64) for c in
            ("prob_mapping", "prob_alt", "prob_ref", "prob_missed_allele", "prob_sample_alt", "prob_double_overlap",
             "prob_hit_base")}
    cat = np.zeros(n, np.uint16)
    for k, o in enumerate(raw):
        indel = 2 if not o["indel_ops"] else (0 if o["indel_ops"] == major else 1)
        cat[k] = (o["strand"] | (o["orientation"] << 2) | (1 << 6) | (int(o["softclipped"]) << 7) | (indel << 8) |
                  (int(o["paired"]) << 10))  # read_position: None -> ReadPosition::Some
    # ---- through the wire format, as `call` reads what `preprocess` wrote
    return obs_codec.decode_record(obs_codec.encode_record(cols, cat))


def prob_read_base(read_base, ref_base, q):
    """variants/evidence/bases.rs:13-21 with the LUTs of 33-48."""
    ln_eps = q * (-math.log(10.0) / 10.0)
    if chr(read_base).upper() == ref_base.upper():
        return ln_one_minus_exp(ln_eps)
    return ln_eps + math.log(0.3333)


def extract_observations_snv(tc, sample_name, backend, omit_read_orientation_bias=False, max_window=64):
    """Observable<SingleEndEvidence>::extract_observations (types/mod.rs:132-183) for an SNV
    (types/snv.rs:75-165), Observation::process and the call-time removal of non-standard
    alignments (calling.rs:423-441, observation.rs:303-322)."""
    smp = tc["samples"][sample_name]
    props = dict(smp["props"])
    max_window = int(smp["opts"].get("realignment_window", smp["opts"].get("indel_window", max_window)))
    variant = make_variant(tc, int(max_window * 1.5))
    start = variant["locus"][0]
    records = smp.get("records")
    if records is None:
        records = [r for r in read_bam(smp["bam"]) if r["contig"] == tc["contig"]]
    lo, hi = max(0, start - props["max_read_len"]), variant["locus"][1]  # RecordBuffer::fetch(locus, false)
    evidences = []
    for r in records:
        if not is_valid_record(r) or r["pos"] >= hi or end_pos(r) < lo:
            continue
        if overlap(variant["locus"], r, False) == "enclosing":
            evidences.append(r)
    sub = Subsampler(int(smp["opts"].get("max_depth", 200)), len(evidences))  # types/mod.rs:166-172
    evidences = [r for r in evidences if sub.keep()]
    batch = RealignBatch(tc, variant, max_window)
    for r in evidences:
        if any(op in (I, D) for op, _ in r["cigar"]):
            batch.add(r)
    batch.run(backend)
    raw = []
    for r in evidences:
        if any(op in (I, D) for op, _ in r["cigar"]):
            sup = batch.support(r)
            qpos = None
        else:
            # one base (snv.rs:109-151) or every base of the MNV (mnv.rs:116-197); the read position is
            # the first base's
            qpos, p_ref, p_alt, gone, st_aux = None, 0.0, 0.0, False, 3
            for k in range(variant["len"]):
                qp = read_pos(r, start + k, False, False)
                if qp is None:   # position deleted / skipped in this read: no observation
                    gone = True
                    break
                if qpos is None:
                    qpos = qp
                rb, q = r["seq"][qp], r["qual"][qp]
                ab, fb = variant["alt_base"][k], variant["ref_base"][k]
                non_alt = chr(rb) if chr(rb).upper() != ab else fb
                b_alt, b_ref = prob_read_base(rb, ab, q), prob_read_base(rb, non_alt, q)
                if b_alt != b_ref and r.get("si") is not None:   # per-base strand of consensus reads (SI tag)
                    st_aux = strand_or(st_aux, strand_from_aux(r["si"][qp:qp + 1]))
                p_alt += b_alt
                p_ref += b_ref
            if gone:
                continue
            if r.get("si") is not None:   # snv.rs:136-143 / mnv.rs:160-183
                strand = st_aux
            else:
                strand = (1 if r["flag"] & 0x10 else 0) if p_ref != p_alt else 3
            sup = dict(prob_ref=p_ref, prob_alt=p_alt, strand=strand)
        if sup["strand"] == 3:
            continue
        raw.append(dict(prob_mapping=ln_one_minus_exp(r["mapq"] * (-math.log(10.0) / 10.0)), prob_alt=sup["prob_alt"],
                        prob_ref=sup["prob_ref"], prob_missed_allele=ln_sum_exp([sup["prob_ref"], sup["prob_alt"]]) - LN2,
                        prob_sample_alt=0.0, prob_double_overlap=0.0 if sup["strand"] == 2 else -math.inf,
                        prob_hit_base=-math.log(len(r["seq"])),
                        strand=sup["strand"], orientation=read_orientation(r),
                        softclipped=leading_softclips(r) > 0 or trailing_softclips(r) > 0, paired=bool(r["flag"] & 1),
                        read_position=qpos))
    counter = Counter(o["read_position"] for o in raw if o["read_position"] is not None)
    major = counter.most_common(1)[0][0] if counter else None
    if not omit_read_orientation_bias:
        raw = [o for o in raw if o["orientation"] in (0, 1, 8)]
    cols = {c: np.array([o[c] for o in raw], np.float64) for c in
            ("prob_mapping", "prob_alt", "prob_ref", "prob_missed_allele", "prob_sample_alt", "prob_double_overlap",
             "prob_hit_base")}
    cat = np.zeros(len(raw), np.uint16)
    for k, o in enumerate(raw):
        rp = 0 if (o["read_position"] is not None and o["read_position"] == major) else 1
        cat[k] = (o["strand"] | (o["orientation"] << 2) | (rp << 6) | (int(o["softclipped"]) << 7) | (2 << 8) |
                  (int(o["paired"]) << 10))  # SNVs do not report indel operations: IndelOperations::None
    return obs_codec.decode_record(obs_codec.encode_record(cols, cat))


def single_sample_scenario(tc, events, resolution=100, prior=None):
    """One-sample Generic scenario, hand-flattened for the testcase's variant (grammar formulas with
    variant-type atoms such as `C>T | G>A` are resolved against the candidate by the host).
    events: list of (name, node spec) with node spec ('range', lo, hi, left_excl, right_excl),
    ('set', [afs]) or None (formula false for this variant).  Bias sets: calling.rs:398-412."""
    b = scenario._Builder()
    snv = make_variant(tc, 96)["kind"] == "snv"
    om = tc["omit"]
    arts = scenario.artifact_biases(not om["strand_bias"], snv and not om["read_orientation_bias"],
                                    snv and not om["read_position_bias"], snv and not om["softclip_bias"],
                                    not om["divindel_bias"])
    named = []
    for name, spec in events:
        if spec is None:
            root = b.node(scenario.NODE_FALSE)
        elif spec[0] == "range":
            root = b.chain([(0, scenario.NODE_RANGE, [spec[1], spec[2], float(spec[3]), float(spec[4])])])
        else:
            root = b.chain([(0, scenario.NODE_SET, list(spec[1]))])
        named.append((name, [root]))
    ev, biases = scenario._with_events(b, named, 1, arts)
    scn = dict(n_samples=1, resolution=[resolution], contaminated=[0], by=[0], purity=[1.0], nodes=b.nodes,
               events=ev, biases=biases)
    if prior is not None:
        scn["prior_fn"] = prior
    return scn


def scenario_from_yaml(tc, text):
    """Scenario of a testcase derived from its scenario.yaml by the grammar front end
    (libprosic_b200/grammar.py) for the testcase's contig, candidate and omit_* flags."""
    from . import grammar
    variant = make_variant(tc, 96)
    snv = variant["kind"] == "snv"
    om = tc["omit"]
    g = grammar.Scenario(text)
    return grammar.flatten(g, contig=tc["contig"],
                           snv=(tc["ref_allele"].upper(), tc["alt_allele"].upper()) if snv else None,
                           snv_or_mnv=variant["kind"] in ("snv", "mnv"), omit_strand_bias=om["strand_bias"],
                           omit_read_orientation_bias=om["read_orientation_bias"],
                           omit_read_position_bias=om["read_position_bias"],
                           omit_softclip_bias=om["softclip_bias"], omit_divindel_bias=om["divindel_bias"])


def run_single_sample_testcase(path, backend, sample, events, prior=None, scenario_yaml=None):
    """preprocess + call of a one-sample Generic testcase; same result layout as run_testcase."""
    tc = load_fixture(path) if os.path.isfile(path) else load_testcase(path)
    variant = make_variant(tc, 96)
    if variant["kind"] == "snv":
        cols, cat = extract_observations_snv(tc, sample, backend, tc["omit"]["read_orientation_bias"])
    else:
        cols, cat = extract_observations(tc, sample, backend)
    off = np.array([0, len(cat)], np.int64)
    scn = single_sample_scenario(tc, events, prior=prior)
    if scenario_yaml is not None:
        scn = scenario_from_yaml(tc, scenario_yaml)
        if prior is not None:
            scn["prior_fn"] = prior
    if prior is not None:  # Prior::compute on the leaves the library reports (set-only universes)
        _, afs = backend.model_leaves(scn)
        scn["prior_ln"] = np.array([prior(list(a)) for a in afs])
    post = backend.posterior(scn, cols, cat, off)
    rec = backend.call_records(scn, post, cols["prob_mapping"], off)
    out = {"PROB_" + n.upper(): float(v) for n, v in zip(rec["names"], rec["prob"][0])}
    out["af"] = {sample: float(rec["af"][0][0])}
    out["n_obs"] = {sample: int(off[1])}
    out["expected"] = tc["expected"]
    return out


def _range(lo, hi, left_excl, right_excl):
    return ("range", lo, hi, left_excl, right_excl)


# One-sample Generic testcases of the reference, scenarios flattened by hand for the testcase's
# candidate (scenario.yaml next to each testcase; `(C>T | G>A)`-style atoms resolved against REF>ALT):
#   name -> (sample, [(event, VAF spec | None)], heterozygosity of the species prior | None)
GENERIC_CASES = {
    # present: "tumor:]0.0,1.0]"
    "test41": ("tumor", [("present", _range(0.0, 1.0, True, False))], None),
    # T>C matches "(T>C | G>A)": ffpe_artifact ]0.0,0.1[, present ]0.1,1.0]
    "test40": ("tumor", [("ffpe_artifact", _range(0.0, 0.1, True, True)), ("present", _range(0.1, 1.0, True, False))], None),
    # G>T does not match "(C>T | G>A)": ffpe_artifact is false, present ]0.0,1.0]
    "test57": ("tumor", [("ffpe_artifact", None), ("present", _range(0.0, 1.0, True, False))], None),
    # a deletion does not match either
    "test59": ("tumor", [("ffpe_artifact", None), ("present", _range(0.0, 1.0, True, False))], None),
    # present: "NA12878:0.5 | NA12878:1.0", species heterozygosity 0.001, ploidy 2
    "test_giab_03": ("NA12878", [("present", ("set", [0.5, 1.0]))], 0.001),
}


def run_generic_case(path, backend, scenario_yaml=None):
    name = os.path.basename(path).split(".")[0]
    sample, events, het = GENERIC_CASES[name]
    return run_single_sample_testcase(path, backend, sample, events, germline_prior(het) if het else None,
                                      scenario_yaml=scenario_yaml)


# ------------------------------------------------------------------------------- pedigree (trio) testcases
def _hypergeom_pmf(N, K, n, k):
    """statrs Hypergeometric::new(N, K, n).pmf(k)."""
    if k > K or n - k > N - K or k > n:
        return 0.0
    return math.comb(K, k) * math.comb(N - K, n - k) / math.comb(N, n)


def mendelian_trio_prior(heterozygosity, germline_mutation_rate, child, parents, ploidy=2):
    """Prior::calc_prob for a diploid trio without somatic terms (variants/model/prior.rs:288-431):
    population prior of the parents (prob_population_germline, 628-656) and Mendelian inheritance of
    the child (prob_mendelian_alt_counts / prob_mendelian_inheritance, 674-786).  child / parents are
    sample indices in the scenario's (alphabetical) order."""
    def f(afs):
        n_alt = [int(round(ploidy * a)) for a in afs]
        m = sum(n_alt[p] for p in parents)
        if m > 0:
            prob = math.log(heterozygosity) - math.log(m)
        else:
            n = ploidy * len(parents)
            prob = ln_one_minus_exp(ln_sum_exp([math.log(heterozygosity) - math.log(k) for k in range(1, n + 1)]))
        half = ploidy // 2  # even ploidies: one meiotic split case
        terms = []
        for a1 in range(0, min(n_alt[parents[0]], half) + 1):
            for a2 in range(0, min(n_alt[parents[1]], half) + 1):
                if a1 + a2 > n_alt[child]:
                    continue
                p = (_ln(_hypergeom_pmf(ploidy, n_alt[parents[0]], half, a1)) +
                     _ln(_hypergeom_pmf(ploidy, n_alt[parents[1]], half, a2)))
                terms.append(p + math.log(germline_mutation_rate) * (n_alt[child] - a1 - a2))
        return prob + ln_sum_exp(terms)
    return f


def trio_scenario(tc):
    """test61-63 (scenario.yaml): samples father = 0, index = 1, mother = 2 (alphabetical, quirk Q12),
    universe {0, .5, 1} each (ploidy 2);
      denovo:          "(index:0.5 | index:1.0) & father:0.0 & mother:0.0"
      not_interesting: "!father:0.0 | !mother:0.0"   (disjoint cover: father != 0, or father = 0 & mother != 0)"""
    b = scenario._Builder()
    S, gt = scenario.NODE_SET, [0.0, 0.5, 1.0]
    denovo = b.chain([(0, S, [0.0]), (1, S, [0.5, 1.0]), (2, S, [0.0])])
    ni_a = b.chain([(0, S, [0.5, 1.0]), (1, S, gt), (2, S, gt)])
    ni_b = b.chain([(0, S, [0.0]), (1, S, gt), (2, S, [0.5, 1.0])])
    om = tc["omit"]
    arts = scenario.artifact_biases(not om["strand_bias"], not om["read_orientation_bias"],
                                    not om["read_position_bias"], not om["softclip_bias"], not om["divindel_bias"])
    ev, biases = scenario._with_events(b, [("denovo", [denovo]), ("not_interesting", [ni_a, ni_b])], 3, arts)
    return dict(n_samples=3, resolution=[100, 100, 100], contaminated=[0, 0, 0], by=[0, 0, 0], purity=[1.0] * 3,
                nodes=b.nodes, events=ev, biases=biases,
                prior_fn=mendelian_trio_prior(0.001, 0.1, child=1, parents=(2, 0)))


def run_trio_testcase(path, backend, scenario_yaml=None):
    """preprocess (father, index, mother) + call for the pedigree SNV testcases test61-63."""
    tc = load_fixture(path) if os.path.isfile(path) else load_testcase(path)
    order = ["father", "index", "mother"]
    pile = [extract_observations_snv(tc, s_, backend, tc["omit"]["read_orientation_bias"]) for s_ in order]
    cols = {c: np.concatenate([p_[0][c] for p_ in pile]) for c in obs_codec.COLUMNS}
    cat = np.concatenate([p_[1] for p_ in pile])
    off = np.concatenate([[0], np.cumsum([len(p_[1]) for p_ in pile])]).astype(np.int64)
    scn = trio_scenario(tc)
    if scenario_yaml is not None:
        prior_fn = scn["prior_fn"]
        scn = scenario_from_yaml(tc, scenario_yaml)
        assert scn["sample_names"] == order
        scn["prior_fn"] = prior_fn
    _, afs = backend.model_leaves(scn)
    scn["prior_ln"] = np.array([scn["prior_fn"](list(a)) for a in afs])
    post = backend.posterior(scn, cols, cat, off)
    rec = backend.call_records(scn, post, cols["prob_mapping"], off)
    out = {"PROB_" + n.upper(): float(v) for n, v in zip(rec["names"], rec["prob"][0])}
    out["af"] = {s_: float(rec["af"][0][k]) for k, s_ in enumerate(order)}
    out["n_obs"] = {s_: int(off[k + 1] - off[k]) for k, s_ in enumerate(order)}
    out["expected"] = tc["expected"]
    return out


TRIO_CASES = ("test61", "test62", "test63")


def germline_prior(heterozygosity, ploidy=2):
    """Prior::prob_population_germline for one diploid sample without inheritance or somatic terms
    (variants/model/prior.rs:288-431, 628-656): m alt alleles -> ln(het) - ln(m); none ->
    ln(1 - sum_k het/k)."""
    def f(afs):
        m = int(round(ploidy * afs[0]))
        if m > 0:
            return math.log(heterozygosity) - math.log(m)
        return ln_one_minus_exp(ln_sum_exp([math.log(heterozygosity) - math.log(k) for k in range(1, ploidy + 1)]))
    return f


def _is_valid_evidence(variant, ev, props):
    left, right = ev
    if variant["kind"] == "bnd":
        return variant["group"].is_valid_evidence(ev, overlap, end_pos) is not None
    loc = variant["locus"]
    if right is None:
        return overlap(loc, left, True) is not None
    any_ov = overlap(loc, left, True) is not None or overlap(loc, right, True) is not None
    if variant["kind"] == "del" and props["insert_size"] is not None:
        return left["pos"] < variant["centerpoint"] < end_pos(right) and any_ov
    return any_ov


def _isize_support(variant, left, right, props):  # deletion.rs:63-96
    isz = estimate_insert_size(left, right)
    mean, sd = props["insert_size"]["mean"], props["insert_size"]["sd"]
    p_ref, p_alt = isize_pmf(isz, mean, sd), isize_pmf(isz, mean + variant["len"], sd)

    def within(shift):
        return abs(isz - (mean + shift)) <= sd
    if (p_ref == -math.inf and not within(variant["len"])) or (p_alt == -math.inf and not within(0.0)):
        return dict(prob_ref=0.0, prob_alt=0.0, strand=3, indel_ops=())
    return dict(prob_ref=p_ref, prob_alt=p_alt, strand=3, indel_ops=())


# ------------------------------------------------------------------------------- calling + expectations
def tn_scenario(tc):
    """Built-in tumor/normal scenario (cli.rs:918-940) with the indel bias set of calling.rs:398-412."""
    scn = scenario.tumor_normal(purity=tc["purity"], indel=True)
    arts = scenario.artifact_biases(not tc["omit"]["strand_bias"], False, False, False, not tc["omit"]["divindel_bias"])
    b = scenario._Builder()
    b.nodes = scn["nodes"]
    named, seen = [], set()
    for ev in scn["events"]:
        if ev["name"] != "absent" and ev["name"] not in seen:
            seen.add(ev["name"])
            named.append((ev["name"], ev["roots"]))
    absent = [ev for ev in scn["events"] if ev["name"] == "absent"][0]
    events, biases = [dict(name="absent", roots=absent["roots"], biases=[0])], [scenario.bias()] + arts
    for name, roots in named:
        events.append(dict(name=name, roots=roots, biases=[0]))
        if arts:
            events.append(dict(name=name, roots=roots, biases=list(range(1, len(biases)))))
    return dict(scn, events=events, biases=biases)


def run_testcase(path, backend):
    """preprocess (both samples) + call: dict(PROB_*: PHRED f32, af: {sample: f32}, n_obs).
    path: a reference testcase directory or a compact fixture written by save_fixture."""
    tc = load_fixture(path) if os.path.isfile(path) else load_testcase(path)
    pile = {s: extract_observations(tc, s, backend) for s in ("normal", "tumor")}  # alphabetical (quirk Q12)
    names = obs_codec.COLUMNS
    cols = {c: np.concatenate([pile["normal"][0][c], pile["tumor"][0][c]]) for c in names}
    cat = np.concatenate([pile["normal"][1], pile["tumor"][1]])
    off = np.array([0, len(pile["normal"][1]), len(cat)], np.int64)
    scn = tn_scenario(tc)
    post = backend.posterior(scn, cols, cat, off)
    rec = backend.call_records(scn, post, cols["prob_mapping"], off)
    out = {"PROB_" + n.upper(): float(v) for n, v in zip(rec["names"], rec["prob"][0])}
    out["af"] = {"normal": float(rec["af"][0][0]), "tumor": float(rec["af"][0][1])}
    out["n_obs"] = {"normal": int(off[1]), "tumor": int(off[2] - off[1])}
    out["expected"] = tc["expected"]
    return out


def run_scenario_testcase(path, backend, scenario_yaml=None):
    """preprocess + call of a Generic-mode testcase driven by its scenario.yaml through the grammar
    front end (any number of samples; samples of the scenario without a BAM file contribute an
    empty pileup, calling.rs:442-444).  Samples without a `universe` get the species prior of
    libprosic_b200/prior.py (heterozygosity, Mendelian inheritance; no somatic / clonal terms)."""
    from . import grammar
    tc = load_fixture(path) if os.path.isfile(path) else load_testcase(path)
    text = scenario_yaml or tc.get("scenario_yaml")
    if text is None:
        raise NotImplementedError("testcase without a scenario")
    from . import prior as prior_mod
    g = grammar.Scenario(text)
    variant = make_variant(tc, 96)
    scn = scenario_from_yaml(tc, text)
    order = scn["sample_names"]
    if any(smp.universe is None for smp in g.samples.values()):
        # Prior::compute on the host (libprosic_b200/prior.py): callback for the oracle, table over the
        # leaves the library reports for the GPU
        # VariantTypeFraction::get (grammar/mod.rs:399-409): replacements count as indels, the other
        # breakend groups as structural variants
        pk = variant["kind"] if variant["kind"] != "bnd" else ("del" if variant["vtype"] == "rep" else "sv")
        scn["prior_fn"], prog = prior_mod.make_prior(g, tc["contig"], pk, want_program=True)
        leaves = backend.model_leaves(scn)
        if leaves is None:
            # a tree path binds a Range: its grid points depend on the site's depth, no table over
            # the leaves exists -- the device evaluates the prior program at every leaf
            scn["prior_prog"] = prog
        else:
            scn["prior_ln"] = np.array([scn["prior_fn"](list(a)) for a in leaves[1]])
    pile = []
    for name in order:
        if name not in tc["samples"]:
            pile.append(({c: np.zeros(0) for c in obs_codec.COLUMNS}, np.zeros(0, np.uint16)))
        elif variant["kind"] in ("snv", "mnv"):
            pile.append(extract_observations_snv(tc, name, backend, tc["omit"]["read_orientation_bias"]))
        else:
            pile.append(extract_observations(tc, name, backend))
    cols = {c: np.concatenate([p_[0][c] for p_ in pile]) for c in obs_codec.COLUMNS}
    cat = np.concatenate([p_[1] for p_ in pile]).astype(np.uint16)
    off = np.concatenate([[0], np.cumsum([len(p_[1]) for p_ in pile])]).astype(np.int64)
    post = backend.posterior(scn, cols, cat, off)
    rec = backend.call_records(scn, post, cols["prob_mapping"], off)
    out = {"PROB_" + n.upper(): float(v) for n, v in zip(rec["names"], rec["prob"][0])}
    out["af"] = {s_: float(rec["af"][0][k]) for k, s_ in enumerate(order)}
    out["n_obs"] = {s_: int(off[k + 1] - off[k]) for k, s_ in enumerate(order)}
    out["expected"] = tc["expected"]
    return out


# Generic testcases of the reference with a scenario.yaml and a uniform prior that the suites run from
# compact fixtures (all that run in the build container: tests/golden/scenario_results.txt)
SCENARIO_CASES = ("omit_sb", "test34", "test37", "test40", "test49", "test52", "test58", "test57", "test70",
                  "test_low_cov_vaf", "test_panel_overlap", "test_contig_universe", "test_giab_02",
                  # breakend groups (libprosic_b200/breakends.py): inversion, duplication, a BND event of four
                  # records, replacements (PCR homopolymer errors called as DivIndel artifacts)
                  "test43", "test44", "test45", "test51", "test60", "test_giab_05", "test_pcr_homopolymer_error1",
                  "test_pcr_homopolymer_error2", "test_pcr_homopolymer_error3",
                  # deep amplicon / consensus-read data: subsampling above max_depth (Subsampler) and the
                  # per-base strand (SI) / read orientation (RO) aux tags
                  "test55", "test_panel_unknown_orientation_bias", "pattern_too_long")
# tumor/normal testcases whose candidate is a breakend event (an insertion written as two BND records)
TN_BND_CASES = ("test42", "test47")


# ... and with a prior over continuous allele frequencies (somatic mutation rates, Mendelian and clonal
# inheritance): the device evaluates the prior as a program (vlr_prior_prog) at every leaf
SCENARIO_CASES_PRIOR = ("test64", "test69", "test71")


def check_expectations(result):
    """tests/common/mod.rs:289-343: list of (expression, passed)."""
    env = dict(result["af"])
    env.update({k: v for k, v in result.items() if k.startswith("PROB_")})
    env["inf"] = math.inf
    verdicts = []
    for group in ("allelefreqs", "posteriors"):
        for expr in (result["expected"] or {}).get(group, []) or []:
            py = expr.replace("&&", " and ").replace("||", " or ")
            try:
                ok = bool(eval(py, {"__builtins__": {}}, env))  # noqa: S307 -- fixture expressions
            except Exception:
                ok = False
            verdicts.append((expr, ok))
    return verdicts


class GpuBackend:
    """The two compute stages on the GPU context (C-ABI entry points)."""

    def __init__(self, ctx, gap=None, flags=None):
        from . import GapParams, HMM_DEFAULT
        self.ctx, self.gap, self.flags = ctx, gap or GapParams(), HMM_DEFAULT if flags is None else flags

    def allele_support(self, jobs):
        haps, extra, reads, quals = [], [], [], []
        rr, rh, rho, rnr = [], [], [0], []
        cache = {}
        for j in jobs:
            ids = []
            for key, ex in [(j["ref"], 0)] + list(j["alts"]):
                if (key, ex) not in cache:
                    cache[(key, ex)] = len(haps)
                    haps.append(np.frombuffer(key, np.uint8))
                    extra.append(ex)
                ids.append(cache[(key, ex)])
            rr.append(len(reads))
            reads.append(np.frombuffer(j["bases"], np.uint8))
            quals.append(np.frombuffer(j["quals"], np.uint8))
            rh += ids
            rho.append(len(rh))
            rnr.append(1)
        hap_off = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int64)
        read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
        got = self.ctx.allele_support_batch(np.array(rr, np.int32), np.array(rho, np.int32), np.array(rh, np.int32),
                                            np.array(rnr, np.int32), np.concatenate(haps), hap_off,
                                            np.concatenate(reads), np.concatenate(quals), read_off, self.gap,
                                            shrink_extra=np.array(extra, np.int32), flags=self.flags)
        out = []
        for k in range(len(jobs)):
            n = int(got["n_indel_ops"][k])
            if got["indel_overflow"][k]:
                raise NotImplementedError("job %d: more than 32 indel runs in the alt hit's alignment" % k)
            out.append(dict(prob_ref=float(got["prob_ref"][k]), prob_alt=float(got["prob_alt"][k]),
                            indel_ops=[int(x) for x in got["indel_ops"][k][:n]]))
        return out

    def posterior(self, scn, cols, cat, off):
        return self.ctx.posterior_batch(scn, cols, cat, off)

    def call_records(self, scn, post, prob_mapping, off):
        return self.ctx.call_records(scn, post, prob_mapping, off)

    def model_leaves(self, scn):
        from . import VlrError
        try:
            return self.ctx.model_leaves({k: v for k, v in scn.items() if k not in ("prior_fn", "prior_prog")})
        except VlrError as e:
            if e.code == -4:  # VLR_ERR_UNSUPPORTED: leaves are enumerable only for set-only tree paths
                return None
            raise
