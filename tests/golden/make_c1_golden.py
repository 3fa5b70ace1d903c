#!/usr/bin/env python
"""Config C1 (bundled reference testcases) for part 1: real read windows and haplotype windows.

Reads the tumor/normal indel fixtures of the reference (tests/resources/testcases/testNN:
candidates.vcf, tumor.bam, normal.bam, reference sequence inside testcase.yaml), restates the host
glue that feeds the realigner and writes the (read window, ref window, alt window) triples it
produces to tests/golden/c1_realign.json:

  * Realigner::candidate_region (src/variants/evidence/realignment/mod.rs:61-156) with the
    recollected semantics of rust-htslib 0.36 CigarStringView::read_pos;
    max_window 64 (cli.rs:654-661), ref_window = 96 (mod.rs:158-162), max_pattern_len 128
  * reference window: ReferenceEmissionParams (realignment/pairhmm.rs:213-237)
  * alt windows: DeletionEmissionParams (types/deletion.rs:129-152, 300-312),
    InsertionEmissionParams (types/insertion.rs:47-74, 200-222; the window grows by ins_len)
  * variant loci from the VCF record as preprocessing/mod.rs:318-356 builds them

This is test-input generation (it needs /root/reference, i.e. this container); parity on these
inputs is GPU vs oracle (tests/test_c1_gpu.py).  BAM decoding: BGZF members via gzip, records per
the SAM/BAM specification.
"""
import gzip
import json
import os
import struct

import yaml

REF = "/root/reference/tests/resources/testcases"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_realign.json")
CASES = ["test02", "test03", "test04", "test05", "test09", "test10", "test12", "test16", "test20", "test24"]
MAX_WINDOW, REF_WINDOW, MAX_PATTERN = 64, 96, 128
SEQ_CODE = "=ACMGRSVTWYHKDBN"


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
        (ref_id, p, l_rn, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl) = struct.unpack_from("<iiBBHHHiiii", data, pos + 4)
        q = pos + 36
        name = data[q:q + l_rn - 1].decode()
        q += l_rn
        cigar = [(v & 15, v >> 4) for v in struct.unpack_from("<%dI" % n_cig, data, q)]
        q += 4 * n_cig
        raw = data[q:q + (l_seq + 1) // 2]
        seq = "".join(SEQ_CODE[(raw[k >> 1] >> (4 if k % 2 == 0 else 0)) & 15] for k in range(l_seq))
        q += (l_seq + 1) // 2
        qual = list(data[q:q + l_seq])
        recs.append(dict(name=name, contig=refs[ref_id] if ref_id >= 0 else None, pos=p, mapq=mapq, flag=flag,
                         cigar=cigar, seq=seq, qual=qual))
        pos += 4 + bs
    return recs


M, I, D, N, S, H, P, EQ, X = range(9)


def read_pos(rec, ref_pos, include_softclips=True, include_dels=True):
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


def end_pos(rec):
    return rec["pos"] + sum(l for op, l in rec["cigar"] if op in (M, D, N, EQ, X))


def candidate_region(rec, l0, l1, ref_len):
    """mod.rs:61-156 -> (overlap, read interval, ref interval)."""
    seq_len = len(rec["seq"])

    def ref_iv(bp):
        return max(0, bp - REF_WINDOW), min(bp + REF_WINDOW, ref_len)
    qs, qe = read_pos(rec, l0), read_pos(rec, l1)
    if qs is not None and qe is not None:
        mw = max(0, MAX_WINDOW - (qe - qs) // 2)
        ro, re_ = max(0, qs - mw), min(qe + mw, seq_len)
        exceed = max(0, (re_ - ro) - MAX_PATTERN)
        if exceed > 0:
            ro += exceed // 2
            re_ -= (exceed + 1) // 2
        return True, (ro, re_), ref_iv(l0)
    if qs is not None:
        return True, (max(0, qs - MAX_WINDOW), min(qs + MAX_WINDOW, seq_len)), ref_iv(l0)
    if qe is not None:
        return True, (max(0, qe - MAX_WINDOW), min(qe + MAX_WINDOW, seq_len)), ref_iv(l1)
    m = seq_len // 2
    enclosed = rec["pos"] >= l0 and end_pos(rec) <= l1
    return enclosed, (max(0, m - MAX_WINDOW), min(m + MAX_WINDOW - 1, seq_len)), ref_iv(rec["pos"] + m)


def main():
    out = []
    for case in CASES:
        d = os.path.join(REF, case)
        y = yaml.safe_load(open(os.path.join(d, "testcase.yaml")))
        ref = y["reference"]["seq"].encode()
        rec = [l for l in open(os.path.join(d, "candidates.vcf")) if not l.startswith("#")][0].split("\t")
        start, r, a = int(rec[1]) - 1, rec[3], rec[4].split(",")[0]
        if a.startswith("<") or len(r) == len(a):
            continue
        if len(r) > len(a):      # deletion: interval start .. start + svlen (preprocessing/mod.rs:324-352)
            dl = len(r) - len(a)
            l0, l1 = start, start + dl
            ro, re_ = max(0, start - REF_WINDOW), min(start + REF_WINDOW, len(ref) - dl)
            alt = bytes(ref[i] if i <= start else ref[i + dl] for i in range(ro, re_))
            extra, kind = 0, "DEL%d" % dl
        else:                    # insertion: locus start .. start + 1, ins_seq = ALT without the anchor
            ins = a[len(r):].encode() if a.startswith(r) else a[1:].encode()
            il = len(ins)
            l0, l1 = start, start + 1
            ro, re_ = max(0, start - REF_WINDOW), min(start + il + REF_WINDOW, len(ref))
            alt = bytes(ref[i] if i <= start else (ref[i - il] if i > start + il else ins[i - start - 1])
                        for i in range(ro, re_ + il))
            extra, kind = il, "INS%d" % il
        regions, windows = [], {}
        for sample in ("tumor", "normal"):
            n = 0
            for b in read_bam(os.path.join(d, sample + ".bam")):
                if b["flag"] & (4 | 256 | 512 | 1024 | 2048) or not b["cigar"] or b["mapq"] == 0:
                    continue
                if end_pos(b) < l0 - 150 or b["pos"] > l1 + 150:
                    continue
                ov, (q0, q1), (r0, r1) = candidate_region(b, l0, l1, len(ref))
                if not ov or q1 - q0 < 8:
                    continue
                windows.setdefault((r0, r1), len(windows))
                regions.append(dict(ref=windows[(r0, r1)], bases=b["seq"][q0:q1], quals=b["qual"][q0:q1],
                                    sample=sample, name=b["name"]))
                n += 1
                if n >= 24:
                    break
        out.append(dict(case=case, kind=kind, alt=alt.decode(), alt_extra=extra,
                        ref_windows=[ref[r0:r1].decode() for (r0, r1), _ in sorted(windows.items(), key=lambda kv: kv[1])],
                        regions=regions))
        print(case, kind, "ref windows", len(windows), "regions", len(regions))
    json.dump(out, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
