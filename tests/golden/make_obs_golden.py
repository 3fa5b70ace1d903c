#!/usr/bin/env python
"""Extract the observation records embedded in the reference's testcases into a small fixture.

The candidate VCFs of many testcases under /root/reference/tests/resources/testcases still carry
the INFO fields `varlociraptor preprocess` wrote upstream (PROB_MAPPING ... PAIRED; format version
7 = version 8 without INDEL_OPERATIONS).  They are genuine outputs of the reference's encoder and
pin the wire format (oracle/obs_codec.py) and serve as real-data pileups for part 2.

    python tests/golden/make_obs_golden.py     # needs /root/reference (this container only)
writes tests/golden/obs_records.json: [{"name", "version", "chrom", "pos", "ref", "alt", "fields"}].
"""
import glob
import json
import os
import re

REF = "/root/reference/tests/resources/testcases"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "obs_records.json")
TAGS = ["PROB_MAPPING", "PROB_REF", "PROB_ALT", "PROB_MISSED_ALLELE", "PROB_SAMPLE_ALT",
        "PROB_DOUBLE_OVERLAP", "PROB_HIT_BASE", "STRAND", "READ_ORIENTATION", "READ_POSITION",
        "SOFTCLIPPED", "INDEL_OPERATIONS", "PAIRED"]


def main():
    recs = []
    for path in sorted(glob.glob(os.path.join(REF, "*", "candidates*.vcf"))):
        txt = open(path, encoding="latin-1").read()
        if txt[:2] == "\x1f\x8b":
            continue  # compressed candidate file
        m = re.search(r"observation_format_version=(\d+)", txt)
        if not m or int(m.group(1)) < 7:
            continue  # older formats lack fields the current reader needs
        for line in txt.split("\n"):
            if not line or line.startswith("#"):
                continue
            c = line.split("\t")
            info = dict(kv.split("=", 1) for kv in c[7].split(";") if "=" in kv)
            if "PROB_MAPPING" not in info:
                continue
            fields = {t: [int(x) for x in info[t].split(",")] for t in TAGS if t in info}
            n_obs = fields["PROB_MAPPING"][0]
            if n_obs > 400:
                continue  # keep the fixture small
            recs.append(dict(name=os.path.basename(os.path.dirname(path)), version=int(m.group(1)),
                             chrom=c[0], pos=int(c[1]), ref=c[3], alt=c[4][:40], fields=fields))
    json.dump(recs, open(OUT, "w"), separators=(",", ":"))
    print("wrote %d records, %d bytes" % (len(recs), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
