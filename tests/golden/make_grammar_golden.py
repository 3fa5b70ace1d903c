#!/usr/bin/env python
"""Writes tests/golden/grammar_scenarios.json: every scenario.yaml of the reference's testcases
(tests/resources/testcases/*/scenario.yaml, quoted verbatim -- they are test inputs) with the
normalised events and the size of the flattened model that libprosic_b200/grammar.py derives from it.
Run in the build container (needs /root/reference); the test suite only reads the JSON.  The file
freezes the restatement; it is NOT output of the reference (no Rust toolchain here)."""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from libprosic_b200 import grammar as G  # noqa: E402

out = {}
for p in sorted(glob.glob("/root/reference/tests/resources/testcases/*/scenario.yaml")):
    name = p.split("/")[-2]
    text = open(p).read()
    s = G.Scenario(text)
    entry = {"yaml": text, "samples": s.sample_names, "contigs": {}}
    for contig in ("1", "X"):
        try:
            m = G.flatten(s, contig=contig, snv=("C", "T"))
        except ValueError as e:
            entry["contigs"][contig] = {"error": str(e)}
            continue
        entry["contigs"][contig] = {
            "normalized": {k: G.format_normalized(G.normalize(f, s, contig)) for k, f in s.events.items()},
            "n_nodes": len(m["nodes"]), "n_events": len(m["events"]), "n_biases": len(m["biases"]),
            "universe": {n: [list(sp[1]) if sp[0] == "set" else list(sp[1:]) for sp in s.contig_universe(n, contig)]
                         for n in s.sample_names}}
    out[name] = entry
json.dump(out, open(os.path.join(HERE, "grammar_scenarios.json"), "w"), indent=1, sort_keys=True)
print(len(out), "scenarios")
