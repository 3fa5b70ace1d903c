"""Grammar front end (libprosic_b200/grammar.py): formula / universe parser,
VAFRange algebra, normalisation, VAF trees and the flattening to the vlr_model layout.

Checked against (i) the reference's own unit tests of the module (src/grammar/formula.rs:1343-1428,
restated below with the same scenario texts), (ii) the scenarios that libprosic_b200/scenario.py and
fixtures.py flatten by hand -- the grammar-derived model must give the same posteriors -- and
(iii) the reference's integration testcases that come with a scenario.yaml (the scenario texts are
quoted from tests/resources/testcases/*/scenario.yaml and src/cli.rs:918-940)."""
import glob
import os

import numpy as np
import pytest

from libprosic_b200 import fixtures as F
from libprosic_b200 import grammar as G
from libprosic_b200 import scenario, synth

HERE = os.path.dirname(os.path.abspath(__file__))

TN_YAML = """
samples:
  tumor:
    resolution: 100
    contamination:
      by: normal
      fraction: 0.25
    universe: "[0.0,1.0]"
  normal:
    resolution: 5
    universe: "[0.0,0.5[ | 0.5 | 1.0"
events:
  somatic_tumor:  "tumor:]0.0,1.0] & normal:0.0"
  somatic_normal: "tumor:]0.0,1.0] & normal:]0.0,0.5["
  germline_het:   "tumor:]0.0,1.0] & normal:0.5"
  germline_hom:   "tumor:]0.0,1.0] & normal:1.0"
"""

TRIO_YAML = """
species:
  heterozygosity: 0.001
  ploidy: 2
  genome-size: 3.5e9
  germline-mutation-rate: 0.1 #1.2e-8

samples:
  index:
    inheritance:
      mendelian:
        from:
          - mother
          - father
  mother:
    sex: female
  father:
    sex: male

events:
  denovo: "(index:0.5 | index:1.0) & father:0.0 & mother:0.0"
  not_interesting: "!father:0.0 | !mother:0.0"
"""

FFPE_YAML = """
samples:
  tumor:
    resolution: 100
    universe: "[0.0,1.0]"

events:
  # In reality, it should be C>T, but this test case is T>C.
  ffpe_artifact: "(T>C | G>A) & tumor:]0.0,0.1["
  present: "((T>C | G>A) & tumor:]0.1,1.0]) | (!(T>C | G>A) & tumor:]0.0,1.0[)"
"""

FFPE57_YAML = """
samples:
  tumor:
    resolution: 100
    universe: "[0.0,1.0]"
#    contamination:
#      fraction: 0.0

events:
  ffpe_artifact: "(C>T | G>A) & tumor:]0.0,0.05["
  present: "((C>T | G>A) & tumor:[0.05,1.0]) | (!(C>T | G>A) & tumor:]0.0,1.0])"
"""

PRESENT_YAML = """
samples:
  tumor:
    resolution: 100
    universe: "[0.0,1.0]"

events:
  present: "tumor:]0.0,1.0]"
"""

GIAB03_YAML = """
species:
  heterozygosity: 0.001
  ploidy:
    male:
      all: 2
      Y: 1
      X: 1
    female:
      all: 2
      X: 2
      Y: 0

samples:
  NA12878:
    sex: female

events:
  present: "NA12878:0.5 | NA12878:1.0"
"""


def _single(universe="[0.0,1.0]", **events):
    return G.Scenario("samples:\n  normal:\n    resolution: 100\n    universe: \"%s\"\nevents:\n%s" % (
        universe, "".join("  %s: \"%s\"\n" % kv for kv in events.items())))


# ------------------------------------------------------------------ parser
def test_parser_atoms_and_operators():
    assert G.parse_formula("tumor:0.5") == ("atom", "tumor", ("set", (0.5,)))
    assert G.parse_formula("tumor:]0.0,1.0]") == ("atom", "tumor", ("range", 0.0, 1.0, True, False))
    assert G.parse_formula("a:0.5 & b:1.0 & c:[0.1,0.2[")[0] == "and"
    f = G.parse_formula("(index:0.5 | index:1.0) & father:0.0 & mother:0.0")
    assert f[0] == "and" and f[1][0][0] == "or" and len(f[1]) == 3
    assert G.parse_formula("!father:0.0 | !mother:0.0") == (
        "or", [("not", ("atom", "father", ("set", (0.0,)))), ("not", ("atom", "mother", ("set", (0.0,))))])
    assert G.parse_formula("$hom") == ("expr", "hom", False)
    assert G.parse_formula("C>T") == ("variant", True, "C", "T")
    assert G.parse_formula("A:0.5") == ("atom", "A", ("set", (0.5,)))  # an identifier, not a variant
    assert G.parse_formula("  a:0.5   /* comment */ | a:1.0 ")[0] == "or"
    assert G.parse_formula("((a:0.5)) & b:1.0")[1][0] == ("atom", "a", ("set", (0.5,)))
    with pytest.raises(G.ParseError):   # formula.pest: no parenthesised subformula at the top level
        G.parse_formula("(a:0.5)")
    f = G.parse_formula("((T>C | G>A) & tumor:]0.1,1.0]) | (!(T>C | G>A) & tumor:]0.0,1.0[)")
    assert f[0] == "or" and f[1][1][1][0] == ("not", ("or", [("variant", True, "T", "C"), ("variant", True, "G", "A")]))


@pytest.mark.parametrize("bad", ["a:0.5 & b:1.0 | c:0.0", "a:1.5", "a:0.", "a:[0.1,0.2", "a", "& a:0.5", "a:0.5 &", "a:2"])
def test_parser_rejects(bad):
    with pytest.raises(G.ParseError):
        G.parse_formula(bad)


def test_universe_parser():
    assert G.parse_universe("[0.0,0.5[ | 0.5 | 1.0") == [("range", 0.0, 0.5, False, True), ("set", (0.5,)), ("set", (1.0,))]
    assert G.parse_universe("0.0 | 0.5 | 1.0 | ]0.0,0.5[")[3] == ("range", 0.0, 0.5, True, True)
    with pytest.raises(G.ParseError):
        G.parse_universe("[0.0,1.0] &")


# ------------------------------------------------------------------ the reference's unit tests
def test_vaf_range_overlap():  # formula.rs:1350-1368
    r1, r2 = G.vrange(0.0, 0.7, False, True), G.vrange(0.3, 1.0, False, False)
    assert G.range_and(r1, r2) == G.vrange(0.3, 0.7, False, True)


def test_range_conjunction():  # formula.rs:1370-1393
    scn = _single(full="normal:[0.0,1.0]", part1="normal:[0.0,0.7]", part2="normal:[0.3,1.0]",
                  expected="normal:[0.3,0.7]")
    conj = G.normalize(("and", [scn.events["part1"], scn.events["part2"]]), scn)
    assert conj == G.normalize(scn.events["expected"], scn)
    assert conj != G.normalize(scn.events["full"], scn)


def test_nested_range_disjunction():  # formula.rs:1395-1410
    scn = _single(full="(normal:[0.0, 0.25] | normal:[0.5,0.75]) | (normal:[0.25,0.5] | normal:[0.75,1.0]) | "
                       "normal:[0.1,0.4] | normal:0.1", expected="normal:[0.0,1.0]")
    assert G.normalize(scn.events["full"], scn) == G.normalize(scn.events["expected"], scn)


def test_two_separate_range_disjunctions():  # formula.rs:1412-1427
    scn = _single(full="(normal:[0.0, 0.25] | normal:[0.5,0.6]) | ((normal:[0.25,0.5] | normal:[0.7,0.9]) | "
                       "normal:[0.9,1.0]) | normal:]0.8,0.9[ | normal:0.75",
                  expected="normal:[0.0,0.6] | normal:[0.7,1.0]")
    assert G.normalize(scn.events["full"], scn) == G.normalize(scn.events["expected"], scn)


# ------------------------------------------------------------------ range algebra / negation
def test_range_split_and_contains():
    r = G.vrange(0.0, 1.0, False, False)
    assert G.range_split_at(r, 0.5) == (G.vrange(0.0, 0.5, False, True), G.vrange(0.5, 1.0, True, False))
    assert G.range_split_at(r, 0.0) == (G.vset([0.0]), G.vrange(0.0, 1.0, True, False))
    # right end of an inclusive range: the right part is the point itself ... unless both ends are exclusive
    assert G.range_split_at(r, 1.0)[0] == G.vrange(0.0, 1.0, False, True)
    assert G.range_contains(G.vrange(0.0, 0.5, True, True), 0.25) and not G.range_contains(G.vrange(0.0, 0.5, True, True), 0.5)
    assert G.range_overlap(G.vrange(0.0, 0.25, False, False), G.vrange(0.25, 0.5, False, False)) == "end"
    assert G.range_overlap(G.vrange(0.0, 0.25, False, True), G.vrange(0.25, 0.5, False, False)) == "none"
    assert G.range_overlap(G.vrange(0.1, 0.2, False, False), G.vrange(0.0, 1.0, False, False)) == "contained"
    assert G.range_overlap(G.vrange(0.0, 1.0, False, False), G.vrange(0.1, 0.2, False, False)) == "contains"


def test_negation_against_the_universe():
    scn = G.Scenario(TRIO_YAML)
    # ploidy 2 without a somatic rate: universe {0, .5, 1}; "!father:0.0" = father:{.5, 1}
    n = G.normalize(G.parse_formula("!father:0.0"), scn)
    assert n == ("atom", "father", ("set", (0.5, 1.0)))
    # tumor/normal: "!normal:0.5" splits nothing of [0,0.5[, keeps 1.0
    tn = G.Scenario(TN_YAML)
    n = G.normalize(G.parse_formula("!normal:0.5"), tn)
    assert n[0] == "or" and set(n[1]) == {("atom", "normal", ("range", 0.0, 0.5, False, True)),
                                         ("atom", "normal", ("set", (1.0,)))}
    # a range negated inside a range universe
    s = _single(e="!normal:[0.2,0.4]")
    n = G.normalize(s.events["e"], s)
    assert n[0] == "or" and set(n[1]) == {("atom", "normal", ("range", 0.0, 0.2, False, True)),
                                         ("atom", "normal", ("range", 0.4, 1.0, True, False))}
    # double negation and De Morgan
    assert G.normalize(G.parse_formula("!(!father:0.0)"), scn) == ("atom", "father", ("set", (0.0,)))
    n = G.normalize(G.parse_formula("!(father:0.0 | mother:0.0)"), scn)
    assert n[0] == "and" and set(n[1]) == {("atom", "father", ("set", (0.5, 1.0))), ("atom", "mother", ("set", (0.5, 1.0)))}


def test_expressions_and_false():
    scn = G.Scenario("""
samples:
  normal:
    resolution: 5
    universe: "0.0 | 0.5 | 1.0 | ]0.0,0.5["
expressions:
  hom: "normal:1.0"
events:
  germline_hom: "$hom"
  germline_het: "normal:0.5"
  nothing: "normal:0.5 & normal:1.0"
  not_hom: "!$hom"
""")
    assert G.normalize(scn.events["germline_hom"], scn) == ("atom", "normal", ("set", (1.0,)))
    assert G.normalize(scn.events["nothing"], scn) == G.FALSE
    n = G.normalize(scn.events["not_hom"], scn)
    assert n[0] == "or" and ("atom", "normal", ("range", 0.0, 0.5, True, True)) in n[1]
    # events are usable as expressions, and so is `absent` (grammar/mod.rs:150-166)
    assert G.normalize(G.parse_formula("$germline_het | $absent"), scn) == ("atom", "normal", ("set", (0.0, 0.5)))
    with pytest.raises(ValueError):
        G.normalize(G.parse_formula("$undefined"), scn)


def test_sample_universes():
    scn = G.Scenario(GIAB03_YAML)
    assert scn.contig_universe("NA12878", "1") == [("set", (0.0, 0.5, 1.0))]
    assert scn.contig_universe("NA12878", "Y") == [("set", (0.0,))]       # ploidy 0
    s2 = G.Scenario("""
species:
  ploidy: 2
samples:
  a:
    somatic-effective-mutation-rate: 1e-10
  b:
    universe:
      all: "0.0 | 0.5 | 1.0"
      X: "0.0 | 1.0"
events:
  e: "a:0.5"
""")
    assert s2.contig_universe("a", "1") == [("range", 0.0, 0.5, True, True), ("range", 0.5, 1.0, True, True),
                                            ("set", (0.0, 0.5, 1.0))]
    assert s2.contig_universe("b", "X") == [("set", (0.0,)), ("set", (1.0,))]
    assert s2.contig_universe("b", "7") == [("set", (0.0,)), ("set", (0.5,)), ("set", (1.0,))]


def test_overlapping_events_are_reported():
    scn = G.Scenario("""
samples:
  n:
    universe: "0.0 | 0.5 | 1.0"
events:
  het: "n:0.5"
  hom: "n:1.0"
  present: "n:0.5 | n:1.0"
""")
    with pytest.raises(ValueError, match="overlapping"):
        G.flatten(scn)


def test_all_reference_scenarios():
    """Every scenario.yaml of the reference's testcases parses, normalises and flattens; frozen in
    tests/golden/grammar_scenarios.json by tests/golden/make_grammar_golden.py."""
    import json
    gold = json.load(open(os.path.join(HERE, "golden", "grammar_scenarios.json")))
    assert len(gold) == 47
    for name, entry in gold.items():
        s = G.Scenario(entry["yaml"])
        assert s.sample_names == entry["samples"] == sorted(entry["samples"]), name
        for contig, want in entry["contigs"].items():
            if "error" in want:
                with pytest.raises(ValueError):
                    G.flatten(s, contig=contig, snv=("C", "T"))
                continue
            got = {k: G.format_normalized(G.normalize(f, s, contig)) for k, f in s.events.items()}
            assert got == want["normalized"], (name, contig)
            m = G.flatten(s, contig=contig, snv=("C", "T"))
            assert (len(m["nodes"]), len(m["events"]), len(m["biases"])) == (want["n_nodes"], want["n_events"], want["n_biases"])
            # every tree path binds every sample exactly once (what flatten_model of the library requires)
            for ev in m["events"]:
                stack = [(r, ()) for r in ev["roots"]]
                while stack:
                    k, seen = stack.pop()
                    nd = m["nodes"][k]
                    if nd["kind"] == scenario.NODE_FALSE:
                        continue
                    if nd["kind"] in (scenario.NODE_SET, scenario.NODE_RANGE):
                        assert nd["sample"] not in seen, (name, ev["name"])
                        seen = seen + (nd["sample"],)
                    if not nd["children"]:
                        assert sorted(seen) == list(range(m["n_samples"])), (name, ev["name"])
                    stack += [(c, seen) for c in nd["children"]]


# ------------------------------------------------------------------ VAF trees
def _paths(roots):
    out = []

    def walk(n, acc):
        acc = acc + [(n.kind, n.sample, n.vafs, n.variant)]
        if not n.children:
            out.append(acc)
        for c in n.children:
            walk(c, acc)
    for r in roots:
        walk(r, [])
    return out


def test_vaftree_adds_missing_samples():
    scn = G.Scenario(TRIO_YAML)   # father = 0, index = 1, mother = 2
    tree = G.vaftree(G.normalize(scn.events["denovo"], scn), scn)
    paths = _paths(tree)
    # sum of products: (index:0.5 & father:0 & mother:0) | (index:1.0 & father:0 & mother:0)
    assert len(paths) == 2 and all(sorted(n[1] for n in p) == [0, 1, 2] for p in paths)
    ni = G.vaftree(G.normalize(scn.events["not_interesting"], scn), scn)
    paths = _paths(ni)
    # two roots (father != 0 | mother != 0), the other two samples appended with their universe
    assert len(ni) == 2 and all(sorted(n[1] for n in p) == [0, 1, 2] for p in paths)
    # a universe of several spectra fans out
    tn = G.Scenario(TN_YAML)
    t = G.vaftree(G.normalize(G.parse_formula("tumor:0.5"), tn), tn)
    assert len(_paths(t)) == 3


# ------------------------------------------------------------------ same posteriors as the hand-flattened scenarios
def _posterior(scn, cols, cat, off):
    from _oracle_backend import OracleBackend
    return OracleBackend().posterior(scn, cols, cat, off)


def _by_name(scn, post):
    """event name -> ln of the summed posterior mass of its (unbiased, artifact) entries per site."""
    out = {}
    for k, ev in enumerate(scn["events"]):
        key = (ev["name"], ev["biases"] == [0])
        assert key not in out
        out[key] = post["event_ln"][:, k]
    return out


def test_tumor_normal_scenario_from_yaml_matches_the_hand_flattened_one():
    g = G.flatten(G.Scenario(TN_YAML), snv=None, snv_or_mnv=False)
    h = scenario.tumor_normal(purity=0.75, indel=True)
    assert g["sample_names"] == ["normal", "tumor"] and g["contaminated"] == [0, 1] and g["by"] == [0, 0]
    assert g["purity"] == [1.0, 0.75] and g["resolution"] == [5, 100] and g["biases"] == h["biases"]
    cols, cat, off, _ = synth.make_pileups(6, [12, 25], seed=11, kind="indel",
                                           af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]],
                                           purity=[1.0, 0.75], artifact_frac=0.2)
    pg, ph = _posterior(g, cols, cat, off), _posterior(h, cols, cat, off)
    bg, bh = _by_name(g, pg), _by_name(h, ph)
    assert set(bg) == set(bh)
    for k in bg:
        both = np.isfinite(bg[k]) & np.isfinite(bh[k])
        assert np.array_equal(np.isfinite(bg[k]), np.isfinite(bh[k]))
        assert np.max(np.abs(bg[k][both] - bh[k][both]), initial=0.0) < 1e-9, k
    assert np.allclose(pg["marginal"], ph["marginal"], rtol=0, atol=1e-9)
    assert np.array_equal(pg["map_af"], ph["map_af"], equal_nan=True)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "trio", "*.json.gz"))),
                         ids=lambda p: os.path.basename(p).split(".")[0])
def test_trio_scenario_from_yaml(path):
    from _oracle_backend import OracleBackend
    want = F.run_trio_testcase(path, OracleBackend())
    got = F.run_trio_testcase(path, OracleBackend(), scenario_yaml=TRIO_YAML)
    assert all(v for _, v in F.check_expectations(got))
    assert got["af"] == want["af"]
    for k, w in want.items():
        if k.startswith("PROB_"):
            assert (np.isinf(w) and np.isinf(got[k])) or abs(got[k] - w) <= 1e-6 * max(1.0, abs(w)), (k, got[k], w)


@pytest.mark.parametrize("name,text", [("test40", FFPE_YAML), ("test41", PRESENT_YAML), ("test57", FFPE57_YAML),
                                       ("test59", FFPE57_YAML), ("test_giab_03", GIAB03_YAML)])
def test_generic_scenarios_from_yaml(name, text):
    """Variant-type atoms (`T>C | G>A`) resolved against the candidate; ploidy-derived universe."""
    from _oracle_backend import OracleBackend
    path = os.path.join(HERE, "golden", "gen", name + ".json.gz")
    want = F.run_generic_case(path, OracleBackend())
    got = F.run_generic_case(path, OracleBackend(), scenario_yaml=text)
    assert all(v for _, v in F.check_expectations(got))
    assert got["af"] == want["af"]
    for k, w in want.items():
        if k.startswith("PROB_"):
            assert (np.isinf(w) and np.isinf(got[k])) or abs(got[k] - w) <= 1e-6 * max(1.0, abs(w)), (k, got[k], w)


@pytest.mark.gpu
def test_grammar_scenarios_on_the_gpu(ctx):
    """The grammar-derived models run through vlr_posterior_batch like the hand-flattened ones."""
    g = G.flatten(G.Scenario(TN_YAML), snv=None, snv_or_mnv=False)
    cols, cat, off, _ = synth.make_pileups(64, [20, 45], seed=5, kind="indel",
                                           af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]],
                                           purity=[1.0, 0.75], artifact_frac=0.2)
    got = ctx.posterior_batch(g, cols, cat, off)
    want = _posterior(g, cols, cat, off)
    ev = got["event_ln"].reshape(want["event_ln"].shape)
    both = np.isfinite(ev) & np.isfinite(want["event_ln"])
    assert np.array_equal(np.isfinite(ev), np.isfinite(want["event_ln"]))
    assert np.max(np.abs(ev[both] - want["event_ln"][both])) < 1e-6
    assert np.array_equal(got["map_af"].reshape(want["map_af"].shape), want["map_af"], equal_nan=True)
    for path in sorted(glob.glob(os.path.join(HERE, "golden", "trio", "*.json.gz"))):
        r = F.run_trio_testcase(path, F.GpuBackend(ctx), scenario_yaml=TRIO_YAML)
        assert all(v for _, v in F.check_expectations(r))


def test_prior_module_matches_the_hand_written_priors():
    """libprosic_b200/prior.py (general restatement of Prior::calc_prob without somatic terms) against
    the two special cases fixtures.py carried before: diploid trio and one germline sample."""
    import itertools
    from libprosic_b200 import prior as PR
    trio = G.Scenario(TRIO_YAML)          # father = 0, index = 1, mother = 2
    f = PR.make_prior(trio, "1", "snv")
    h = F.mendelian_trio_prior(0.001, 0.1, child=1, parents=(2, 0))
    for afs in itertools.product([0.0, 0.5, 1.0], repeat=3):
        a, b = f(list(afs)), h(list(afs))
        assert (np.isinf(a) and np.isinf(b)) or abs(a - b) < 1e-12, (afs, a, b)
    assert f([0.25, 0.0, 0.0]) == -np.inf          # not a germline VAF at ploidy 2
    one = G.Scenario(GIAB03_YAML)
    f1, h1 = PR.make_prior(one, "1", "snv"), F.germline_prior(0.001)
    for af in (0.0, 0.5, 1.0):
        assert abs(f1([af]) - h1([af])) < 1e-12
    # indels: heterozygosity scaled by the variant-type fraction (default 0.0125)
    fi = PR.make_prior(one, "1", "del")
    assert abs(fi([0.5]) - np.log(0.001 * 0.0125)) < 1e-12
    # ploidy 0 (female, Y): only AF 0 is possible
    fy = PR.make_prior(one, "Y", "snv")
    assert fy([0.5]) == -np.inf and fy([0.0]) == 0.0
    # somatic mutation rate (prior.rs:433-455): a non-germline VAF gets the density rate / (vaf^2 * genome size),
    # summed over the possible germline genotypes; a germline VAF the complement of the integrated density
    som = G.Scenario(TRIO_YAML.replace("  mother:\n    sex: female",
                                       "  mother:\n    sex: female\n    somatic-effective-mutation-rate: 1e-6"))
    fs = PR.make_prior(som, "1", "snv")
    base = f([0.0, 0.0, 0.0])
    want = np.log(1e-6) - (2 * np.log(0.25) + np.log(3.5e9))
    got = fs([0.0, 0.0, 0.25])
    assert np.isfinite(got) and abs((got - base) - want) < 1e-3     # germline 0 dominates the sum
    assert abs(fs([0.0, 0.0, 0.0]) - base) < 1e-6                   # ln(1 - tiny)
    with pytest.raises(NotImplementedError):
        PR.make_prior(G.Scenario(TRIO_YAML.replace("  mother:\n    sex: female",
                                                   "  mother:\n    sex: female\n    inheritance:\n      subclonal:\n"
                                                   "        from: father\n        origin: single-cell")), "1")
