"""Grammar front end: scenario YAML -> events -> normalised formulas -> VAF
trees -> the flattened `vlr_model` layout of include/vlr_gpu.h (see scenario.py for the dict form).

Host-side mirror of the reference's `grammar` module; it produces the INPUT of the GPU path
(vlr_posterior_batch), it is not on the path itself.

  * formula / universe syntax        src/grammar/formula.pest, parse_formula formula.rs:1244-1340
  * VAFRange algebra                 formula.rs:949-1120 (contains, split_at, overlap, &, |)
  * expressions, negation            formula.rs:403-442, 631-862
  * normalisation                    formula.rs:391-401: expand -> negate -> simplify -> merge_atoms ->
                                     simplify -> strip_false
  * VAF tree + missing samples       src/grammar/vaftree.rs:100-214
  * sample universes                 src/grammar/mod.rs:453-527 (universe | ploidy | somatic rate)
  * event universe + bias sets       src/calling/variants/calling.rs:453-504, bias/mod.rs:101-179

One step cannot be restated literally: `simplify` goes through the `boolean_expression` crate's BDD
(not vendored with the reference).  After `apply_negations` every formula is MONOTONE in its
terminals (negation only survives inside Variant terminals, which are terminals of their own), and
for a monotone function the irredundant sum of products is unique: the set of prime implicants.
`simplify` is therefore restated as "distribute to a sum of products, drop absorbed products" --
the same set of products the BDD cube list yields, in a deterministic operand order (the reference's
order is unspecified anyway: HashMap grouping).  Operand order only changes the
summation order of the posterior, i.e. the last bits.
"""
import itertools

import yaml

from . import scenario as _scn

# ----------------------------------------------------------------------------- spectra
# ("set", (v0, v1, ...)) sorted ascending (BTreeSet)  |  ("range", start, end, left_excl, right_excl)


def vset(vals):
    return ("set", tuple(sorted(set(float(v) for v in vals))))


def vrange(start, end, lex, rex):
    return ("range", float(start), float(end), bool(lex), bool(rex))


EMPTY_RANGE = vrange(0.0, 0.0, True, True)  # VAFRange::empty, formula.rs:950-956


def range_contains(r, v):  # formula.rs:958-965
    _, s, e, lex, rex = r
    return (s < v if lex else s <= v) and (e > v if rex else e >= v)


def spectrum_contains(sp, v):
    return v in sp[1] if sp[0] == "set" else range_contains(sp, v)


def range_split_at(r, v):  # formula.rs:967-996
    assert range_contains(r, v), "split_at is only defined if the VAF is contained in the range"
    _, s, e, lex, rex = r
    left, right = vrange(s, v, lex, True), vrange(v, e, True, rex)

    def to_spectrum(x):
        if x[1] == x[2]:
            return vset([x[1]]) if not (x[3] and rex) else None
        return x
    return to_spectrum(left), to_spectrum(right)


def range_overlap(a, b):  # VAFRange::overlap, formula.rs:998-1027
    if a == b:
        return "equal"
    _, s, e, lex, rex = a
    _, os_, oe, olex, orex = b
    start_right = (s >= os_) if (lex and not olex) else (s > os_)
    end_left = (e <= oe) if (rex and not orex) else (e < oe)
    if ((e < os_ or s > oe) or (e <= os_ and (rex or olex)) or (s >= oe and (lex or orex))):
        return "none"
    return {(True, True): "contained", (True, False): "start", (False, True): "end",
            (False, False): "contains"}[(start_right, end_left)]


def range_and(a, b):  # impl & , formula.rs:1076-1093
    o = range_overlap(a, b)
    if o in ("contained", "equal"):
        return a
    if o == "contains":
        return b
    if o == "start":
        return vrange(a[1], b[2], a[3], b[4])
    if o == "end":
        return vrange(b[1], a[2], b[3], a[4])
    return EMPTY_RANGE


def range_or(a, b):  # impl | , formula.rs:1095-1112 (first component)
    o = range_overlap(a, b)
    if o == "contained":
        return b
    if o in ("contains", "equal", "none"):
        return a
    if o == "start":
        return vrange(b[1], a[2], b[3], a[4])
    return vrange(a[1], b[2], a[3], b[4])  # end


# ----------------------------------------------------------------------------- formulas
# ("and", [f...]) | ("or", [f...]) | ("not", f) | ("atom", sample, spectrum)
# | ("variant", positive, refbase, altbase) | ("expr", identifier, negated) | ("false",)
FALSE = ("false",)
IUPAC = "ACGTRYSWKMBDHVN"
_IUPAC_SETS = {"R": "AG", "Y": "CT", "S": "GC", "W": "AT", "K": "GT", "M": "AC", "B": "CGT", "D": "AGT",
               "H": "ACT", "V": "ACG", "N": "ACGTN"}


def iupac_contains(code, base):  # Iupac::contains, formula.rs:22-41
    return code == base or base in _IUPAC_SETS.get(code, "") or code == "N"


class ParseError(ValueError):
    pass


class _Parser:
    """PEG of formula.pest, ordered choice and backtracking as pest does it."""

    def __init__(self, text):
        self.s, self.n = text, len(text)

    def ws(self, i):
        while i < self.n:
            if self.s[i] == " ":
                i += 1
            elif self.s.startswith("/*", i):
                j = self.s.find("*/", i + 2)
                if j < 0:
                    return i
                i = j + 2
            else:
                break
        return i

    def lit(self, i, tok):
        i = self.ws(i)
        return i + len(tok) if self.s.startswith(tok, i) else None

    # atomic rules (no inner whitespace)
    def vaf(self, i):
        i = self.ws(i)
        if self.s.startswith("1.0", i):
            return i + 3, 1.0
        if self.s.startswith("0.", i):
            j = i + 2
            while j < self.n and self.s[j].isdigit() and self.s[j].isascii():
                j += 1
            if j > i + 2:
                return j, float(self.s[i:j])
        return None

    def bound(self, i):
        i = self.ws(i)
        return (i + 1, self.s[i]) if i < self.n and self.s[i] in "[]" else None

    def vafrange(self, i):
        b0 = self.bound(i)
        if not b0:
            return None
        lo = self.vaf(b0[0])
        if not lo:
            return None
        c = self.lit(lo[0], ",")
        if c is None:
            return None
        hi = self.vaf(c)
        if not hi:
            return None
        b1 = self.bound(hi[0])
        if not b1:
            return None
        return b1[0], vrange(lo[1], hi[1], b0[1] == "]", b1[1] == "[")

    def identifier(self, i):
        i = self.ws(i)
        j = i
        while j < self.n and (self.s[j].isascii() and (self.s[j].isalnum() or self.s[j] in "_-.")):
            j += 1
        return (j, self.s[i:j]) if j > i else None

    def variant(self, i):
        i = self.ws(i)
        if i < self.n and self.s[i] in IUPAC:
            g = self.lit(i + 1, ">")
            if g is not None:
                g = self.ws(g)
                if g < self.n and self.s[g] in IUPAC:
                    return g + 1, ("variant", True, self.s[i], self.s[g])
        return None

    def sample_vafdef(self, i):
        ident = self.identifier(i)
        if not ident:
            return None
        c = self.lit(ident[0], ":")
        if c is None:
            return None
        v = self.vaf(c)  # sample_vaf first, then sample_vafrange
        if v:
            return v[0], ("atom", ident[1], vset([v[1]]))
        r = self.vafrange(c)
        if r:
            return r[0], ("atom", ident[1], r[1])
        return None

    def expression(self, i):
        d = self.lit(i, "$")
        if d is None:
            return None
        ident = self.identifier(d)
        return (ident[0], ("expr", ident[1], False)) if ident else None

    def nary(self, i, op, kind):
        first = self.subformula(i)
        if not first:
            return None
        ops, j = [first[1]], first[0]
        while True:
            k = self.lit(j, op)
            if k is None:
                break
            nxt = self.subformula(k)
            if not nxt:
                break
            ops.append(nxt[1])
            j = nxt[0]
        return (j, (kind, ops)) if len(ops) > 1 else None

    def paren(self, i, inner):
        o = self.lit(i, "(")
        if o is None:
            return None
        r = inner(o)
        if not r:
            return None
        c = self.lit(r[0], ")")
        return (c, r[1]) if c is not None else None

    def negation(self, i):
        b = self.lit(i, "!")
        if b is None:
            return None
        r = self.subformula(b)
        return (r[0], ("not", r[1])) if r else None

    def subformula(self, i):
        for alt in (self.variant, self.sample_vafdef,
                    lambda k: self.paren(k, lambda q: self.nary(q, "&", "and")),
                    lambda k: self.paren(k, lambda q: self.nary(q, "|", "or")),
                    self.negation, self.expression, lambda k: self.paren(k, self.subformula)):
            r = alt(i)
            if r:
                return r
        return None

    def formula(self):
        for alt in (lambda k: self.nary(k, "&", "and"), lambda k: self.nary(k, "|", "or"), self.negation,
                    self.sample_vafdef, self.variant, self.expression):
            r = alt(0)
            if r and self.ws(r[0]) == self.n:
                return r[1]
        raise ParseError("invalid VAF formula: %r" % self.s)

    def universe(self):
        out, i = [], 0
        while True:
            r = self.vaf(i)
            if r:
                out.append(vset([r[1]]))
            else:
                r = self.vafrange(i)
                if not r:
                    raise ParseError("invalid VAF universe: %r" % self.s)
                out.append(r[1])
            i = r[0]
            k = self.lit(i, "|")
            if k is None:
                break
            i = k
        if self.ws(i) != self.n:
            raise ParseError("invalid VAF universe: %r" % self.s)
        return out


def parse_formula(text):
    return _Parser(text).formula()


def parse_universe(text):
    """VAFUniverse (a HashSet in the reference; here: the listed order, duplicates removed)."""
    out = []
    for sp in _Parser(text).universe():
        if sp not in out:
            out.append(sp)
    return out


# ----------------------------------------------------------------------------- scenario
class Sample:
    def __init__(self, name, d):
        d = d or {}
        known = {"contamination", "resolution", "universe", "somatic-effective-mutation-rate",
                 "germline-mutation-rate", "ploidy", "inheritance", "sex"}
        if set(d) - known:
            raise ParseError("sample %s: unknown field(s) %s" % (name, sorted(set(d) - known)))  # deny_unknown_fields
        self.name = name
        self.contamination = d.get("contamination")
        self.resolution = int(d.get("resolution", 100))
        u = d.get("universe")
        self.universe = None if u is None else (
            {k: parse_universe(v) for k, v in u.items()} if isinstance(u, dict) else parse_universe(u))
        self.somatic_rate = d.get("somatic-effective-mutation-rate")
        self.germline_rate = d.get("germline-mutation-rate")
        self.ploidy = d.get("ploidy")
        self.inheritance = d.get("inheritance")
        self.sex = d.get("sex")


def _contig_ploidy_def(p, contig):  # PloidyDefinition::contig_ploidy, grammar/mod.rs:300-317
    if isinstance(p, dict):
        if contig in p:
            return int(p[contig])
        if "all" not in p:
            raise ValueError("ploidy of contig %s not found" % contig)
        return int(p["all"])
    return int(p)


class Scenario:
    """grammar::Scenario (grammar/mod.rs:131-271).  Samples and events are BTreeMaps: iteration and
    sample indices follow the sorted names (quirk Q12)."""

    def __init__(self, text):
        y = yaml.safe_load(text)
        if set(y) - {"expressions", "events", "samples", "species"}:
            raise ParseError("unknown scenario field(s)")
        self.samples = {k: Sample(k, v) for k, v in sorted(y["samples"].items())}
        self.events = {k: parse_formula(v) for k, v in sorted(y["events"].items())}
        self.expressions = {k: parse_formula(v) for k, v in (y.get("expressions") or {}).items()}
        self.species = y.get("species")
        self.sample_names = list(self.samples)
        # from_path registers every event, and `absent`, as an expression unless defined (mod.rs:150-166)
        for k, f in self.events.items():
            self.expressions.setdefault(k, f)
        self.expressions.setdefault("absent", absent_formula(self))

    @classmethod
    def from_path(cls, path):
        return cls(open(path).read())

    def idx(self, name):
        return self.sample_names.index(name) if name in self.samples else None

    def contig_ploidy(self, sample, contig):  # Sample::contig_ploidy, mod.rs:529-541 + SexPloidyDefinition 319-341
        s = self.samples[sample]
        if s.ploidy is not None:
            return _contig_ploidy_def(s.ploidy, contig)
        sp = (self.species or {}).get("ploidy")
        if sp is None:
            return None
        if isinstance(sp, dict) and sp and all(isinstance(v, dict) for v in sp.values()) and \
                set(sp) <= {"male", "female"}:
            if s.sex is None:
                raise ValueError("sex specific ploidy definition found but no sex specified in sample")
            if s.sex not in sp:
                raise ValueError("ploidy definition for %s not found" % s.sex)
            return _contig_ploidy_def(sp[s.sex], contig)
        return _contig_ploidy_def(sp, contig)

    def contig_universe(self, sample, contig):  # Sample::contig_universe, mod.rs:453-527
        s = self.samples[sample]
        if s.universe is not None:
            if isinstance(s.universe, dict):
                if contig in s.universe:
                    return s.universe[contig]
                if "all" not in s.universe:
                    raise ValueError("universe of contig %s not found" % contig)
                return s.universe["all"]
            return s.universe
        ploidy = self.contig_ploidy(sample, contig)
        somatic = (s.somatic_rate is not None)
        if ploidy is None and not somatic:
            raise ValueError("sample needs to define either universe, ploidy or somatic_mutation_rate")
        if ploidy is None:
            return [vrange(0.0, 1.0, False, False)]
        spec = [(k / ploidy if ploidy > 0 else 0.0) for k in range(ploidy + 1)]
        if not somatic:
            return [vset(spec)]
        uniq = sorted(set(spec))
        return [vrange(a, b, True, True) for a, b in zip(uniq, uniq[1:])] + [vset(spec)]


def absent_formula(scn):  # Formula::absent, formula.rs:375-389
    return ("and", [("atom", s, vset([0.0])) for s in scn.samples])


# ----------------------------------------------------------------------------- normalisation
def is_terminal_false(f):  # formula.rs:348-358
    return f == FALSE or (f[0] == "atom" and f[2][0] == "set" and not f[2][1])


def expand_expressions(f, scn):  # formula.rs:403-442
    k = f[0]
    if k in ("and", "or"):
        return (k, [expand_expressions(o, scn) for o in f[1]])
    if k == "not":
        return ("not", expand_expressions(f[1], scn))
    if k == "expr":
        if f[1] not in scn.expressions:
            raise ValueError("undefined expression $%s" % f[1])
        g = scn.expressions[f[1]]
        return ("not", g) if f[2] else g
    return f


def negate(f, scn, contig):  # Formula::negate, formula.rs:631-785
    k = f[0]
    if k == "and":
        return ("or", [negate(o, scn, contig) for o in f[1]])
    if k == "or":
        return ("and", [negate(o, scn, contig) for o in f[1]])
    if k == "not":
        return f[1]
    if k == "variant":
        return ("variant", not f[1], f[2], f[3])
    if k == "expr":
        return ("expr", f[1], not f[2])
    if k == "false":
        raise ValueError("negation is not defined for the false terminal")
    _, sample, vafs = f
    if sample not in scn.samples:
        raise ValueError("invalid sample name %s" % sample)
    universe = scn.contig_universe(sample, contig)
    out = []
    if vafs[0] == "set":
        stack = list(universe)
        while stack:
            u = stack.pop(0)
            if u[0] == "set":
                diff = tuple(v for v in u[1] if v not in vafs[1])
                if diff:
                    out.append(("set", diff))
            else:
                for v in vafs[1]:
                    if range_contains(u, v):
                        left, right = range_split_at(u, v)
                        if right is not None:
                            stack.append(right)
                        if left is not None:
                            out.append(left)
                    else:
                        out.append(u)
    else:
        for u in universe:
            if u[0] == "set":
                rest = tuple(v for v in u[1] if not range_contains(vafs, v))
                if rest:
                    out.append(("set", rest))
            else:
                o = range_overlap(vafs, u)
                if o == "contained":
                    left, right = range_split_at(u, vafs[1])[0], range_split_at(u, vafs[2])[1]
                    out += [x for x in (left, right) if x is not None]
                elif o == "end":
                    x = range_split_at(u, vafs[2])[1]
                    if x is not None:
                        out.append(x)
                elif o == "start":
                    x = range_split_at(u, vafs[1])[0]
                    if x is not None:
                        out.append(x)
                elif o == "none":
                    out.append(u)
    if not out:
        return ("atom", sample, ("set", ()))
    return ("or", [("atom", sample, sp) for sp in out])


def apply_negations(f, scn, contig):  # formula.rs:787-836
    k = f[0]
    if k == "not":
        return apply_negations(negate(f[1], scn, contig), scn, contig)
    if k in ("and", "or"):
        return (k, [apply_negations(o, scn, contig) for o in f[1]])
    if k == "expr":
        raise ValueError("expressions must be expanded before negations are applied")
    return f


def _products(f):
    """Sum of products of a negation-free formula: list of products, a product = tuple of distinct
    terminals in first-seen order; terminals that are false kill their product."""
    k = f[0]
    if is_terminal_false(f):
        return []
    if k == "or":
        out = []
        for o in f[1]:
            for p in _products(o):
                if p not in out:
                    out.append(p)
        return out
    if k == "and":
        acc = [()]
        for o in f[1]:
            nxt = []
            for a in acc:
                for p in _products(o):
                    q = a + tuple(t for t in p if t not in a)
                    nxt.append(q)
            acc = nxt
        return acc
    return [(f,)]


def simplify(f):
    """Formula::simplify (formula.rs:624-628) for negation-free formulas: the prime implicants of
    the monotone function (see the module docstring)."""
    prods = _products(f)
    sets = [frozenset(p) for p in prods]
    keep = []
    for i, p in enumerate(prods):
        if any((sets[j] < sets[i]) or (sets[j] == sets[i] and j < i) for j in range(len(prods)) if j != i):
            continue
        keep.append(p)
    if not keep:
        return FALSE
    terms = [p[0] if len(p) == 1 else ("and", list(p)) for p in keep]
    return terms[0] if len(terms) == 1 else ("or", terms)


def _merge_conj(a, b):  # FormulaTerminal::merge_conjunctions, formula.rs:102-136
    if a[0] == "range" and b[0] == "range":
        return range_and(a, b)
    if a[0] == "range":
        return ("set", tuple(v for v in b[1] if range_contains(a, v)))
    if b[0] == "range":
        return ("set", tuple(v for v in a[1] if range_contains(b, v)))
    return ("set", tuple(v for v in a[1] if v in b[1]))


def _try_merge_disj(a, b):  # FormulaTerminal::try_merge_disjunction, formula.rs:141-177
    if a[0] == "range" and b[0] == "range":
        return None if range_overlap(a, b) == "none" else range_or(a, b)
    if a[0] == "range":
        return a if all(range_contains(a, v) for v in b[1]) else None
    if b[0] == "range":
        return b if all(range_contains(b, v) for v in a[1]) else None
    return ("set", tuple(sorted(set(a[1]) | set(b[1]))))


def merge_atoms(f):
    """Formula::merge_atoms (formula.rs:503-607): only the TOP-LEVEL operands are grouped by sample;
    operands that are not atoms are left as they are (the reference does not recurse)."""
    k = f[0]
    if k not in ("and", "or"):
        return f
    groups, rest = {}, []
    for o in f[1]:
        if o[0] == "atom":
            groups.setdefault(o[1], []).append(o)
        else:
            rest.append(o)
    out = []
    for sample, stmts in groups.items():
        if k == "and":
            m = stmts[-1][2]                      # statements.pop(), then merge the others into it
            for s in stmts[:-1]:
                m = _merge_conj(m, s[2])
            out.append(("atom", sample, m))
        else:
            stmts = sorted(stmts, key=lambda s: (min(s[2][1]) if s[2][1] else -1.0) if s[2][0] == "set" else s[2][1])
            cur = stmts[0][2]
            for s in stmts[1:]:
                m = _try_merge_disj(cur, s[2])
                if m is not None:
                    cur = m
                else:
                    out.append(("atom", sample, cur))
                    cur = s[2]
            out.append(("atom", sample, cur))
    return (k, out + rest)


def strip_false(f):  # formula.rs:609-622
    if f[0] != "or":
        return f
    keep = [o for o in f[1] if not is_terminal_false(o) and
            not (o[0] == "and" and any(is_terminal_false(x) for x in o[1]))]
    return ("or", keep)


def _merge_products(f):
    """Atoms of one sample inside a product of a sum of products are intersected; contradictory
    products are dropped.  NOT in the reference: its merge_atoms only looks at top-level operands,
    which leaves e.g. `germline:]0.0,0.1] & 18_D:0.0 & germline:0.0` (testcase test58) as a product
    that binds a sample twice -- a tree path that neither the reference's VecMap of base events (the
    inner node overwrites the outer one) nor vlr_model can express.  What the reference's BDD makes
    of such formulas cannot be known here (crate not vendored); this is the meaning of the formula."""
    if f[0] != "or":
        return f
    out = []
    for o in f[1]:
        if o[0] == "and":
            merged, seen = [], {}
            for t in o[1]:
                if t[0] == "atom" and t[1] in seen:
                    k = seen[t[1]]
                    merged[k] = ("atom", t[1], _merge_conj(merged[k][2], t[2]))
                else:
                    if t[0] == "atom":
                        seen[t[1]] = len(merged)
                    merged.append(t)
            if any(is_terminal_false(t) or (t[0] == "atom" and t[2] == EMPTY_RANGE) for t in merged):
                continue
            o = merged[0] if len(merged) == 1 else ("and", merged)
        if o not in out:
            out.append(o)
    if not out:
        return FALSE
    return out[0] if len(out) == 1 else ("or", out)


def normalize(f, scn, contig="all"):
    """Formula::normalize (formula.rs:391-401), then _merge_products."""
    g = apply_negations(expand_expressions(f, scn), scn, contig)
    g = simplify(merge_atoms(simplify(g)))
    return _merge_products(strip_false(g))


def canonical(f):
    """Order-insensitive form of a normalised formula (operands sorted), for comparisons."""
    if f[0] in ("and", "or"):
        return (f[0], tuple(sorted((canonical(o) for o in f[1]), key=repr)))
    return f


def validate(scn, contig="all"):
    """Scenario::validate (grammar/mod.rs:223-271): two events whose union is itself a third event
    are reported as overlapping."""
    norm = {}
    for name, f in scn.events.items():
        if name != "absent":
            norm.setdefault(canonical(normalize(f, scn, contig)), []).append(name)
    forms = sorted(norm, key=repr)
    overlapping = []
    for e1, e2 in itertools.combinations(forms, 2):
        if e1 == FALSE or e2 == FALSE:
            continue
        def plain(c):
            return (c[0], [plain(o) for o in c[1]]) if c[0] in ("and", "or") else c
        d = canonical(normalize(("or", [plain(e1), plain(e2)]), scn, contig))
        if d in norm:
            overlapping.append((norm[e1], norm[e2], norm[d]))
    if overlapping:
        raise ValueError("overlapping events: " + ", ".join("(%s | %s) = %s" % o for o in overlapping))


# ----------------------------------------------------------------------------- VAF tree
class Node:
    __slots__ = ("kind", "sample", "vafs", "variant", "children")

    def __init__(self, kind, sample=None, vafs=None, variant=None):
        self.kind, self.sample, self.vafs, self.variant, self.children = kind, sample, vafs, variant, []

    def clone(self):
        n = Node(self.kind, self.sample, self.vafs, self.variant)
        n.children = [c.clone() for c in self.children]
        return n

    def leafs(self):
        if not self.children:
            return [self]
        return [l for c in self.children for l in c.leafs()]


def vaftree(norm, scn, contig="all"):
    """VAFTree::new (vaftree.rs:100-214): list of root nodes."""
    def build(f):
        k = f[0]
        if k == "atom":
            i = scn.idx(f[1])
            if i is None:
                raise ValueError("invalid sample name %s" % f[1])
            return [Node("sample", i, f[2])]
        if k == "or":
            return [t for o in f[1] for t in build(o)]
        if k == "and":
            ops = sorted(f[1], key=lambda o: 1 if o[0] == "or" else 0)  # stable: disjunctions last
            roots = build(ops[0])
            for o in ops[1:]:
                sub = build(o)
                for r in roots:
                    for leaf in r.leafs():
                        leaf.children = [s.clone() for s in sub]
            return roots
        if k == "variant":
            return [Node("variant", variant=(f[1], f[2], f[3]))]
        if k == "false":
            return [Node("false")]
        raise ValueError("formula is not normalised: %r" % (f,))

    def add_missing(node, seen):
        if node.kind == "false":
            return
        if node.kind == "sample":
            seen.add(node.sample)
        if not node.children:
            for name in scn.samples:
                i = scn.idx(name)
                if i not in seen:
                    seen.add(i)
                    node.children = [Node("sample", i, sp) for sp in scn.contig_universe(name, contig)]
                    add_missing(node, seen)
                    break
        else:
            for c in node.children[1:]:
                add_missing(c, set(seen))
            add_missing(node.children[0], seen)

    roots = build(norm)
    for r in roots:
        add_missing(r, set())
    return roots


def absent_tree(n_samples):  # VAFTree::absent, vaftree.rs:16-39
    root = cur = Node("sample", 0, vset([0.0]))
    for s in range(1, n_samples):
        nxt = Node("sample", s, vset([0.0]))
        cur.children = [nxt]
        cur = nxt
    return [root]


# ----------------------------------------------------------------------------- flattening
def flatten(scn, contig="all", snv=None, snv_or_mnv=True, omit_strand_bias=False,
            omit_read_orientation_bias=False, omit_read_position_bias=False, omit_softclip_bias=False,
            omit_divindel_bias=False, min_divindel_other_rate=0.25):
    """Scenario -> the dict layout of scenario.py / vlr_model for one contig and one candidate.

    snv: (refbase, altbase) of an SNV candidate or None; Variant nodes are resolved against it the
    way GenericPosterior::density does (generic.rs:211-233): a contradicted node becomes FALSE, a
    satisfied one is skipped.  Event universe and bias sets: calling.rs:398-412, 466-504."""
    names = scn.sample_names
    nodes = []

    def emit(node):
        kids = [emit(c) for c in node.children]
        if node.kind == "false":
            d = dict(kind=_scn.NODE_FALSE, sample=0, vafs=[], children=[])
        elif node.kind == "variant":
            positive, refb, altb = node.variant
            if snv is not None:
                contains = iupac_contains(refb, snv[0]) and iupac_contains(altb, snv[1])
                dead = (positive and not contains) or (not positive and contains)
            else:
                dead = positive
            d = dict(kind=_scn.NODE_FALSE if dead else _scn.NODE_SKIP, sample=0, vafs=[],
                     children=[] if dead else kids)
        elif node.vafs[0] == "set":
            d = dict(kind=_scn.NODE_SET, sample=node.sample, vafs=list(node.vafs[1]), children=kids)
        else:
            _, s, e, lex, rex = node.vafs
            d = dict(kind=_scn.NODE_RANGE, sample=node.sample, vafs=[s, e, float(lex), float(rex)], children=kids)
        nodes.append(d)
        return len(nodes) - 1

    arts = _scn.artifact_biases(not omit_strand_bias, snv_or_mnv and not omit_read_orientation_bias,
                                snv_or_mnv and not omit_read_position_bias,
                                snv_or_mnv and not omit_softclip_bias, not omit_divindel_bias,
                                min_divindel_other_rate)
    biases = [_scn.bias()] + arts
    validate(scn, contig)
    events = [dict(name="absent", roots=[emit(r) for r in absent_tree(len(names))], biases=[0])]
    for name, formula in scn.events.items():
        roots = [emit(r) for r in vaftree(normalize(formula, scn, contig), scn, contig)]
        events.append(dict(name=name, roots=roots, biases=[0]))
        if arts:
            events.append(dict(name=name, roots=roots, biases=list(range(1, len(biases)))))
    contaminated, by, purity = [], [], []
    for n in names:
        c = scn.samples[n].contamination
        if c:
            if scn.idx(c["by"]) is None:
                raise ValueError("invalid sample name %s" % c["by"])
            contaminated.append(1); by.append(scn.idx(c["by"])); purity.append(1.0 - float(c["fraction"]))  # float(): PyYAML reads `1e-3` as str
        else:
            contaminated.append(0); by.append(0); purity.append(1.0)
    return dict(n_samples=len(names), resolution=[scn.samples[n].resolution for n in names],
                contaminated=contaminated, by=by, purity=purity, nodes=nodes, events=events, biases=biases,
                sample_names=names)


def format_normalized(f):
    """Display of NormalizedFormula (formula.rs:861-915), for logs and tests."""
    k = f[0]
    if k == "atom":
        sp = f[2]
        if sp[0] == "set":
            if len(sp[1]) == 1:
                return "%s:%s" % (f[1], repr(sp[1][0]))
            return "false" if not sp[1] else "%s:{%s}" % (f[1], ", ".join("%.1f" % v for v in sp[1]))
        return "%s:%s%.1f,%.1f%s" % (f[1], "]" if sp[3] else "[", sp[1], sp[2], "[" if sp[4] else "]")
    if k == "variant":
        return "%s(%s>%s)" % ("" if f[1] else "!", f[2], f[3])
    if k == "false":
        return "false"
    sep = " & " if k == "and" else " | "
    return sep.join(format_normalized(o) if o[0] in ("atom", "variant") else "(%s)" % format_normalized(o)
                    for o in f[1])
