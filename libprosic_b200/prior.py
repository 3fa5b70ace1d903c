"""Host-side restatement of the reference's `Prior` (src/variants/model/prior.rs) for the cases the
fixture runner meets: uniform samples, germline samples (ploidy + species heterozygosity), Mendelian
and clonal inheritance, somatic mutation rates, with the variant-type fractions.  Subclonal
inheritance is not restated (NotImplementedError; the reference itself warns that it "will likely
yield wrong results", prior.rs:511).

The prior is data independent: the GPU path takes it as a table over the leaf allele-frequency
vectors that `vlr_model_leaves` enumerates (include/vlr_gpu.h); this module evaluates that table.

  * Prior::calc_prob                      prior.rs:288-431
  * prob_population_germline              prior.rs:628-656
  * prob_mendelian_alt_counts/inheritance prior.rs:658-786 (statrs Hypergeometric pmf)
  * prob_somatic_mutation                 prior.rs:433-455 (density rate / (vaf^2 * genome size))
  * prob_clonal_inheritance               prior.rs:457-500
  * variant-type fractions                grammar/mod.rs:373-420, prior.rs:242-260
"""
import math

from . import grammar as G

LN_ZERO = -math.inf
SOMATIC_EPSILON = 0.0001   # prior.rs:36


def _ln(x):
    return math.log(x) if x > 0.0 else LN_ZERO


def _ln_sum_exp(ps):
    ps = [p for p in ps if p != LN_ZERO]
    if not ps:
        return LN_ZERO
    m = max(ps)
    k = ps.index(m)
    return m + math.log1p(sum(math.exp(p - m) for i, p in enumerate(ps) if i != k))


def _ln_one_minus_exp(p):
    if p > 0.0:
        raise ValueError("probability above 1")
    if p == 0.0:
        return LN_ZERO
    return math.log1p(-math.exp(p)) if p < -0.693 else math.log(-math.expm1(p))


def _hypergeom_pmf(N, K, n, k):
    """statrs Hypergeometric::new(N, K, n).pmf(k)."""
    if k > K or n - k > N - K or k > n or k < 0:
        return 0.0
    return math.comb(K, k) * math.comb(N - K, n - k) / math.comb(N, n)


def variant_type_fraction(species, kind):
    """VariantTypeFraction::get (grammar/mod.rs:399-409); kind: 'snv' | 'ins' | 'del' | 'mnv' | 'sv'."""
    f = dict(indel=0.0125, mnv=0.001, sv=0.01)
    f.update((species or {}).get("variant-fractions") or {})
    return {"ins": f["indel"], "del": f["indel"], "mnv": f["mnv"], "sv": f["sv"]}.get(kind, 1.0)


def _setup(scn, contig, kind):
    """Everything Prior::new + set_universe_and_ploidies + set_variant_type fix for one contig and
    variant type: per-sample ploidies, universes, rates (times the variant-type fraction), inheritance."""
    sp = scn.species or {}
    names = scn.sample_names
    frac = variant_type_fraction(sp, kind)
    def num(x):   # PyYAML reads `1e-3` (no dot) as a string, serde_yaml as a float
        return None if x is None else float(x)
    het = num(sp.get("heterozygosity"))
    ln_het = None if het is None else math.log(het * frac)            # vartype_heterozygosity
    uniform = [scn.samples[n].universe is not None for n in names]
    ploidy = [scn.contig_ploidy(n, contig) for n in names]
    universe = [scn.contig_universe(n, contig) if u else None for n, u in zip(names, uniform)]
    genome_size = num(sp.get("genome-size"))
    germ_rate, som_rate, inherit = [], [], []
    for n in names:
        s = scn.samples[n]
        r = num(s.germline_rate if s.germline_rate is not None else sp.get("germline-mutation-rate"))
        germ_rate.append(None if r is None else r * frac)
        r = num(s.somatic_rate if s.somatic_rate is not None else sp.get("somatic-effective-mutation-rate"))
        som_rate.append(None if r is None else r * frac)
        inh = s.inheritance
        if inh is None:
            inherit.append(None)
        elif "mendelian" in inh:
            a, b = inh["mendelian"]["from"]
            inherit.append(("mendelian", scn.idx(a), scn.idx(b)))
        elif "clonal" in inh:
            inherit.append(("clonal", scn.idx(inh["clonal"]["from"]), bool(inh["clonal"]["somatic"])))
        else:
            raise NotImplementedError("inheritance pattern %s" % list(inh))
    return names, het, ln_het, uniform, ploidy, universe, genome_size, germ_rate, som_rate, inherit


def make_prior(scn, contig, kind="snv", want_program=False):
    """ln prior of an allele-frequency vector (sample order of the scenario), Prior::compute.
    want_program: return (f, program) where program is the dict form of vlr_prior_prog
    (include/vlr_gpu.h) that lets the device evaluate the same prior at every leaf."""
    names, het, ln_het, uniform, ploidy, universe, genome_size, germ_rate, som_rate, inherit = \
        _setup(scn, contig, kind)

    def mendelian_alt_counts(sp_, tp, sa, ta, rate):  # prob_mendelian_alt_counts
        def cases(p):
            return [p // 2] if p % 2 == 0 else [p // 2, p // 2 + 1]
        pairs = [(c1, c2) for c1 in cases(sp_[0]) for c2 in cases(sp_[1])]
        if not any(c1 + c2 == tp for c1, c2 in pairs):
            raise ValueError("ploidies of child and parents do not match (%d, %d => %d)" % (sp_[0], sp_[1], tp))
        terms = []
        for c1, c2 in pairs:
            if c1 + c2 != tp:
                continue
            for a1 in range(0, min(sa[0], c1) + 1):
                for a2 in range(0, min(sa[1], c2) + 1):
                    if a1 + a2 <= ta:
                        p = _ln(_hypergeom_pmf(sp_[0], sa[0], c1, a1)) + _ln(_hypergeom_pmf(sp_[1], sa[1], c2, a2))
                        terms.append(p + math.log(rate) * (ta - a1 - a2))
        return _ln_sum_exp(terms)

    def somatic_mutation(rate, vaf):  # prob_somatic_mutation, prior.rs:433-455
        def density(v):
            return math.log(rate) - (2.0 * math.log(v) + math.log(genome_size))
        if abs(vaf) <= SOMATIC_EPSILON:
            # ln_simpsons_integrate_exp over [eps, 1] with 11 grid points, then ln(1 - .)
            n, lo, hi = 11, SOMATIC_EPSILON, 1.0
            hstep = (hi - lo) / (n - 1)
            terms = []
            for i in range(n):
                w = 0.0 if i in (0, n - 1) else (math.log(4.0) if i % 2 == 1 else math.log(2.0))
                terms.append(w + density(abs(lo + hstep * i)))
            return _ln_one_minus_exp(_ln_sum_exp(terms) + math.log(hstep) - math.log(3.0))
        return density(abs(vaf))

    def close(a, b):   # approx::relative_eq! with the crate's defaults
        return a == b or abs(a - b) <= max(abs(a), abs(b)) * 2.220446049250313e-16 or abs(a - b) <= 2.220446049250313e-16

    def end(afs, germ):   # the recursion end of calc_prob
        prob = 0.0
        if ln_het is not None:                                         # step 1: population
            pop = [s for s in range(len(afs)) if inherit[s] is None and ploidy[s] is not None and not uniform[s]]
            m = sum(int(round(ploidy[s] * germ[s])) for s in pop)
            if m > 0:
                prob = ln_het - math.log(m)
            else:
                n = sum(ploidy[s] for s in pop)
                prob = _ln_one_minus_exp(_ln_sum_exp([ln_het - math.log(k) for k in range(1, n + 1)]))
        eff = [afs[s] - germ[s] for s in range(len(afs))]              # effective_somatic_vaf
        for s, inh in enumerate(inherit):                              # step 2: inheritance, somatic mutation
            if uniform[s]:
                continue
            if inh is None:
                if som_rate[s] is not None:
                    prob += somatic_mutation(som_rate[s], eff[s])
            elif inh[0] == "mendelian":
                if germ_rate[s] is None:
                    raise ValueError("no germline mutation rate for child %s" % names[s])
                def n_alt(k):
                    return int(round(germ[k] * ploidy[k]))
                prob += mendelian_alt_counts((ploidy[inh[1]], ploidy[inh[2]]), ploidy[s],
                                             (n_alt(inh[1]), n_alt(inh[2])), n_alt(s), germ_rate[s])
                if som_rate[s] is not None:
                    prob += somatic_mutation(som_rate[s], eff[s])
            else:                                                      # clonal, prior.rs:457-500
                parent, somatic = inh[1], inh[2]
                if not close(germ[s], germ[parent]):
                    return LN_ZERO
                if somatic and som_rate[s] is not None:
                    prob += somatic_mutation(som_rate[s], afs[s] - germ[s] - eff[parent])
                elif somatic:
                    if not close(eff[s], eff[parent]):
                        return LN_ZERO
                elif som_rate[s] is not None:
                    prob += somatic_mutation(som_rate[s], eff[s])
            if prob == LN_ZERO:
                return LN_ZERO
        return prob

    def rec(afs, germ):   # the recursion of calc_prob over the samples
        s = len(germ)
        if s == len(afs):
            return end(afs, germ)
        af = afs[s]
        if ploidy[s] == 0 and af != 0.0:
            return LN_ZERO
        if uniform[s]:
            if not any(G.spectrum_contains(u, af) for u in universe[s]):
                return LN_ZERO
            return rec(afs, germ + [0.0])
        if som_rate[s] is not None:
            if ploidy[s] is None:
                raise ValueError("sample %s with a somatic mutation rate but no ploidy" % names[s])
            return _ln_sum_exp([rec(afs, germ + [(k / ploidy[s]) if ploidy[s] > 0 else 0.0])
                                for k in range(ploidy[s] + 1)])
        if ploidy[s] is not None and het is not None:
            n_alt = ploidy[s] * af
            if not math.isclose(n_alt, round(n_alt), rel_tol=1e-9, abs_tol=1e-9):  # is_valid_germline_vaf
                return LN_ZERO
            return rec(afs, germ + [af])
        raise ValueError("not enough information for the prior of sample %s" % names[s])

    cache = {}

    def f(afs):
        key = tuple(afs)
        if key not in cache:
            cache[key] = rec(list(afs), [])
        return cache[key]
    if not want_program:
        return f

    # ---- the same prior as a program for the device (vlr_prior_prog).  The allele-frequency
    # independent part of the recursion end is tabulated per germline alt-count vector.
    ns = len(names)
    pkind, som_kind, parent, ln_rate, som_zero = [], [], [], [], []
    for s_ in range(ns):
        if uniform[s_]:
            pkind.append(0)
        elif som_rate[s_] is not None:
            if ploidy[s_] is None:
                raise ValueError("sample %s with a somatic mutation rate but no ploidy" % names[s_])
            pkind.append(1)
        elif ploidy[s_] is not None and het is not None:
            pkind.append(2)
        else:
            raise ValueError("not enough information for the prior of sample %s" % names[s_])
        inh, sk, par = inherit[s_], 0, -1
        if not uniform[s_]:
            if inh is None or inh[0] == "mendelian":
                sk = 1 if som_rate[s_] is not None else 0
            else:
                par = inh[1]
                if inh[2] and som_rate[s_] is not None:
                    sk = 2
                elif inh[2]:
                    sk = 3
                elif som_rate[s_] is not None:
                    sk = 1
        som_kind.append(sk)
        parent.append(par)
        ln_rate.append(math.log(som_rate[s_]) if som_rate[s_] is not None else 0.0)
        som_zero.append(somatic_mutation(som_rate[s_], 0.0) if sk in (1, 2) else 0.0)
    radix = [1 if (pkind[s_] == 0 or not ploidy[s_]) else ploidy[s_] + 1 for s_ in range(ns)]
    n_germ = 1
    for r in radix:
        n_germ *= r
    germ_ln = []
    for idx in range(n_germ):
        rem, counts = idx, [0] * ns
        for s_ in range(ns - 1, -1, -1):
            counts[s_] = rem % radix[s_]
            rem //= radix[s_]
        germ = [(counts[s_] / ploidy[s_]) if radix[s_] > 1 else 0.0 for s_ in range(ns)]
        prob = 0.0
        if ln_het is not None:
            pop = [s_ for s_ in range(ns) if inherit[s_] is None and ploidy[s_] is not None and not uniform[s_]]
            m = sum(int(round(ploidy[s_] * germ[s_])) for s_ in pop)
            if m > 0:
                prob = ln_het - math.log(m)
            else:
                n = sum(ploidy[s_] for s_ in pop)
                prob = _ln_one_minus_exp(_ln_sum_exp([ln_het - math.log(k) for k in range(1, n + 1)]))
        for s_, inh in enumerate(inherit):
            if uniform[s_] or inh is None:
                continue
            if inh[0] == "mendelian":
                if germ_rate[s_] is None:
                    raise ValueError("no germline mutation rate for child %s" % names[s_])
                prob += mendelian_alt_counts((ploidy[inh[1]], ploidy[inh[2]]), ploidy[s_],
                                             (int(round(germ[inh[1]] * ploidy[inh[1]])),
                                              int(round(germ[inh[2]] * ploidy[inh[2]]))),
                                             int(round(germ[s_] * ploidy[s_])), germ_rate[s_])
            elif not close(germ[s_], germ[inh[1]]):
                prob = LN_ZERO
            if prob == LN_ZERO:
                break
        germ_ln.append(prob)
    uni = []
    for s_ in range(ns):
        ent = []
        for spc in (universe[s_] or []):
            if spc[0] == "set":
                ent += [(0.0, float(v), float(v), 0.0, 0.0) for v in sorted(spc[1])]
            else:
                ent.append((1.0, float(spc[1]), float(spc[2]), float(bool(spc[3])), float(bool(spc[4]))))
        uni.append(ent)
    program = dict(kind=pkind, ploidy=[-1 if p_ is None else int(p_) for p_ in ploidy], som_kind=som_kind,
                   parent=parent, ln_rate=ln_rate, som_zero=som_zero,
                   ln_genome_size=math.log(genome_size) if genome_size else 0.0, uni=uni, germ_ln=germ_ln)
    return f, program


def eval_program(prog, afs):
    """Host evaluation of a prior program (the dict form of vlr_prior_prog), operation for operation
    what the device does (csrc/posterior_kernel.cuh: prior_prog_eval).  Used by the tests as the
    callback of the oracle for random programs and to check make_prior's program against its closure."""
    ns = len(afs)
    radix = [1 if (prog["kind"][s] == 0 or prog["ploidy"][s] <= 0) else prog["ploidy"][s] + 1 for s in range(ns)]
    stride, st = [0] * ns, 1
    for s in range(ns - 1, -1, -1):
        stride[s] = st
        st *= radix[s]
    germ, fixed, branch = [0.0] * ns, 0, []
    for s in range(ns):
        af = afs[s]
        if prog["ploidy"][s] == 0 and af != 0.0:
            return LN_ZERO
        if prog["kind"][s] == 0:
            ok = False
            for k, a, b, lex, rex in prog["uni"][s]:
                if k == 0.0:
                    ok = ok or af == a
                else:
                    ok = ok or ((af > a if lex else af >= a) and (af < b if rex else af <= b))
            if not ok:
                return LN_ZERO
        elif prog["kind"][s] == 2:
            x = prog["ploidy"][s] * af
            r = float(round(x))
            if abs(x - r) > max(1e-9 * max(abs(x), abs(r)), 1e-9):
                return LN_ZERO
            fixed += int(r) * stride[s]
            germ[s] = af
        elif radix[s] > 1:
            branch.append(s)
    total = 1
    for s in branch:
        total *= radix[s]
    terms = []
    for combo in range(total):
        idx, rem = fixed, combo
        for s in reversed(branch):
            k = rem % radix[s]
            rem //= radix[s]
            idx += k * stride[s]
            germ[s] = k / prog["ploidy"][s]
        t = prog["germ_ln"][idx]
        if t == LN_ZERO:
            continue
        for s in range(ns):
            sk = prog["som_kind"][s]
            if sk == 0 or t == LN_ZERO:
                continue
            eff = afs[s] - germ[s]
            v = eff
            if sk != 1:
                par = prog["parent"][s]
                effp = afs[par] - germ[par]
                if sk == 3:
                    d = abs(eff - effp)
                    if not (eff == effp or d <= 2.220446049250313e-16 or d <= max(abs(eff), abs(effp)) * 2.220446049250313e-16):
                        t = LN_ZERO
                    continue
                v = afs[s] - germ[s] - effp
            v = abs(v)
            t += prog["som_zero"][s] if v <= SOMATIC_EPSILON else prog["ln_rate"][s] - (2.0 * math.log(v) + prog["ln_genome_size"])
        if t != LN_ZERO:
            terms.append(t)
    return _ln_sum_exp(terms)
