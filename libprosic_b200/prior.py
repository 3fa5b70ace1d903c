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


def make_prior(scn, contig, kind="snv"):
    """ln prior of an allele-frequency vector (sample order of the scenario), Prior::compute."""
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
    return f
