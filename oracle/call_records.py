"""CPU restatement (numpy) of the numeric fields of the final call record -- TEST INFRASTRUCTURE.

Follows src/calling/variants/calling.rs:573-602 (event posteriors, the artifact event) and
src/calling/variants/mod.rs:185-187, 309-333 (AF as f32, DP = expected_depth, PROB_* as the
absolute PHRED value in f32, float-missing for sites without any observation);
expected_depth: src/variants/evidence/observation.rs:31-36."""
import numpy as np

BCF_FLOAT_MISSING = np.array([0x7F800001], np.uint32).view(np.float32)[0]


def ln_sum_exp(ps):
    """LogProb::ln_sum_exp: first maximum + ln_1p(sum of the others)."""
    ps = list(ps)
    if not ps:
        return -np.inf
    imax = int(np.argmax(ps))
    pmax = ps[imax]
    if pmax == -np.inf:
        return -np.inf
    acc = 0.0
    for k, p in enumerate(ps):
        if k != imax and p != -np.inf:
            acc += np.exp(p - pmax)
    return pmax + np.log1p(acc)


def call_record(event_ln, marginal, map_af, is_artifact, prob_mapping_per_sample):
    """One site.  Returns (prob f32 [n_named + 1], af f32 [n_samples], dp int32 [n_samples])."""
    named = [e for e, a in enumerate(is_artifact) if not a]
    dp = np.array([np.uint32(np.round(np.exp(ln_sum_exp(pm)))) for pm in prob_mapping_per_sample], np.int32)
    af = np.asarray(map_af, np.float64).astype(np.float32)
    if all(len(pm) == 0 for pm in prob_mapping_per_sample):
        return np.full(len(named) + 1, BCF_FLOAT_MISSING, np.float32), af, dp
    post = np.asarray(event_ln, np.float64) - marginal
    k = -10.0 / np.log(10.0)
    prob = [np.float32(abs(post[e] * k)) for e in named]
    art = ln_sum_exp([post[e] for e, a in enumerate(is_artifact) if a])
    prob.append(np.float32(abs(art * k)))
    return np.array(prob, np.float32), af, dp


# ---- OBS / SOBS strings and bias flags (src/calling/variants/mod.rs:147-260, utils/mod.rs:101-146)
def _letter(prob_alt, prob_ref):
    """utils::bayes_factor_to_letter(BayesFactor::new(prob_alt, prob_ref)), Kass-Raftery scale."""
    bf = np.exp(prob_alt - prob_ref)
    if bf <= 1.0:
        d = abs(bf - 1.0)
        return "E" if (bf == 1.0 or d <= 2.220446049250313e-16 or d <= max(abs(bf), 1.0) * 2.220446049250313e-16) else "N"
    return "B" if bf <= 3.0 else ("P" if bf <= 20.0 else ("S" if bf <= 150.0 else "V"))


def obs_items(prob_alt, prob_ref, prob_mapping, cat):
    """(OBS item, SOBS item) of every observation row."""
    out = []
    for pa, pr, pm, c in zip(prob_alt, prob_ref, prob_mapping, cat):
        c = int(c)
        ltr = _letter(pa, pr)
        ltr = ltr.lower() if pm < np.log(0.95) else ltr
        strand = {2: "*", 1: "-", 0: "+"}[c & 3]
        orr = (c >> 2) & 15
        orient = ">" if orr == 0 else ("<" if orr == 1 else ("*" if orr == 8 else "!"))
        item = ltr + ("p" if (c >> 10) & 1 else "s") + strand + orient + ("*" if (c >> 6) & 1 else "^") + \
            ("$" if (c >> 7) & 1 else ".") + {0: "*", 1: "#", 2: "."}[(c >> 8) & 3]
        out.append((item, ltr))
    return out


def generalized_cigar(items):
    """utils::generalized_cigar(items, keep_order=false, aux_sort): counts per distinct item, most
    common first, then stably E.. after the others and N.. last.  Returns the list of (count, item);
    the reference's order among equal counts is unspecified (Counter over a HashMap), so callers
    compare the multiset per (class, count)."""
    from collections import Counter
    cnt = Counter(items)
    lst = sorted(cnt.items(), key=lambda kv: -kv[1])
    lst = sorted(lst, key=lambda kv: 2 if kv[0].startswith("N") else (1 if kv[0].startswith("E") else 0))
    return [(n, it) for it, n in lst]


def bias_chars(b):
    return ("+" if b["strand"] == 1 else "-" if b["strand"] == 2 else ".") + \
        (">" if b["orientation"] == 1 else "<" if b["orientation"] == 2 else ".") + \
        ("^" if b["read_position"] else ".") + ("$" if b["softclip"] else ".") + ("#" if b["divindel"] else ".")
