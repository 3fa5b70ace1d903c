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
