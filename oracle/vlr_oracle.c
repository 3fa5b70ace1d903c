/*
 * vlr_oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of Varlociraptor 4.1.0's
 * realignment + pileup-likelihood hot path.  See vlr_oracle.h for the scope statement and
 * the "PARITY UNPINNED" note.  Compile with -O2 -ffp-contract=off (oracle/Makefile) so that
 * the arithmetic is plain IEEE fp64 in the order written here.
 */
#include "vlr_oracle.h"
#include "detmath.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-INFINITY)

/* =====================================================================================
 * bio::stats::probs
 * ===================================================================================== */

/* the oracle's libm (detmath.h), exported for tests/test_detmath.py: 0 exp, 1 expm1, 2 log, 3 log1p */
void vlro_ref_math(int fn, long n, const double *x, double *out) {
    for (long k = 0; k < n; ++k)
        out[k] = fn == 0 ? vlro_exp(x[k]) : (fn == 1 ? vlro_expm1(x[k]) : (fn == 2 ? vlro_log(x[k]) : vlro_log1p(x[k])));
}

/* LogProb::ln_add_exp */
double vlro_ln_add_exp(double a, double b) {
    double p0 = a, p1 = b;
    if (p1 > p0) { double t = p0; p0 = p1; p1 = t; }
    if (p0 == NEG_INF) return NEG_INF;
    if (p0 == INFINITY) return INFINITY;
    if (p1 == NEG_INF) return p0;
    return p0 + vlro_log1p(vlro_exp(p1 - p0));
}

/* LogProb::ln_sum_exp: first max on ties, inner sum in slice order */
double vlro_ln_sum_exp(const double *p, size_t n) {
    if (n == 0) return NEG_INF;
    double pmax = p[0];
    size_t imax = 0;
    for (size_t i = 1; i < n; ++i)
        if (p[i] > pmax) { pmax = p[i]; imax = i; }
    if (pmax == NEG_INF) return NEG_INF;
    if (pmax == INFINITY) return INFINITY;
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) {
        if (i == imax || p[i] == NEG_INF) continue;
        s += vlro_exp(p[i] - pmax);
    }
    return pmax + vlro_log1p(s);
}

/* LogProb::ln_one_minus_exp */
double vlro_ln_one_minus_exp(double p) {
    if (p < -0.693) return vlro_log1p(-vlro_exp(p));
    return vlro_log(-vlro_expm1(p));
}

/* LogProb::cap_numerical_overshoot; NaN marks the reference's panic */
double vlro_cap_numerical_overshoot(double p, double eps) {
    if (p <= 0.0) return p;
    if (p <= eps) return 0.0;
    return NAN;
}

/* =====================================================================================
 * src/variants/evidence/bases.rs:23-48 -- base quality LUTs
 * LogProb::from(PHREDProb(q)) = q * (-ln10/10)
 * ===================================================================================== */
static const double PHRED_TO_LOG_FACTOR = -0.23025850929940456; /* -ln(10)/10 */

double vlro_ln_prob_miscall(uint8_t qual) { return (double)qual * PHRED_TO_LOG_FACTOR; }
double vlro_ln_prob_call(uint8_t qual) { return vlro_ln_one_minus_exp(vlro_ln_prob_miscall(qual)); }

/* ln(0.3333): pairhmm.rs:21-23, bases.rs:8-10 (quirk Q5: 0.3333, not 1/3) */
static double ln_confusion(void) { return vlro_log(0.3333); }

/* =====================================================================================
 * MiniLogProb (src/utils/mod.rs:381-407): f16 if p < -10 and floor(f16(p)) == floor(p), else f32.
 * half::f16::from_f64 is a direct round-to-nearest-even conversion.
 * ===================================================================================== */
uint16_t vlro_f64_to_f16_bits(double v) {
    uint64_t x;
    memcpy(&x, &v, 8);
    uint16_t sign = (uint16_t)((x >> 48) & 0x8000u);
    int64_t exp = (int64_t)((x >> 52) & 0x7FF);
    uint64_t man = x & 0xFFFFFFFFFFFFFull;
    if (exp == 0x7FF) { /* inf / nan */
        if (man == 0) return (uint16_t)(sign | 0x7C00u);
        return (uint16_t)(sign | 0x7C00u | 0x0200u | (uint16_t)(man >> 42));
    }
    int64_t e = exp - 1023; /* unbiased */
    if (e > 15) return (uint16_t)(sign | 0x7C00u); /* overflow -> inf */
    if (e >= -14) {
        /* normal half: 10 mantissa bits, drop 42 */
        uint64_t m = man >> 42;
        uint64_t rem = man & ((1ull << 42) - 1);
        uint64_t half = 1ull << 41;
        uint32_t h = (uint32_t)(((e + 15) << 10) | (int64_t)m);
        if (rem > half || (rem == half && (m & 1))) h += 1; /* may carry into exponent: correct */
        return (uint16_t)(sign | h);
    }
    if (e < -25) return sign; /* underflow to zero (|v| < 2^-25 rounds to 0) */
    /* subnormal half */
    uint64_t full = man | (1ull << 52);
    int shift = (int)(42 + (-14 - e));
    uint64_t m = full >> shift;
    uint64_t rem = full & ((1ull << shift) - 1);
    uint64_t half = 1ull << (shift - 1);
    if (rem > half || (rem == half && (m & 1))) m += 1;
    return (uint16_t)(sign | (uint16_t)m);
}

double vlro_f16_bits_to_f64(uint16_t h) {
    int sign = (h >> 15) & 1;
    int exp = (h >> 10) & 0x1F;
    int man = h & 0x3FF;
    double v;
    if (exp == 0) v = ldexp((double)man, -24);
    else if (exp == 31) v = man ? NAN : INFINITY;
    else v = ldexp((double)(man | 0x400), exp - 25);
    return sign ? -v : v;
}

double vlro_minilogprob_roundtrip(double p) {
    double proj = vlro_f16_bits_to_f64(vlro_f64_to_f16_bits(p));
    /* `proj.floor() as i64 == prob.floor() as i64` with Rust's saturating float->int casts */
    double fa = floor(proj), fb = floor(p);
    int same;
    if (isinf(fa) || isinf(fb)) {
        /* saturating cast: -inf -> i64::MIN, +inf -> i64::MAX */
        int64_t ia = isinf(fa) ? (fa < 0 ? INT64_MIN : INT64_MAX) : (int64_t)fa;
        int64_t ib = isinf(fb) ? (fb < 0 ? INT64_MIN : INT64_MAX) : (int64_t)fb;
        same = ia == ib;
    } else {
        same = fa == fb;
    }
    if (p < -10.0 && same) return proj;
    return (double)(float)p;
}

/* =====================================================================================
 * Pair-HMM forward: bio::stats::pairhmm::PairHMM::{new, prob_related} (Appendix A.2)
 * called from PairHMMRealigner::calculate_prob_allele, realignment/mod.rs:430-444.
 * Emissions: realignment/pairhmm.rs:123-211 (default_emission!, ReadEmission).
 * ===================================================================================== */

typedef struct {
    double gap_x, gap_y, gap_x_ext, gap_y_ext; /* ln */
    double no_gap;                             /* ln(1 - (e^gx + e^gy)) */
    double no_gap_x_ext, no_gap_y_ext;         /* ln(1 - e^gxe), ln(1 - e^gye) */
    double close_x_state, close_y_state;       /* constants paired with fx / fy in the M update */
} hmm_consts_t;

static void hmm_consts(const double gap_ln[4], unsigned flags, hmm_consts_t *c) {
    c->gap_x = gap_ln[0];
    c->gap_y = gap_ln[1];
    c->gap_x_ext = gap_ln[2];
    c->gap_y_ext = gap_ln[3];
    /* PairHMM::new; same table as PathHMMRealigner::new, realignment/mod.rs:465-476 */
    c->no_gap = vlro_ln_one_minus_exp(vlro_ln_add_exp(c->gap_x, c->gap_y));
    c->no_gap_x_ext = vlro_ln_one_minus_exp(c->gap_x_ext);
    c->no_gap_y_ext = vlro_ln_one_minus_exp(c->gap_y_ext);
    if (flags & VLRO_HMM_PATH_CLOSE) {
        /* PathHMM pairing: after Del (x-only state, extended with gap_y_ext) close with 1-gye;
         * after Ins (y-only state) close with 1-gxe (realignment/mod.rs:469-470,513-523) */
        c->close_x_state = c->no_gap_y_ext;
        c->close_y_state = c->no_gap_x_ext;
    } else {
        /* upstream recollection (Appendix A.2, ambiguity 1): fx pairs with no_gap_x_extend and
         * fy with no_gap_y_extend.  Identical whenever gxe == gye (CLI defaults 0/0). */
        c->close_x_state = c->no_gap_x_ext;
        c->close_y_state = c->no_gap_y_ext;
    }
}

/* bio pairhmm ln_sum3_exp_approx.  The crate source is not in /root/reference; two recollections
 * exist and both are implemented (flag, measured sensitivity in DESIGN.md section 5):
 *   VLRO_HMM_SUM3_APPROX3 (default of the library): sort descending; p0 - p1 > 10 -> p0;
 *       else p1 - p2 > 10 -> p0.ln_add_exp(p1); else the exact sum of three
 *   VLRO_HMM_SUM3_APPROX  (round-1 recollection): p0 - p1 > 10 -> p0, else the exact sum of three */
static double ln_sum3(double a, double b, double c, unsigned flags) {
    if (flags & (VLRO_HMM_SUM3_APPROX | VLRO_HMM_SUM3_APPROX3)) {
        double p0 = a, p1 = b, p2 = c, t;
        if (p1 < p2) { t = p1; p1 = p2; p2 = t; }
        if (p1 > p0) { t = p1; p1 = p0; p0 = t; }
        if (p2 > p1) { t = p2; p2 = p1; p1 = t; }
        /* p0 >= p1 >= p2 */
        if (p0 - p1 > 10.0) return p0;
        if ((flags & VLRO_HMM_SUM3_APPROX3) && p1 - p2 > 10.0) return vlro_ln_add_exp(p0, p1);
        double v[3] = {p0, p1, p2};
        return vlro_ln_sum_exp(v, 3);
    }
    double v[3] = {a, b, c};
    return vlro_ln_sum_exp(v, 3);
}

static inline uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

#define MED_MAX 0x3fffffff

double vlro_pairhmm_prob_related(const uint8_t *hap, int len_x, const uint8_t *read,
                                 const uint8_t *qual, int len_y, const double gap_ln[4],
                                 int max_edit_dist, unsigned flags) {
    hmm_consts_t c;
    hmm_consts(gap_ln, flags, &c);
    const int banded = max_edit_dist >= 0;
    const double lnconf = ln_confusion();
    size_t n = (size_t)len_y + 1;
    double *buf = (double *)malloc(sizeof(double) * n * 6);
    int *med = (int *)malloc(sizeof(int) * n * 2);
    double *any_miscall = (double *)malloc(sizeof(double) * n * 2);
    double *no_miscall = any_miscall + n;
    double *prob_cols = (double *)malloc(sizeof(double) * 3 * (size_t)(len_x > 0 ? len_x : 1));
    double *fm[2] = {buf, buf + n}, *fx[2] = {buf + 2 * n, buf + 3 * n},
           *fy[2] = {buf + 4 * n, buf + 5 * n};
    int *md[2] = {med, med + n};
    /* ReadEmission::new (pairhmm.rs:160-180) */
    for (int j = 0; j < len_y; ++j) {
        any_miscall[j] = vlro_ln_prob_miscall(qual[j]);
        no_miscall[j] = vlro_ln_one_minus_exp(any_miscall[j]);
    }
    for (int k = 0; k < 2; ++k)
        for (size_t j = 0; j < n; ++j) {
            fm[k][j] = fx[k][j] = fy[k][j] = NEG_INF;
            md[k][j] = MED_MAX;
        }
    int prev = 0, curr = 1;
    size_t ncols = 0;
    for (int i = 0; i < len_x; ++i) {
        /* semiglobal start: prob_start_gap_x(i) = ln 1, free_start_gap_x (pairhmm.rs:103-121) */
        fm[prev][0] = 0.0;
        md[prev][0] = 0;
        uint8_t xb = up(hap[i]);
        for (int j = 0; j < len_y; ++j) {
            int j_ = j + 1;
            int m_tl = md[prev][j], m_t = md[curr][j], m_l = md[prev][j_];
            if (banded) {
                int mn = m_tl < m_t ? m_tl : m_t;
                if (m_l < mn) mn = m_l;
                if (mn > max_edit_dist) continue; /* cell stays at its reset state */
            }
            int is_match = read[j] == xb;
            /* prob_emit_xy: pairhmm.rs:182-196 */
            double e_xy = is_match ? no_miscall[j] : (any_miscall[j] + lnconf);
            double m = e_xy + ln_sum3(c.no_gap + fm[prev][j], c.close_x_state + fx[prev][j],
                                      c.close_y_state + fy[prev][j], flags);
            /* x-only state (gap in y, "Del"): emit_x = ln 1 (pairhmm.rs:132-135) */
            double x = 0.0 + vlro_ln_add_exp(c.gap_y + fm[prev][j_], c.gap_y_ext + fx[prev][j_]);
            /* y-only state (gap in x, "Ins"): emit_y = any_miscall (pairhmm.rs:198-200) */
            double y = any_miscall[j] +
                       vlro_ln_add_exp(c.gap_x + fm[curr][j], c.gap_x_ext + fy[curr][j]);
            if (banded) {
                int a = is_match ? m_tl : (m_tl >= MED_MAX ? MED_MAX : m_tl + 1);
                int b = m_l >= MED_MAX ? MED_MAX : m_l + 1;
                int d = m_t >= MED_MAX ? MED_MAX : m_t + 1;
                int mn = a < b ? a : b;
                if (d < mn) mn = d;
                md[curr][j_] = mn;
            }
            fm[curr][j_] = m;
            fx[curr][j_] = x;
            fy[curr][j_] = y;
        }
        /* free_end_gap_x: collect the last row of every column */
        prob_cols[ncols++] = fm[curr][len_y];
        prob_cols[ncols++] = fx[curr][len_y];
        prob_cols[ncols++] = fy[curr][len_y];
        /* swap and reset the new curr column (oracle choice: reset all four arrays) */
        int t = prev; prev = curr; curr = t;
        if (flags & VLRO_HMM_RESET_FM_ONLY) {
            /* the other recollection of upstream ("reset next column to zeros": fm only); skipped
             * band cells then keep fx / fy / min_edit_dist from two columns back */
            for (size_t j = 0; j < n; ++j) fm[curr][j] = NEG_INF;
        } else {
            for (size_t j = 0; j < n; ++j) {
                fm[curr][j] = fx[curr][j] = fy[curr][j] = NEG_INF;
                md[curr][j] = MED_MAX;
            }
        }
    }
    double p = vlro_ln_sum_exp(prob_cols, ncols);
    free(buf); free(med); free(any_miscall); free(prob_cols);
    if (p > 0.0) p = 0.0; /* "sum of paths can exceed 1" cap */
    return p;
}

/* ---- brute force over explicit paths; validates the recurrence on tiny inputs ------- */
typedef struct {
    const uint8_t *hap, *read, *qual;
    int len_x, len_y;
    hmm_consts_t c;
    double lnconf;
    double total; /* linear-space accumulator */
} bf_t;

/* state: 0=M 1=X(del, consumes hap) 2=Y(ins, consumes read); i,j = consumed counts */
static void bf_rec(bf_t *b, int i, int j, int state, double lp) {
    if (j == b->len_y && state >= 0) b->total += vlro_exp(lp); /* free end: any column, last row */
    /* extend by M */
    if (i < b->len_x && j < b->len_y) {
        double tr = state < 0 ? b->c.no_gap : (state == 0 ? b->c.no_gap
                                               : (state == 1 ? b->c.close_x_state : b->c.close_y_state));
        double eps = vlro_ln_prob_miscall(b->qual[j]);
        double e = (b->read[j] == up(b->hap[i])) ? vlro_ln_one_minus_exp(eps) : eps + b->lnconf;
        bf_rec(b, i + 1, j + 1, 0, lp + tr + e);
    }
    /* X: consume hap only; from M (open) or X (extend).  Not from the start row (row 0 has fx=0). */
    if (i < b->len_x && j > 0 && j <= b->len_y && (state == 0 || state == 1)) {
        double tr = state == 0 ? b->c.gap_y : b->c.gap_y_ext;
        if (tr > NEG_INF) bf_rec(b, i + 1, j, 1, lp + tr);
    }
    /* Y: consume read only; from M (open) or Y (extend); fm[curr][0] = 0 so never from the start */
    if (j < b->len_y && (state == 0 || state == 2)) {
        double tr = state == 0 ? b->c.gap_x : b->c.gap_x_ext;
        if (tr > NEG_INF) bf_rec(b, i, j + 1, 2, lp + tr + vlro_ln_prob_miscall(b->qual[j]));
    }
}

double vlro_pairhmm_bruteforce(const uint8_t *hap, int len_x, const uint8_t *read,
                               const uint8_t *qual, int len_y, const double gap_ln[4],
                               unsigned flags) {
    bf_t b;
    b.hap = hap; b.read = read; b.qual = qual; b.len_x = len_x; b.len_y = len_y;
    hmm_consts(gap_ln, flags & ~VLRO_HMM_SUM3_APPROX, &b.c);
    b.lnconf = ln_confusion();
    b.total = 0.0;
    /* free start: the path may begin (state -1 = start, transition no_gap from fm[prev][0]=1)
     * at any hap offset */
    for (int s = 0; s < len_x; ++s) bf_rec(&b, s, 0, -1, 0.0);
    double p = vlro_log(b.total);
    if (p > 0.0) p = 0.0;
    return p;
}

/* =====================================================================================
 * Myers best hit: EditDistanceCalculation::calc_best_hit (edit_distance.rs:56-154) on top of
 * bio::pattern_matching::myers (Appendix A.3).  The bit-vector algorithm computes exactly the
 * semiglobal edit-distance matrix D[j][i] (pattern row j, text column i, D[0][i] = 0,
 * D[j][0] = j); the oracle keeps the explicit matrix so that the traceback rule can be read
 * off directly (default, rule 7 below):
 *   at (j,i):  if mismatch and D[j-1][i-1] + 1 == D[j][i]                     -> Subst, go diagonally
 *              else if D[j][i] == D[j-1][i] + 1 (vertical delta +1, `pv` bit) -> Ins, go up
 *              else if D[j][i-1] + 1 == D[j][i]                               -> Del, go left
 *              else                                                            -> Match, go diagonally
 * ops: 0=Match 1=Subst 2=Del 3=Ins, forward order.
 * ===================================================================================== */
/* Traceback tie-breaking among equally good predecessors.  bio's rule is not in /root/reference; the
 * default (rule 7: a substitution first, then Ins, then Del, a match last -- the branch order of
 * bio's myers traceback as recollected) is the ONLY order family that satisfies every `expected:`
 * block the reference's suite asserts on the testcases that run here (72 of 72 with the restated
 * preprocess/call glue, tools/traceback_rules.py): "diagonal first" (rules 1, 2; the round-1
 * calibration on the tumor/normal testcases alone) fails test_pcr_homopolymer_error3, "Ins, Del,
 * diagonal" (rule 0) fails test08 / test10 / test15, the others fail two or three.  The rule decides
 * DivIndelBias classes (IndelOperations::{Major,Other}) through best_indel_operations, which is why
 * the testcases see it.
 * 0 = Ins, Del, diag; 1 = diag, Ins, Del; 2 = diag, Del, Ins; 3 = match, Ins, Del, subst;
 * 4 = Ins, diag, Del; 5 = Del, diag, Ins; 6 = match, Del, Ins, subst;
 * 7 = subst, Ins, Del, match; 8 = subst, Del, Ins, match */
static int g_traceback_rule = 7;
void vlro_set_traceback_rule(int rule) { g_traceback_rule = rule; }

int vlro_best_hit(const uint8_t *hap, int len_x, const uint8_t *read, int len_y, vlro_hit_t *hit,
                  uint8_t *indel_ops, int indel_ops_cap, int32_t *aln_starts, int aln_cap,
                  uint8_t *aln_ops, int32_t *aln_ops_off, int aln_ops_cap) {
    const int m = len_y, n = len_x;
    if (n <= 0) return 0;
    size_t W = (size_t)n + 1;
    uint16_t *D = (uint16_t *)malloc(sizeof(uint16_t) * W * (size_t)(m + 1));
#define DD(j, i) D[(size_t)(j) * W + (size_t)(i)]
    for (int i = 0; i <= n; ++i) DD(0, i) = 0;
    for (int j = 1; j <= m; ++j) {
        DD(j, 0) = (uint16_t)j;
        for (int i = 1; i <= n; ++i) {
            int sub = DD(j - 1, i - 1) + (read[j - 1] == up(hap[i - 1]) ? 0 : 1);
            int ins = DD(j - 1, i) + 1;
            int del = DD(j, i - 1) + 1;
            int v = sub < ins ? sub : ins;
            if (del < v) v = del;
            DD(j, i) = (uint16_t)v;
        }
    }
    /* find_all_lazy(text, max_dist = len_y): every end position qualifies; keep all minima */
    int best = 1 << 30, n_best = 0;
    for (int i = 1; i <= n; ++i) {
        int d = DD(m, i);
        if (d < best) { best = d; n_best = 1; }
        else if (d == best) n_best++;
    }
    hit->dist = best;
    hit->n_best = n_best;
    /* tracebacks, in increasing end position */
    int k = 0, first_start = 0, last_start = 0;
    int best_n_indel = 1 << 30;
    int ops_total = 0;
    uint8_t *tmp = (uint8_t *)malloc((size_t)(m + n + 2));
    uint8_t *best_ops_tmp = (uint8_t *)malloc((size_t)(m + n + 2));
    int best_ops_len = 0;
    for (int i = 1; i <= n; ++i) {
        if (DD(m, i) != best) continue;
        int j = m, ii = i, len = 0;
        while (j > 0) {
            const int can_ins = DD(j, ii) == DD(j - 1, ii) + 1;
            const int can_del = ii > 0 && DD(j, ii - 1) + 1 == DD(j, ii);
            const int can_diag = ii > 0 && DD(j - 1, ii - 1) + (read[j - 1] == up(hap[ii - 1]) ? 0 : 1) == DD(j, ii);
            int op; /* 0 diag, 2 del, 3 ins */
            const int is_match = ii > 0 && read[j - 1] == up(hap[ii - 1]);
            if (g_traceback_rule == 0) op = can_ins ? 3 : (can_del ? 2 : 0);
            else if (g_traceback_rule == 1) op = can_diag ? 0 : (can_ins ? 3 : 2);
            else if (g_traceback_rule == 2) op = can_diag ? 0 : (can_del ? 2 : 3);
            else if (g_traceback_rule == 3) op = (can_diag && is_match) ? 0 : (can_ins ? 3 : (can_del ? 2 : 0));
            else if (g_traceback_rule == 4) op = can_ins ? 3 : (can_diag ? 0 : 2);
            else if (g_traceback_rule == 5) op = can_del ? 2 : (can_diag ? 0 : 3);
            else if (g_traceback_rule == 6) op = (can_diag && is_match) ? 0 : (can_del ? 2 : (can_ins ? 3 : 0));
            else if (g_traceback_rule == 7) op = (can_diag && !is_match) ? 0 : (can_ins ? 3 : (can_del ? 2 : 0));
            else op = (can_diag && !is_match) ? 0 : (can_del ? 2 : (can_ins ? 3 : 0));
            if (op == 3) { tmp[len++] = 3; j--; }
            else if (op == 2) { tmp[len++] = 2; ii--; }
            else {
                tmp[len++] = (DD(j - 1, ii - 1) == DD(j, ii)) ? 0 : 1;
                j--; ii--;
            }
        }
        int start = ii; /* 0-based text index of first aligned char */
        if (k == 0) first_start = start;
        last_start = start;
        /* count indel ops; keep the alignment with the fewest (first on ties: min_by_key) */
        int n_indel = 0;
        for (int t = 0; t < len; ++t) if (tmp[t] >= 2) n_indel++;
        if (n_indel < best_n_indel) {
            best_n_indel = n_indel;
            best_ops_len = 0;
            for (int t = len - 1; t >= 0; --t) if (tmp[t] >= 2) best_ops_tmp[best_ops_len++] = tmp[t];
        }
        if (aln_starts && k < aln_cap) aln_starts[k] = start;
        if (aln_ops && aln_ops_off && k < aln_cap) {
            aln_ops_off[k] = ops_total;
            for (int t = len - 1; t >= 0; --t)
                if (ops_total < aln_ops_cap) aln_ops[ops_total++] = tmp[t];
            aln_ops_off[k + 1] = ops_total;
        }
        k++;
    }
    hit->start = first_start;
    int end = last_start + m + best;
    hit->end = end < n ? end : n;
    hit->n_indel_ops = best_ops_len;
    if (indel_ops)
        for (int t = 0; t < best_ops_len && t < indel_ops_cap; ++t) indel_ops[t] = best_ops_tmp[t];
    free(tmp); free(best_ops_tmp); free(D);
#undef DD
    return 1;
}

/* =====================================================================================
 * PathHMMRealigner::calculate_prob_allele (realignment/mod.rs:499-576): probability of the
 * best Myers alignment path, with the transition table of PathHMMRealigner::new (465-486).
 * ===================================================================================== */
double vlro_pathhmm_prob(const uint8_t *hap, int len_x, const uint8_t *read, const uint8_t *qual,
                         int len_y, const double gap_ln[4], int n_aln, const int32_t *aln_starts,
                         const uint8_t *aln_ops, const int32_t *aln_ops_off) {
    (void)len_x; (void)len_y;
    double gx = gap_ln[0], gy = gap_ln[1], gxe = gap_ln[2], gye = gap_ln[3];
    double no_gap = vlro_ln_one_minus_exp(vlro_ln_add_exp(gx, gy));
    double close_x = vlro_ln_one_minus_exp(gxe);
    double close_y = vlro_ln_one_minus_exp(gye);
    double ext_or_reopen_x = vlro_ln_add_exp(gxe, close_x + gx);
    double ext_or_reopen_y = vlro_ln_add_exp(gye, close_y + gy);
    double lnconf = ln_confusion();
    double best = NEG_INF;
    int have = 0;
    for (int a = 0; a < n_aln; ++a) {
        double prob = 0.0;
        int pos_ref = aln_starts[a], pos_read = 0, prev = -1;
        for (int t = aln_ops_off[a]; t < aln_ops_off[a + 1]; ++t) {
            int op = aln_ops[t];
            if (op == 0 || op == 1) {
                if (prev == 2) prob += close_y;
                else if (prev == 3) prob += close_x;
                else if (prev == 0 || prev == 1) prob += no_gap;
                double eps = vlro_ln_prob_miscall(qual[pos_read]);
                prob += (read[pos_read] == up(hap[pos_ref])) ? vlro_ln_one_minus_exp(eps) : eps + lnconf;
                pos_ref++; pos_read++;
            } else if (op == 2) { /* Del */
                if (prev == 2) prob += ext_or_reopen_y;
                else if (prev == 3) prob += close_x + gy;
                else prob += gy;
                prob += 0.0; /* prob_emit_x */
                pos_ref++;
            } else { /* Ins */
                if (prev == 3) prob += ext_or_reopen_x;
                else if (prev == 2) prob += close_y + gx;
                else prob += gx;
                prob += vlro_ln_prob_miscall(qual[pos_read]);
                pos_read++;
            }
            prev = op;
        }
        if (!have || prob > best) { best = prob; have = 1; }
    }
    return best;
}

/* =====================================================================================
 * Realigner::prob_allele (realignment/mod.rs:338-385) and the per-region epilogue of
 * Realigner::allele_support (225-319): min-dist haplotype selection, dist==0 certainty
 * short-circuit (quirk Q7), exact/fast HMM on the shrunk window (pairhmm.rs:38-44), max over
 * haplotypes, normalisation iff both non-zero, 0.5/0.5 fallback.
 * ===================================================================================== */
typedef struct {
    double prob;
    vlro_hit_t hit;
    uint8_t indel_ops[512];
} allele_res_t;

static void prob_allele(const uint8_t *haps, const int32_t *hap_off, int h0, int nh,
                        const int32_t *shrink_extra, const uint8_t *read, const uint8_t *qual,
                        int len_y, const double gap_ln[4], unsigned flags, int fast_mode,
                        allele_res_t *res) {
    vlro_hit_t *hits = (vlro_hit_t *)malloc(sizeof(vlro_hit_t) * (size_t)nh);
    int *has = (int *)calloc((size_t)nh, sizeof(int));
    int best_dist = -1;
    for (int h = 0; h < nh; ++h) {
        const uint8_t *hp = haps + hap_off[h0 + h];
        int lx = hap_off[h0 + h + 1] - hap_off[h0 + h];
        has[h] = vlro_best_hit(hp, lx, read, len_y, &hits[h], NULL, 0, NULL, 0, NULL, NULL, 0);
        if (has[h] && (best_dist < 0 || hits[h].dist < best_dist)) best_dist = hits[h].dist;
    }
    int have = 0;
    res->prob = NEG_INF;
    memset(&res->hit, 0, sizeof(res->hit));
    for (int h = 0; h < nh; ++h) {
        if (!has[h] || hits[h].dist != best_dist) continue;
        const uint8_t *hp = haps + hap_off[h0 + h];
        int lx = hap_off[h0 + h + 1] - hap_off[h0 + h];
        double p;
        int take = 0;
        if (hits[h].dist == 0) {
            /* certainty_est (pairhmm.rs:208-210): sum of ln(1-eps_j), in order */
            p = 0.0;
            for (int j = 0; j < len_y; ++j) p += vlro_ln_prob_call(qual[j]);
            res->prob = p; /* `prob = Some(certainty_est)`: unconditional overwrite */
            have = 1;
            take = 1;
        } else {
            /* shrink_to_hit (pairhmm.rs:38-44) */
            int s = 0, e = lx;
            if (shrink_extra[h0 + h] >= 0) {
                int extra = shrink_extra[h0 + h];
                int ref_len = lx - extra; /* ref_end - ref_offset */
                int e0 = hits[h].end + 2 < ref_len ? hits[h].end + 2 : ref_len;
                s = hits[h].start >= 2 ? hits[h].start - 2 : 0;
                e = e0 + extra;
            }
            if (fast_mode) {
                /* PathHMM uses alignment.start() relative to the (already shrunk) window: the
                 * reference shrinks only in the exact realigner, so fast mode sees the full window */
                int cap = hits[h].n_best + 1;
                int32_t *st = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
                int32_t *off = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap + 1));
                int ops_cap = cap * (len_y + lx + 2);
                uint8_t *ops = (uint8_t *)malloc((size_t)ops_cap);
                vlro_hit_t tmp;
                vlro_best_hit(hp, lx, read, len_y, &tmp, NULL, 0, st, cap, ops, off, ops_cap);
                p = vlro_pathhmm_prob(hp, lx, read, qual, len_y, gap_ln, tmp.n_best, st, ops, off);
                free(st); free(off); free(ops);
            } else {
                p = vlro_pairhmm_prob_related(hp + s, e - s, read, qual, len_y, gap_ln,
                                              hits[h].dist + 2, flags);
            }
            if (!have || p > res->prob) { res->prob = p; have = 1; take = 1; }
        }
        if (take) {
            res->hit = hits[h];
            vlro_hit_t tmp;
            vlro_best_hit(hp, lx, read, len_y, &tmp, res->indel_ops, (int)sizeof(res->indel_ops),
                          NULL, 0, NULL, NULL, 0);
        }
    }
    free(hits); free(has);
}

int vlro_allele_support_region(const uint8_t *haps, const int32_t *hap_off, const int32_t n_hap[2],
                               const int32_t *shrink_extra, const uint8_t *read,
                               const uint8_t *qual, int len_y, const double gap_ln[4],
                               unsigned flags, int fast_mode, vlro_support_t *out,
                               uint8_t *indel_ops, int indel_ops_cap) {
    allele_res_t *r = (allele_res_t *)malloc(sizeof(allele_res_t) * 2);
    prob_allele(haps, hap_off, 0, n_hap[0], shrink_extra, read, qual, len_y, gap_ln, flags,
                fast_mode, &r[0]);
    prob_allele(haps, hap_off, n_hap[0], n_hap[1], shrink_extra, read, qual, len_y, gap_ln, flags,
                fast_mode, &r[1]);
    double pr = r[0].prob, pa = r[1].prob;
    out->raw_ref = pr;
    out->raw_alt = pa;
    if (pr != NEG_INF && pa != NEG_INF) {
        double tot = vlro_ln_add_exp(pa, pr);
        pr -= tot;
        pa -= tot;
    }
    if (pr == NEG_INF && pa == NEG_INF) pr = pa = vlro_log(0.5);
    out->prob_ref = pr;
    out->prob_alt = pa;
    out->informative = pr != pa;
    out->ref_dist = r[0].hit.dist;
    out->alt_dist = r[1].hit.dist;
    out->n_indel_ops = r[1].hit.n_indel_ops;
    if (indel_ops)
        for (int t = 0; t < r[1].hit.n_indel_ops && t < indel_ops_cap; ++t)
            indel_ops[t] = r[1].indel_ops[t];
    free(r);
    return 0;
}

/* =====================================================================================
 * Bias models: src/variants/model/bias/{mod,strand_bias,read_orientation_bias,
 * read_position_bias,softclip_bias,divindel_bias}.rs
 * ===================================================================================== */
static const double LN_05 = -0.6931471805599453; /* LogProb::from(Prob(0.5)), utils/mod.rs:36 */

static double bias_strand(const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    unsigned s = VLRO_CAT_STRAND(p->cat[i]); /* 0 F 1 R 2 Both 3 None */
    switch (b->strand) {                     /* strand_bias.rs:21-33 */
    case 1: if (s == 0) return 0.0; if (s == 1 || s == 2) return NEG_INF; return NAN;
    case 2: if (s == 1) return 0.0; if (s == 0 || s == 2) return NEG_INF; return NAN;
    default:
        if (s == 2) return p->prob_double_overlap[i];
        return LN_05 + p->prob_single_overlap[i];
    }
}
static double bias_orient(const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    unsigned o = VLRO_CAT_ORIENT(p->cat[i]); /* 0 F1R2 1 F2R1 */
    if (b->orientation == 1) { if (o == 0) return 0.0; if (o == 1) return NEG_INF; }
    if (b->orientation == 2) { if (o == 1) return 0.0; if (o == 0) return NEG_INF; }
    return LN_05; /* read_orientation_bias.rs:22-36 */
}
static double bias_readpos(const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    if (b->read_position == 0) return p->prob_hit_base[i]; /* read_position_bias.rs:19-29 */
    return VLRO_CAT_READPOS(p->cat[i]) == 0 ? 0.0 : NEG_INF;
}
static double bias_softclip(const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    if (b->softclip == 0) return 0.0; /* softclip_bias.rs:19-29 */
    return VLRO_CAT_SOFTCLIP(p->cat[i]) ? 0.0 : NEG_INF;
}
static double bias_divindel(const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    int other = VLRO_CAT_INDEL(p->cat[i]) == 1; /* divindel_bias.rs:37-59 */
    if (b->divindel == 0) return other ? NEG_INF : 0.0;
    if (b->other_rate == 0.0) return NEG_INF;
    return other ? vlro_log(b->other_rate) : vlro_log(1.0 - b->other_rate);
}

/* Biases::prob (bias/mod.rs:225-232): left-to-right sum */
double vlro_biases_prob(const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    return bias_strand(b, p, i) + bias_orient(b, p, i) + bias_readpos(b, p, i) +
           bias_softclip(b, p, i) + bias_divindel(b, p, i);
}
/* Biases::prob_any (bias/mod.rs:234-243) */
double vlro_biases_prob_any(const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    (void)b;
    return LN_05 + LN_05 + p->prob_hit_base[i] + 0.0 + 0.0;
}
int vlro_biases_is_artifact(const vlro_biases_t *b) {
    return b->strand != 0 || b->orientation != 0 || b->read_position != 0 || b->softclip != 0 ||
           b->divindel != 0;
}
/* Bias::is_strong_obs (bias/mod.rs:79-83): prob_mapping >= ln .95 and Kass-Raftery >= Strong,
 * i.e. exp(prob_alt - prob_ref) > 20 (Appendix A.4) */
int vlro_is_strong_obs(const vlro_pileup_t *p, int i) {
    static const double LN_095 = -0.05129329438755058;
    return p->prob_mapping[i] >= LN_095 && vlro_exp(p->prob_alt[i] - p->prob_ref[i]) > 20.0;
}

typedef double (*comp_fn)(const vlro_biases_t *, const vlro_pileup_t *, int);

static int comp_possible(comp_fn f, const vlro_biases_t *b, const vlro_pileup_t *ps, int ns) {
    for (int s = 0; s < ns; ++s)
        for (int i = 0; i < ps[s].n_obs; ++i)
            if (f(b, &ps[s], i) != NEG_INF) return 1;
    return 0;
}
/* Bias::is_likely (bias/mod.rs:46-67) for one component; `is_default` short-circuits */
static int comp_likely(comp_fn f, int is_default, int divindel, const vlro_biases_t *b,
                       const vlro_pileup_t *ps, int ns) {
    if (is_default) return 1;
    for (int s = 0; s < ns; ++s) {
        int strong_all = 0, strong_ev = 0;
        for (int i = 0; i < ps[s].n_obs; ++i) {
            if (!vlro_is_strong_obs(&ps[s], i)) continue;
            strong_all++;
            int ev = divindel ? (VLRO_CAT_INDEL(ps[s].cat[i]) == 1) : (f(b, &ps[s], i) != NEG_INF);
            if (ev) strong_ev++;
        }
        if (strong_all >= 10) {
            double ratio = (double)strong_ev / (double)strong_all;
            double min_ratio = divindel ? 0.66666 * b->other_rate : 0.66666;
            if (ratio >= min_ratio) return 1;
        } else {
            return 1;
        }
    }
    return 0;
}

int vlro_biases_gate(const vlro_biases_t *b, const vlro_pileup_t *ps, int ns) {
    /* is_possible (bias/mod.rs:192-201; DivIndel override divindel_bias.rs:77-84) */
    if (!comp_possible(bias_strand, b, ps, ns)) return 0;
    if (!comp_possible(bias_orient, b, ps, ns)) return 0;
    if (!comp_possible(bias_readpos, b, ps, ns)) return 0;
    if (!comp_possible(bias_softclip, b, ps, ns)) return 0;
    if (b->divindel) {
        int any = 0;
        for (int s = 0; s < ns && !any; ++s)
            for (int i = 0; i < ps[s].n_obs; ++i)
                if (VLRO_CAT_INDEL(ps[s].cat[i]) == 1) { any = 1; break; }
        if (!any) return 0;
    } else if (!comp_possible(bias_divindel, b, ps, ns)) return 0;
    /* is_informative (bias/mod.rs:203-212) */
    if (b->orientation) { /* read_orientation_bias.rs:42-68 */
        long n_unc = 0, n = 0;
        for (int s = 0; s < ns; ++s) {
            n += ps[s].n_obs;
            for (int i = 0; i < ps[s].n_obs; ++i)
                if (VLRO_CAT_ORIENT(ps[s].cat[i]) > 1) n_unc++;
        }
        if (!((double)n_unc < (double)n / 2.0)) return 0;
    }
    if (b->softclip) { /* softclip_bias.rs:35-43 */
        int any = 0;
        for (int s = 0; s < ns && !any; ++s)
            for (int i = 0; i < ps[s].n_obs; ++i)
                if (VLRO_CAT_SOFTCLIP(ps[s].cat[i])) { any = 1; break; }
        if (!any) return 0;
    }
    if (b->divindel) { /* divindel_bias.rs:65-75 */
        int any = 0;
        for (int s = 0; s < ns && !any; ++s)
            for (int i = 0; i < ps[s].n_obs; ++i)
                if (VLRO_CAT_INDEL(ps[s].cat[i]) == 1) { any = 1; break; }
        if (!any) return 0;
    }
    /* is_likely (bias/mod.rs:214-223) */
    if (!comp_likely(bias_strand, b->strand == 0, 0, b, ps, ns)) return 0;
    if (!comp_likely(bias_orient, b->orientation == 0, 0, b, ps, ns)) return 0;
    if (!comp_likely(bias_readpos, b->read_position == 0, 0, b, ps, ns)) return 0;
    if (!comp_likely(bias_softclip, b->softclip == 0, 0, b, ps, ns)) return 0;
    if (!comp_likely(bias_divindel, b->divindel == 0, 1, b, ps, ns)) return 0;
    return 1;
}

void vlro_biases_learn(vlro_biases_t *b, const vlro_pileup_t *ps, int ns) {
    if (!b->divindel) return;
    long strong_all = 0, strong_other = 0;
    for (int s = 0; s < ns; ++s)
        for (int i = 0; i < ps[s].n_obs; ++i)
            if (vlro_is_strong_obs(&ps[s], i)) {
                strong_all++;
                if (VLRO_CAT_INDEL(ps[s].cat[i]) == 1) strong_other++;
            }
    double r = strong_all > 0 ? (double)strong_other / (double)strong_all : 0.0;
    b->other_rate = r > b->min_other_rate ? r : b->min_other_rate;
}

/* =====================================================================================
 * Likelihood: src/variants/model/likelihood.rs
 * ===================================================================================== */
#define NUMERICAL_EPSILON 1e-3 /* utils/mod.rs:33 */

/* likelihood_mapping, likelihood.rs:191-213 with prob_sample_alt 25-38 */
double vlro_likelihood_mapping(double ln_af, const vlro_biases_t *b, const vlro_pileup_t *p, int i) {
    double psa = (ln_af != 0.0)
                     ? vlro_cap_numerical_overshoot(ln_af + p->prob_sample_alt[i], NUMERICAL_EPSILON)
                     : ln_af;
    double psr = vlro_ln_one_minus_exp(psa);
    double pb = vlro_biases_prob(b, p, i);
    double pab = vlro_biases_prob_any(b, p, i);
    double v[2] = {psa + pb + p->prob_alt[i], psr + p->prob_ref[i] + pab};
    return vlro_ln_sum_exp(v, 2);
}

/* SampleLikelihoodModel::likelihood_observation (165-187) folded over the pileup (220-244) */
double vlro_likelihood_single(const vlro_pileup_t *p, double af, const vlro_biases_t *b) {
    double ln_af = vlro_log(af);
    double lh = 0.0;
    for (int i = 0; i < p->n_obs; ++i) {
        double prob = vlro_likelihood_mapping(ln_af, b, p, i);
        double total = vlro_ln_add_exp(p->prob_mapping[i] + prob,
                                       p->prob_mismapping[i] + p->prob_missed_allele[i] +
                                           vlro_biases_prob_any(b, p, i));
        lh = lh + total;
    }
    return lh;
}

/* ContaminatedSampleLikelihoodModel (61-152) */
double vlro_likelihood_contaminated(const vlro_pileup_t *p, double purity, double af_primary,
                                    const vlro_biases_t *b_primary, double af_secondary,
                                    const vlro_biases_t *b_secondary) {
    double ln_purity = vlro_log(purity);
    double ln_impurity = vlro_ln_one_minus_exp(ln_purity);
    double ln_af_p = vlro_log(af_primary), ln_af_s = vlro_log(af_secondary);
    double lh = 0.0;
    for (int i = 0; i < p->n_obs; ++i) {
        double prob_primary = ln_purity + vlro_likelihood_mapping(ln_af_p, b_primary, p, i);
        double prob_secondary = ln_impurity + vlro_likelihood_mapping(ln_af_s, b_secondary, p, i);
        double total = vlro_ln_add_exp(
            p->prob_mapping[i] + vlro_ln_add_exp(prob_secondary, prob_primary),
            p->prob_mismapping[i] + p->prob_missed_allele[i] + vlro_biases_prob_any(b_primary, p, i));
        lh = lh + total;
    }
    return lh;
}

/* =====================================================================================
 * Posterior: GenericPosterior (generic.rs:109-282), GenericLikelihood (317-365),
 * bio::stats::bayesian::Model (Appendix A.4), MAP selection (calling.rs:635-677)
 * ===================================================================================== */
int vlro_grid_points(int n_obs, int resolution) {
    int n = n_obs + 1 > 5 ? n_obs + 1 : 5;
    if (resolution < n) n = resolution;
    if (n % 2 == 0) n += 1;
    return n;
}

/* VAFRange::observable_max (formula.rs:1056-1071) */
double vlro_observable_max(double start, double end, int left_excl, int right_excl, int n_obs) {
    (void)start; (void)left_excl;
    if (n_obs < 10) return end;
    double obs_count = (double)n_obs * end;
    if (right_excl && fmod(obs_count, 1.0) == 0.0) obs_count -= 1.0;
    return floor(obs_count) / (double)n_obs;
}
/* VAFRange::observable_min (formula.rs:1029-1054) */
double vlro_observable_min(double start, double end, int left_excl, int right_excl, int n_obs) {
    if (n_obs < 10) return start;
    double obs_count = (double)n_obs * start;
    if (left_excl && fmod(obs_count, 1.0) == 0.0) {
        double adjusted_end = vlro_observable_max(start, end, left_excl, right_excl, n_obs);
        double offs[2] = {1.0, 0.0};
        for (int k = 0; k < 2; ++k) {
            double adj = ceil(obs_count + offs[k]) / (double)n_obs;
            if (adj <= 1.0 && adj <= adjusted_end) return adj;
        }
    }
    return ceil(obs_count) / (double)n_obs;
}

/* ---- likelihood cache (likelihood.rs:128,221): open-addressing table per sample ---- */
typedef struct {
    uint64_t af, af2;
    int32_t bias;
    int32_t used;
    double val;
} cache_ent_t;
typedef struct {
    cache_ent_t *e;
    size_t cap, n;
} cache_t;

static uint64_t dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static void cache_init(cache_t *c) { c->cap = 1024; c->n = 0; c->e = (cache_ent_t *)calloc(c->cap, sizeof(cache_ent_t)); }
static void cache_free(cache_t *c) { free(c->e); }
static cache_ent_t *cache_slot(cache_t *c, uint64_t af, uint64_t af2, int32_t bias) {
    uint64_t h = af * 0x9E3779B97F4A7C15ull ^ (af2 + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)bias * 0x165667B19E3779F9ull;
    h ^= h >> 29;
    size_t idx = (size_t)h & (c->cap - 1);
    for (;;) {
        cache_ent_t *e = &c->e[idx];
        if (!e->used || (e->af == af && e->af2 == af2 && e->bias == bias)) return e;
        idx = (idx + 1) & (c->cap - 1);
    }
}
static void cache_grow(cache_t *c) {
    cache_t n;
    n.cap = c->cap * 2; n.n = c->n;
    n.e = (cache_ent_t *)calloc(n.cap, sizeof(cache_ent_t));
    for (size_t i = 0; i < c->cap; ++i)
        if (c->e[i].used) *cache_slot(&n, c->e[i].af, c->e[i].af2, c->e[i].bias) = c->e[i];
    free(c->e);
    *c = n;
}

/* ---- MAP bookkeeping: joint_probs of every visited base event ---- */
typedef struct {
    double p;
    int32_t bias;
    double af[8];
} joint_t;

typedef struct {
    const vlro_model_t *m;
    const vlro_pileup_t *pileups;
    const vlro_biases_t *biases; /* learned copy */
    cache_t *caches;
    int grid[8];
    double cur_af[8];
    int cur_bias;
    joint_t best_any, best_clean; /* max joint over all / over non-artifact base events */
    int have_any, have_clean;
    int64_t n_evals;
    vlro_prior_fn prior_cb;
    void *prior_user;
    int nan_flag;
} post_ctx_t;

/* GenericLikelihood::compute (generic.rs:321-365) + Model::joint_prob */
static double joint_prob(post_ctx_t *c) {
    const vlro_model_t *m = c->m;
    double prior = c->prior_cb ? c->prior_cb(c->cur_af, m->n_samples, c->prior_user) : 0.0;
    double p = 0.0;
    const vlro_biases_t *b = &c->biases[c->cur_bias];
    for (int s = 0; s < m->n_samples; ++s) {
        double af = c->cur_af[s];
        double af2 = m->contaminated[s] ? c->cur_af[m->by[s]] : 0.0;
        uint64_t k1 = dbits(af), k2 = m->contaminated[s] ? dbits(af2) : 0xFFFFFFFFFFFFFFFFull;
        cache_ent_t *e = cache_slot(&c->caches[s], k1, k2, c->cur_bias);
        if (!e->used) {
            double v = m->contaminated[s]
                           ? vlro_likelihood_contaminated(&c->pileups[s], m->purity[s], af, b, af2, b)
                           : vlro_likelihood_single(&c->pileups[s], af, b);
            e->used = 1; e->af = k1; e->af2 = k2; e->bias = c->cur_bias; e->val = v;
            c->n_evals += 1;
            if (++c->caches[s].n * 2 > c->caches[s].cap) {
                cache_grow(&c->caches[s]);
                e = cache_slot(&c->caches[s], k1, k2, c->cur_bias);
            }
        }
        p += e->val;
    }
    double j = prior + p;
    if (isnan(j)) c->nan_flag = 1;
    /* joint_probs.insert(base_event, p): keep the running maxima.  The reference sorts a HashMap
     * (order among exact ties unspecified, quirk Q14); the oracle and the GPU path share one
     * explicit rule: larger joint, then earlier (event, bias) entry, then the lexicographically
     * smaller allele-frequency vector. */
    int art = vlro_biases_is_artifact(b);
    joint_t *slots[2] = {&c->best_any, art ? NULL : &c->best_clean};
    int *have[2] = {&c->have_any, &c->have_clean};
    for (int w = 0; w < 2; ++w) {
        joint_t *bst = slots[w];
        if (!bst || isnan(j)) continue;
        int better = 0;
        if (!*have[w]) better = 1;
        else if (j != bst->p) better = j > bst->p;
        else if (c->cur_bias != bst->bias) better = c->cur_bias < bst->bias;
        else {
            for (int s = 0; s < m->n_samples; ++s)
                if (c->cur_af[s] != bst->af[s]) { better = c->cur_af[s] < bst->af[s]; break; }
        }
        if (better) {
            *have[w] = 1; bst->p = j; bst->bias = c->cur_bias;
            memcpy(bst->af, c->cur_af, sizeof(double) * (size_t)m->n_samples);
        }
    }
    return j;
}

static double density(post_ctx_t *c, int node);

/* the `subdensity` closure, generic.rs:133-165 */
static double subdensity(post_ctx_t *c, const vlro_node_t *nd) {
    if (nd->n_children == 0) return joint_prob(c);
    if (nd->n_children > 1) {
        double *v = (double *)malloc(sizeof(double) * (size_t)nd->n_children);
        double saved[8];
        memcpy(saved, c->cur_af, sizeof(saved));
        for (int k = 0; k < nd->n_children; ++k) {
            memcpy(c->cur_af, saved, sizeof(saved)); /* base_events.clone() per child */
            v[k] = density(c, c->m->children[nd->child_off + k]);
        }
        memcpy(c->cur_af, saved, sizeof(saved));
        double r = vlro_ln_sum_exp(v, (size_t)nd->n_children);
        free(v);
        return r;
    }
    return density(c, c->m->children[nd->child_off]);
}

/* GenericPosterior::density, generic.rs:124-235 */
static double density(post_ctx_t *c, int node) {
    const vlro_model_t *m = c->m;
    const vlro_node_t *nd = &m->nodes[node];
    switch (nd->kind) {
    case VLRO_NODE_FALSE: return NEG_INF;
    case VLRO_NODE_SKIP: return subdensity(c, nd);
    case VLRO_NODE_SET: {
        const double *vafs = m->vafs + nd->vaf_off;
        if (nd->n_vafs == 1) {
            c->cur_af[nd->sample] = vafs[0];
            return subdensity(c, nd);
        }
        double *v = (double *)malloc(sizeof(double) * (size_t)nd->n_vafs);
        for (int k = 0; k < nd->n_vafs; ++k) {
            c->cur_af[nd->sample] = vafs[k];
            v[k] = subdensity(c, nd);
        }
        double r = vlro_ln_sum_exp(v, (size_t)nd->n_vafs);
        free(v);
        return r;
    }
    case VLRO_NODE_RANGE: {
        const double *r = m->vafs + nd->vaf_off;
        int n_obs = c->pileups[nd->sample].n_obs;
        double a = vlro_observable_min(r[0], r[1], r[2] != 0.0, r[3] != 0.0, n_obs);
        double b = vlro_observable_max(r[0], r[1], r[2] != 0.0, r[3] != 0.0, n_obs);
        int n = c->grid[nd->sample];
        /* LogProb::ln_simpsons_integrate_exp (Appendix A.1): interior ascending, then a, then b */
        double *probs = (double *)malloc(sizeof(double) * (size_t)n);
        double step = (b - a) / (double)(n - 1);
        int k = 0;
        for (int i = 1; i < n - 1; ++i) {
            double x = a + step * (double)i;
            c->cur_af[nd->sample] = x;
            double w = (double)(2 + (i % 2) * 2);
            probs[k++] = subdensity(c, nd) + vlro_log(w);
        }
        c->cur_af[nd->sample] = a;
        probs[k++] = subdensity(c, nd);
        c->cur_af[nd->sample] = b;
        probs[k++] = subdensity(c, nd);
        double res = vlro_ln_sum_exp(probs, (size_t)k) + vlro_log(b - a) - vlro_log((double)(n - 1)) - vlro_log(3.0);
        free(probs);
        return res;
    }
    default: return NAN;
    }
}

int vlro_model_compute(const vlro_model_t *m, const vlro_pileup_t *pileups, double *out_event_ln,
                       double *out_marginal, double *out_map_af, int32_t *out_map_bias,
                       int64_t *out_n_evals, vlro_prior_fn prior_cb, void *prior_user) {
    if (m->n_samples > 8) return -1;
    post_ctx_t c;
    memset(&c, 0, sizeof(c));
    c.m = m; c.pileups = pileups; c.prior_cb = prior_cb; c.prior_user = prior_user;
    /* learn_parameters on a per-site copy (calling.rs:557-565) */
    vlro_biases_t *biases = (vlro_biases_t *)malloc(sizeof(vlro_biases_t) * (size_t)(m->n_biases_total > 0 ? m->n_biases_total : 1));
    for (int k = 0; k < m->n_biases_total; ++k) {
        biases[k] = m->biases[k];
        vlro_biases_learn(&biases[k], pileups, m->n_samples);
    }
    c.biases = biases;
    c.caches = (cache_t *)malloc(sizeof(cache_t) * (size_t)m->n_samples);
    for (int s = 0; s < m->n_samples; ++s) {
        cache_init(&c.caches[s]);
        c.grid[s] = vlro_grid_points(pileups[s].n_obs, m->resolution[s]);
    }
    double art_post[64];
    int n_art = 0;
    for (int ev = 0; ev < m->n_events; ++ev) {
        const vlro_event_t *E = &m->events[ev];
        int is_art = 0;
        for (int k = 0; k < E->n_biases; ++k) is_art |= vlro_biases_is_artifact(&biases[E->bias_off + k]);
        /* generic.rs:251-255: |biases| counted before gating */
        double bias_prior = is_art ? LN_05 + vlro_log(1.0 / (double)E->n_biases) : LN_05;
        size_t cap = (size_t)E->n_biases * (size_t)E->n_roots;
        double *terms = (double *)malloc(sizeof(double) * (cap ? cap : 1));
        size_t nt = 0;
        for (int k = 0; k < E->n_biases; ++k) {
            const vlro_biases_t *b = &biases[E->bias_off + k];
            if (!vlro_biases_gate(b, pileups, m->n_samples)) continue;
            for (int r = 0; r < E->n_roots; ++r) {
                c.cur_bias = E->bias_off + k;
                for (int s = 0; s < 8; ++s) c.cur_af[s] = 0.0;
                terms[nt++] = bias_prior + density(&c, m->roots[E->root_off + r]);
            }
        }
        out_event_ln[ev] = vlro_ln_sum_exp(terms, nt);
        free(terms);
        (void)art_post; (void)n_art;
    }
    /* Model::compute: marginal = ln_sum_exp(posterior_probs) (HashMap order in the reference) */
    double marginal = vlro_ln_sum_exp(out_event_ln, (size_t)m->n_events);
    if (out_marginal) *out_marginal = marginal;
    /* is_artifact (calling.rs:571-598) */
    double art_terms[256];
    int na = 0;
    for (int ev = 0; ev < m->n_events && na < 256; ++ev) {
        const vlro_event_t *E = &m->events[ev];
        int is_art = 0;
        for (int k = 0; k < E->n_biases; ++k) is_art |= vlro_biases_is_artifact(&biases[E->bias_off + k]);
        if (is_art) art_terms[na++] = out_event_ln[ev] - marginal;
    }
    double prob_art = vlro_ln_sum_exp(art_terms, (size_t)na);
    int is_artifact = 1;
    for (int ev = 0; ev < m->n_events; ++ev) {
        const vlro_event_t *E = &m->events[ev];
        int is_art = 0;
        for (int k = 0; k < E->n_biases; ++k) is_art |= vlro_biases_is_artifact(&biases[E->bias_off + k]);
        if (!is_art && !((out_event_ln[ev] - marginal) < prob_art)) is_artifact = 0;
    }
    /* sample_infos (calling.rs:635-677): best joint; artifact MAPs skipped unless is_artifact */
    const joint_t *best = NULL;
    if (is_artifact && c.have_any) best = &c.best_any;
    else if (c.have_clean) best = &c.best_clean;
    for (int s = 0; s < m->n_samples; ++s) {
        if (!best) { out_map_af[s] = NAN; if (out_map_bias) out_map_bias[s] = -1; continue; }
        int art = vlro_biases_is_artifact(&biases[best->bias]);
        out_map_af[s] = art ? 0.0 : best->af[s];
        if (out_map_bias) out_map_bias[s] = art ? best->bias : -1;
    }
    if (out_n_evals) *out_n_evals = c.n_evals;
    for (int s = 0; s < m->n_samples; ++s) cache_free(&c.caches[s]);
    free(c.caches);
    free(biases);
    return c.nan_flag ? 1 : 0;
}


/* =====================================================================================
 * All-cores driver for the CPU legs of bench.py: the per-site Model::compute above over a batch of
 * sites given as struct-of-arrays columns (rows obs_off[s*n_samples + m] .. obs_off[s*n_samples + m + 1]),
 * sites dealt round-robin to n_threads POSIX threads (the reference's `call` processes records
 * independently, calling.rs:283-330).  Flat prior only (the bench workloads).
 * ===================================================================================== */
#include <pthread.h>
typedef struct {
    const vlro_model_t *m;
    const vlro_pileup_t *cols;   /* column base pointers (n_obs unused) */
    const int64_t *obs_off;
    int64_t n_sites;
    int t, n_threads;
    double *event_ln, *marginal, *map_af;
    int32_t *map_bias;
    int64_t *n_evals;
    int status;
} batch_job_t;

static void *batch_worker(void *arg) {
    batch_job_t *j = (batch_job_t *)arg;
    const int ns = j->m->n_samples, ne = j->m->n_events;
    vlro_pileup_t pv[8];
    for (int64_t s = j->t; s < j->n_sites; s += j->n_threads) {
        for (int k = 0; k < ns; ++k) {
            const int64_t lo = j->obs_off[s * ns + k], hi = j->obs_off[s * ns + k + 1];
            pv[k].prob_mapping = j->cols->prob_mapping + lo;
            pv[k].prob_mismapping = j->cols->prob_mismapping + lo;
            pv[k].prob_alt = j->cols->prob_alt + lo;
            pv[k].prob_ref = j->cols->prob_ref + lo;
            pv[k].prob_missed_allele = j->cols->prob_missed_allele + lo;
            pv[k].prob_sample_alt = j->cols->prob_sample_alt + lo;
            pv[k].prob_double_overlap = j->cols->prob_double_overlap + lo;
            pv[k].prob_single_overlap = j->cols->prob_single_overlap + lo;
            pv[k].prob_hit_base = j->cols->prob_hit_base + lo;
            pv[k].cat = j->cols->cat + lo;
            pv[k].n_obs = (int32_t)(hi - lo);
        }
        int64_t nev = 0;
        const int st = vlro_model_compute(j->m, pv, j->event_ln + s * ne, j->marginal + s, j->map_af + s * ns,
                                          j->map_bias ? j->map_bias + s * ns : NULL, &nev, NULL, NULL);
        if (j->n_evals) j->n_evals[s] = nev;
        if (st < 0) j->status = st;
    }
    return NULL;
}

int vlro_model_compute_batch(const vlro_model_t *m, const vlro_pileup_t *cols, const int64_t *obs_off,
                             int64_t n_sites, int n_threads, double *out_event_ln, double *out_marginal,
                             double *out_map_af, int32_t *out_map_bias, int64_t *out_n_evals) {
    if (m->n_samples > 8 || n_threads < 1) return -1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    batch_job_t jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (batch_job_t){m, cols, obs_off, n_sites, t, n_threads, out_event_ln, out_marginal,
                                out_map_af, out_map_bias, out_n_evals, 0};
        if (n_threads == 1) batch_worker(&jobs[t]);
        else if (pthread_create(&th[t], NULL, batch_worker, &jobs[t]) != 0) return -2;
    }
    int status = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        if (jobs[t].status) status = jobs[t].status;
    }
    return status;
}
