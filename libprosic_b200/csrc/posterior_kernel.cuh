// posterior_kernel.cuh -- K3/K4: AF-grid pileup likelihood and fused event integration for sm_100a.
//
// Replaces, per candidate site,
//   * SampleLikelihoodModel / ContaminatedSampleLikelihoodModel::compute
//       (reference src/variants/model/likelihood.rs:118-152, 215-245; likelihood_mapping 191-213)
//   * Biases::{prob, prob_any, is_possible, is_informative, is_likely, learn_parameters}
//       (src/variants/model/bias/*.rs)
//   * GenericPosterior::{grid_points, density, compute} (src/variants/model/modes/generic.rs:109-282)
//   * bio Model::compute (marginal) and the MAP selection of calling.rs:635-677.
//
// Arithmetic.  For observation i, bias configuration B and effective alt-sampling probability z
//   T_i = pm_i * ( z * beta_iB * A_i + (1 - z) * R_i * betabar_i ) + pmm_i * mu_i * betabar_i
// with A,R,mu = exp(prob_alt, prob_ref, prob_missed_allele), z = s_i(af) for a plain sample and
// z = rho*s_i(af_primary) + (1-rho)*s_i(af_secondary) for a contaminated one (likelihood_mapping
// is affine in s), s_i(af) = af * exp(prob_sample_alt_i) (exactly 1 for af == 1, quirk Q4).
// The pileup likelihood is sum_i ln T_i = sum_i m_i + ln prod_i T'_i with per-observation scaling
// m_i = max(prob_alt, prob_ref, prob_missed): the kernel multiplies mantissas and adds exponents
// (no per-evaluation log), and takes ONE log per (allele-frequency key, bias configuration).
// Evaluations whose T' leaves the normal fp64 range are recomputed in log space (same formulas as
// the reference), still on the GPU.
//
// Mapping.  A team of TEAM warps owns one site at a time.  Observations sit one per lane
// (strided) in registers; for every likelihood key the lanes evaluate their observations and the
// warp reduces mantissa product / exponent sum by shuffles (segmented reduction over the
// observation columns, struct-of-arrays in HBM).  The likelihood table of the site is then
// integrated: every (event, bias configuration, tree path, grid combination) is one term
// W * exp(joint) of a log-sum-exp, where W carries the Simpson weights (generic.rs:196-208) -- the
// nested ln_sum_exp / ln_simpsons_integrate_exp of the reference are distributive sums.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace vlr {

constexpr int kMaxSamples = 8;
constexpr int kMaxSpectra = 48;
constexpr int kMaxValues = 4096;  // distinct allele-frequency values per site (all samples; dynamic shared memory)
constexpr int kMaxCfg = 32;       // distinct bias configurations
constexpr int kMaxEvents = 64;
constexpr int kTeamThreadsMax = 128;

struct DevSpectrum {
    int32_t sample, kind, vaf_off, n_vafs;
};
struct DevPath {
    int32_t event;
    int32_t spec[kMaxSamples];
};
struct DevEventBias {  // one (event, bias configuration) entry
    int32_t event, cfg;
};
// one (event, bias configuration, tree path) segment of the integration, in the order the
// reference visits them (events, then the event's bias list, then tree paths)
struct DevSeg {
    int32_t event, ebi, cfg, art, path, _pad;
    int64_t leaf_off;            // first prior-table entry of (event, path); -1 without a prior
    int16_t spec[kMaxSamples];   // spectrum bound by the path, per sample
    int16_t nv[kMaxSamples];     // its number of values; -1 = Range (the site's grid size)
};
struct DevEvent {
    int32_t is_artifact;
    int32_t path_off, n_paths;
    int32_t eb_off, n_eb;
    int32_t seg_off, n_seg;  // the event's (bias, path) segments in DevModel::segs
    double bias_prior;  // ln .5 (+ ln 1/|biases| for artifact events), generic.rs:251-255
};

struct DevModel {
    int32_t n_samples, n_events, n_spectra, n_paths, n_cfg, n_eb, n_segs;
    int32_t resolution[kMaxSamples], contaminated[kMaxSamples], by[kMaxSamples];
    double purity[kMaxSamples];
    const DevSpectrum *spectra;
    const DevPath *paths;
    const DevEvent *events;
    const DevEventBias *eb;
    const vlr_biases *cfg;
    const double *vafs;
    const DevSeg *segs;
    const double *prior;  // ln prior per leaf (event, path, combination), nullptr = flat
    // prior as a program (vlr_prior_prog in device memory, its arrays patched to device copies),
    // used when prior == nullptr
    const vlr_prior_prog *prog;
};

struct PostArgs {
    DevModel m;
    vlr_obs_columns obs;
    const int64_t *obs_off;
    int64_t n_sites;
    int32_t *cursor;
    double *table;          // per-team scratch: likelihood tables
    int64_t table_stride;   // doubles per team
    double *out_event_ln, *out_marginal, *out_map_af;
    int32_t *out_map_bias;
    int32_t *status;        // [0] = number of sites that overflowed a kernel limit
    int32_t obs_cap;        // observations per (site, sample) the shared staging can hold
    int64_t gs_cap;         // doubles of per-team shared memory for the blocked contaminated path
    int64_t tab_cap;        // doubles of per-team shared memory for the likelihood table (0: global)
    int32_t seg_cap;        // ints of per-team shared memory for the segment sizes (n_segs, padded)
    int32_t val_cap;        // allele-frequency values (+ ln weights) per site the shared staging can hold
};
__device__ __forceinline__ int m_ncfg_cap(const PostArgs &g) { return g.m.n_cfg; }

// Out-of-line transcendentals: one copy each instead of one per call site (the fully inlined kernel
// was 343 KB of SASS and stalled on instruction fetch; ncu profile r1 `no_instruction`).
static __device__ __noinline__ double exp_ni(double x) { return exp(x); }
static __device__ __noinline__ double log_ni(double x) { return log(x); }
static __device__ __noinline__ double log1p_ni(double x) { return log1p(x); }
static __device__ __noinline__ double ln_one_minus_exp_d(double p) {
    return p < -0.693 ? log1p(-exp(p)) : log(-expm1(p));
}
static __device__ __noinline__ double ln_add_exp2(double a, double b) {
    double p0 = fmax(a, b), p1 = fmin(a, b);
    if (p0 == -INFINITY) return -INFINITY;
    if (p1 == -INFINITY) return p0;
    return p0 + log1p_ni(exp_ni(p1 - p0));
}

// ---- bias terms (bias/*.rs).  Returns the linear probability; NaN where the reference panics.
__device__ __forceinline__ double bias_lin(const vlr_biases &b, double other_rate, unsigned cat,
                                           double strand_none, double hb) {
    const unsigned st = VLR_CAT_STRAND(cat), orr = VLR_CAT_ORIENT(cat);
    double v;
    if (b.strand == 1) v = st == 0 ? 1.0 : (st == 3 ? NAN : 0.0);          // strand_bias.rs:21-33
    else if (b.strand == 2) v = st == 1 ? 1.0 : (st == 3 ? NAN : 0.0);
    else v = strand_none;  // e^double_overlap | .5 * e^single_overlap
    double o = 0.5;                                                          // read_orientation_bias.rs:22-36
    if (b.orientation == 1) o = orr == 0 ? 1.0 : (orr == 1 ? 0.0 : 0.5);
    else if (b.orientation == 2) o = orr == 1 ? 1.0 : (orr == 0 ? 0.0 : 0.5);
    v *= o;
    v *= b.read_position == 0 ? hb : (VLR_CAT_READPOS(cat) == 0 ? 1.0 : 0.0);  // read_position_bias.rs:19-29
    if (b.softclip) v *= VLR_CAT_SOFTCLIP(cat) ? 1.0 : 0.0;                    // softclip_bias.rs:19-29
    const bool other = VLR_CAT_INDEL(cat) == 1;                                // divindel_bias.rs:37-59
    if (b.divindel == 0) v *= other ? 0.0 : 1.0;
    else v *= other_rate == 0.0 ? 0.0 : (other ? other_rate : 1.0 - other_rate);
    return v;
}
static __device__ __noinline__ double bias_ln(const vlr_biases &b, double other_rate, unsigned cat,
                                          double ln_dov, double ln_sov, double ln_hb) {
    const double LN05 = -0.6931471805599453;
    const unsigned st = VLR_CAT_STRAND(cat), orr = VLR_CAT_ORIENT(cat);
    double v;
    if (b.strand == 1) v = st == 0 ? 0.0 : (st == 3 ? NAN : -INFINITY);
    else if (b.strand == 2) v = st == 1 ? 0.0 : (st == 3 ? NAN : -INFINITY);
    else v = st == 2 ? ln_dov : LN05 + ln_sov;
    double o = LN05;
    if (b.orientation == 1) o = orr == 0 ? 0.0 : (orr == 1 ? -INFINITY : LN05);
    else if (b.orientation == 2) o = orr == 1 ? 0.0 : (orr == 0 ? -INFINITY : LN05);
    v += o;
    v += b.read_position == 0 ? ln_hb : (VLR_CAT_READPOS(cat) == 0 ? 0.0 : -INFINITY);
    v += b.softclip ? (VLR_CAT_SOFTCLIP(cat) ? 0.0 : -INFINITY) : 0.0;
    const bool other = VLR_CAT_INDEL(cat) == 1;
    if (b.divindel == 0) v += other ? -INFINITY : 0.0;
    else v += other_rate == 0.0 ? -INFINITY : (other ? log_ni(other_rate) : log_ni(1.0 - other_rate));
    return v;
}

// likelihood_mapping in log space (likelihood.rs:191-213), used by the out-of-range path
static __device__ __noinline__ double lh_mapping_ln(double af, double psa_obs, double pb, double pab, double pa,
                                double pr) {
    const double ln_af = log_ni(af);
    double psa;
    if (ln_af != 0.0) {
        psa = ln_af + psa_obs;
        if (psa > 0.0) psa = psa <= 1e-3 ? 0.0 : NAN;  // cap_numerical_overshoot
    } else {
        psa = ln_af;
    }
    const double psr = ln_one_minus_exp_d(psa);
    const double t0 = psa + pb + pa, t1 = psr + pr + pab;
    // ln_sum_exp of two terms
    const double mx = fmax(t0, t1), mn = fmin(t0, t1);
    if (isnan(t0) || isnan(t1)) return NAN;
    if (mx == -INFINITY) return -INFINITY;
    if (mn == -INFINITY) return mx;
    return mx + log1p_ni(exp_ni(mn - mx));
}

// exact ln T of one observation in log space (likelihood.rs:75-116, 165-187)
static __device__ __noinline__ double ln_T_exact(const vlr_obs_columns &o, int64_t row, const vlr_biases &b,
                             double other_rate, double af_p, double af_s, bool contaminated,
                             double purity) {
    const unsigned cat = o.cat[row];
    const double ln_hb = o.prob_hit_base[row];
    const double pb = bias_ln(b, other_rate, cat, o.prob_double_overlap[row],
                              o.prob_single_overlap[row], ln_hb);
    const double pab = -0.6931471805599453 + -0.6931471805599453 + ln_hb + 0.0 + 0.0;
    const double pa = o.prob_alt[row], pr = o.prob_ref[row], psa = o.prob_sample_alt[row];
    double mapped;
    if (contaminated) {
        const double lp = log_ni(purity);
        const double li = ln_one_minus_exp_d(lp);
        const double a = lp + lh_mapping_ln(af_p, psa, pb, pab, pa, pr);
        const double c = li + lh_mapping_ln(af_s, psa, pb, pab, pa, pr);
        mapped = (isnan(a) || isnan(c)) ? NAN : ln_add_exp2(c, a);
    } else {
        mapped = lh_mapping_ln(af_p, psa, pb, pab, pa, pr);
    }
    const double t0 = o.prob_mapping[row] + mapped;
    const double t1 = o.prob_mismapping[row] + o.prob_missed_allele[row] + pab;
    if (isnan(t0) || isnan(t1)) return NAN;
    return ln_add_exp2(t0, t1);
}

// Prior::compute as a program (vlr_prior_prog, include/vlr_gpu.h): Prior::calc_prob (prior.rs:288-431)
// at one allele-frequency vector.  The nested ln_sum_exp of the recursion is a flat sum here.
static __device__ __noinline__ double prior_prog_eval(const vlr_prior_prog *Pp, int NS, const double *af) {
    const vlr_prior_prog &P = *Pp;
    int stride[kMaxSamples], radix[kMaxSamples], bs[kMaxSamples];
    double germ[kMaxSamples];
    int nb = 0, total = 1;
    for (int s = NS - 1, st = 1; s >= 0; --s) {
        radix[s] = (P.kind[s] == VLR_PRIOR_UNIFORM || P.ploidy[s] <= 0) ? 1 : P.ploidy[s] + 1;
        stride[s] = st;
        st *= radix[s];
    }
    int fixed = 0;
    for (int s = 0; s < NS; ++s) {
        germ[s] = 0.0;
        if (P.ploidy[s] == 0 && af[s] != 0.0) return -INFINITY;  // chromosome absent from the sample
        if (P.kind[s] == VLR_PRIOR_UNIFORM) {
            bool in = false;
            for (int u = P.uni_off[s]; u < P.uni_off[s + 1] && !in; ++u) {
                const double *e = P.uni_spec + 5 * u;
                if (e[0] == 0.0) in = af[s] == e[1];
                else in = (e[3] != 0.0 ? af[s] > e[1] : af[s] >= e[1]) && (e[4] != 0.0 ? af[s] < e[2] : af[s] <= e[2]);
            }
            if (!in) return -INFINITY;
        } else if (P.kind[s] == VLR_PRIOR_GERMLINE) {
            const double x = (double)P.ploidy[s] * af[s], r = rint(x);
            if (fabs(x - r) > fmax(1e-9 * fmax(fabs(x), fabs(r)), 1e-9)) return -INFINITY;  // is_valid_germline_vaf
            fixed += (int)r * stride[s];
            germ[s] = af[s];
        } else if (radix[s] > 1) {
            bs[nb++] = s;
            total *= radix[s];
        }
    }
    double mx = -INFINITY, sum = 0.0;
    for (int combo = 0; combo < total; ++combo) {
        int idx = fixed, rem = combo;
        for (int b = nb - 1; b >= 0; --b) {
            const int s = bs[b], k = rem % radix[s];
            rem /= radix[s];
            idx += k * stride[s];
            germ[s] = (double)k / (double)P.ploidy[s];
        }
        double t = P.germ_ln[idx];
        if (t == -INFINITY) continue;
        for (int s = 0; s < NS && t != -INFINITY; ++s) {
            const int sk = P.som_kind[s];
            if (sk == VLR_SOM_NONE) continue;
            const double eff = af[s] - germ[s];
            double v = eff;
            if (sk != VLR_SOM_PLAIN) {
                const double effp = af[P.parent[s]] - germ[P.parent[s]];
                if (sk == VLR_SOM_CLONAL_SAME) {  // approx::relative_eq! with the crate's defaults
                    const double d = fabs(eff - effp);
                    if (!(eff == effp || d <= 2.220446049250313e-16 ||
                          d <= fmax(fabs(eff), fabs(effp)) * 2.220446049250313e-16))
                        t = -INFINITY;
                    continue;
                }
                v = af[s] - germ[s] - effp;
            }
            v = fabs(v);
            t += v <= 0.0001 ? P.som_zero[s] : P.ln_rate[s] - (2.0 * log_ni(v) + P.ln_genome_size);
        }
        if (t == -INFINITY) continue;
        if (t > mx) { sum = sum * exp_ni(mx - t) + 1.0; mx = t; }
        else sum += exp_ni(t - mx);
    }
    return mx == -INFINITY ? -INFINITY : mx + log_ni(sum);
}

// Bias::is_strong_obs (bias/mod.rs:79-83)
__device__ __forceinline__ bool strong_obs(double prob_mapping, double pa, double pr) {
    return prob_mapping >= -0.05129329438755058 && exp_ni(pa - pr) > 20.0;
}

// VAFRange::observable_{min,max} (grammar/formula.rs:1029-1071)
static __device__ __noinline__ double observable_max_d(double end, bool rex, int n_obs) {
    if (n_obs < 10) return end;
    double c = (double)n_obs * end;
    if (rex && fmod(c, 1.0) == 0.0) c -= 1.0;
    return floor(c) / (double)n_obs;
}
static __device__ __noinline__ double observable_min_d(double start, double end, bool lex, bool rex, int n_obs) {
    if (n_obs < 10) return start;
    const double c = (double)n_obs * start;
    if (lex && fmod(c, 1.0) == 0.0) {
        const double adj_end = observable_max_d(end, rex, n_obs);
        for (int k = 0; k < 2; ++k) {
            const double adj = ceil(c + (k == 0 ? 1.0 : 0.0)) / (double)n_obs;
            if (adj <= 1.0 && adj <= adj_end) return adj;
        }
    }
    return ceil(c) / (double)n_obs;
}

// evidence "variants" of the gating counters
enum { EV_SF = 0, EV_SR, EV_OF1R2, EV_OF2R1, EV_RP, EV_SC, EV_DI, EV_N };
// possibility flags of the unbiased components
enum { PS_STRAND_NONE = 0, PS_RP_NONE, PS_DI_NONE, PS_ANYOBS, PS_N };

struct TeamShared {
    int n_obs[kMaxSamples];
    int grid[kMaxSamples];
    int strong_all[kMaxSamples];
    int strong_ev[kMaxSamples][EV_N];
    int any_ev[EV_N];
    int any_ps[PS_N];
    int n_uncertain, n_total, strong_other_total, strong_all_total;
    int spec_off[kMaxSpectra + 1];       // into val[]
    int sample_voff[kMaxSamples + 1];    // value ranges are grouped by sample: val index base
    int sample_nv[kMaxSamples];
    int64_t tab_off[kMaxSamples];        // offset of sample's table (doubles), per cfg stride below
    int64_t tab_cfg_stride[kMaxSamples];
    double *val, *lnw;                   // [val_cap] each, in the team's dynamic shared memory
    double sum_m[kMaxSamples];
    double other_rate[kMaxCfg];
    int allowed[kMaxCfg];
    int alist[kMaxCfg];   // allowed configurations, compacted
    int n_allowed, slow;
    double ev_ln[kMaxEvents];
    // reduction scratch
    double red_mx[4], red_sum[4];
    double best_j[4][2];
    int best_id[4][2][3];  // eb index, path, combo
    int overflow;
};


// decode the allele-frequency vector of (path, combo) -- LAST sample is the fastest digit
static __device__ __noinline__ void decode_afs(const DevModel &m, const TeamShared &S, const DevPath &path, int64_t combo,
                           double *afs) {
#pragma unroll 1
    for (int smp = m.n_samples - 1; smp >= 0; --smp) {
        const DevSpectrum spc = m.spectra[path.spec[smp]];
        const int cnt = spc.kind == VLR_NODE_RANGE ? S.grid[smp] : spc.n_vafs;
        const int d = (int)(combo % cnt);
        combo /= cnt;
        afs[smp] = S.val[S.spec_off[path.spec[smp]] + d];
    }
}
// MAP tie-break: larger joint first; then the earlier (event, bias) entry; then the
// lexicographically smaller allele-frequency vector (the reference's order among exact ties is
// unspecified: HashMap iteration, SURVEY quirk Q14 -- the oracle uses this same rule)
static __device__ __noinline__ bool map_tie_better(const DevModel &m, const TeamShared &S, double j, int eb, int p, int combo,
                           bool have, double bj, int beb, int bp, int bcombo) {
    (void)have; (void)j; (void)bj;
    if (eb != beb) return eb < beb;
    if (p == bp && combo == bcombo) return false;
    double a[kMaxSamples], b[kMaxSamples];
    decode_afs(m, S, m.paths[m.events[m.eb[eb].event].path_off + p], combo, a);
    decode_afs(m, S, m.paths[m.events[m.eb[beb].event].path_off + bp], bcombo, b);
    for (int smp = 0; smp < m.n_samples; ++smp)
        if (a[smp] != b[smp]) return a[smp] < b[smp];
    return false;
}

__device__ __forceinline__ bool map_better(const DevModel &m, const TeamShared &S, double j, int eb, int p,
                                           int combo, bool have, double bj, int beb, int bp, int bcombo) {
    if (!have) return true;
    if (j != bj) return j > bj;
    return map_tie_better(m, S, j, eb, p, combo, have, bj, beb, bp, bcombo);  // exact tie: rare
}

template <int TEAM>
__device__ __forceinline__ void team_sync() {
    if (TEAM == 1) __syncwarp();
    else __syncthreads();
}

__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Product of the per-observation terms T'_i of one likelihood key, checked variant: mantissas are
// multiplied (two chains), biased exponents summed in an integer; terms outside the normal range
// are handed to the log-space path.  finish(): ln prod = (sum e - 1023*count)*ln2 + ln(mantissas).
struct ProdAcc {
    double m0, m1, lnslow;
    int ex, nslow;
    __device__ __forceinline__ void init() { m0 = 1.0; m1 = 1.0; lnslow = 0.0; ex = 0; nslow = 0; }
    __device__ __forceinline__ bool add(double T, int i) {
        const int hi = __double2hiint(T);
        // sign 0 and 24 <= biased exponent < 2000 (normal, away from both ends)
        if ((unsigned)(hi - 0x01800000) >= (unsigned)(0x7d000000 - 0x01800000)) return false;
        ex += hi >> 20;
        const double mt = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(T));
        if (i & 1) m1 *= mt; else m0 *= mt;
        if ((i & 255) == 255) {  // keep the mantissa products inside fp64 range
            const double mm = m0 * m1;
            const int h2 = __double2hiint(mm);
            ex += (h2 >> 20) - 1023;
            m0 = __hiloint2double((h2 & 0x000fffff) | 0x3ff00000, __double2loint(mm));
            m1 = 1.0;
        }
        return true;
    }
    // exact ln T_i of an out-of-range term; sum_m already holds m_i, so it is taken out again
    __device__ __forceinline__ void slow(double lt, double mi) {
        lnslow += (mi == -INFINITY) ? lt : lt - mi;
        nslow++;
    }
    __device__ __forceinline__ double finish(double sum_m, int n) const {
        if (n == 0) return 0.0;  // empty pileup: fold identity ln 1 (likelihood.rs:135,227)
        const double lm = (double)(ex - 1023 * (n - nslow)) * 0.6931471805599453 + log_ni(m0 * m1);
        if (nslow == 0) return sum_m + lm;
        return (sum_m == -INFINITY) ? lnslow + lm : sum_m + lm + lnslow;
    }
};

// Unchecked variant for pileups whose terms are known to lie in [2^-240, 4] (k_i >= 2^-240, all
// factors finite): NB running products per thread, the exponent is pulled out of each product once
// per four observations (m in [1,2) times four terms stays above 2^-960).
constexpr double kTermMin = 5.659799424266695e-73;  // 2^-240
template <int NB>
struct ProdBlock {
    double m[NB];
    int ex[NB];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int b = 0; b < NB; ++b) { m[b] = 1.0; ex[b] = 0; }
    }
    __device__ __forceinline__ void renorm() {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int hi = __double2hiint(m[b]);
            ex[b] += hi >> 20;
            m[b] = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(m[b]));
        }
    }
    __device__ __forceinline__ double finish(int b, double sum_m, int nren) const {
        return sum_m + ((double)(ex[b] - 1023 * nren) * 0.6931471805599453 + log_ni(m[b]));
    }
};

// s_i(af): effective probability to sample the alt allele (likelihood.rs:25-38)
__device__ __forceinline__ double s_eff(double af, double sigma) {
    if (af == 1.0) return 1.0;
    double s = af * sigma;
    if (s > 1.0) s = s <= 1.0010005001667084 ? 1.0 : NAN;  // e^(1e-3) overshoot cap
    return s;
}

// TEAM = warps per site; MINB = CTAs per SM the register allocation must allow; NSC = number of
// samples when known at compile time (0: run-time) -- the per-sample loops of the integration then
// unroll and their small index arrays stay in registers instead of local memory; CONT = false
// leaves out the contaminated-sample paths (scenarios without contamination) -- the kernel is
// instruction-fetch sensitive, less code is measurably faster.
// PROG = false leaves out the prior program (set-only universes with a prior table, flat priors): the
// call sits in the innermost loop of the integration and costs 3-8 % on the small-site workloads even
// when it is never taken (registers of the allele-frequency vector, code size).
template <int TEAM, int MINB, int NSC, bool CONT, bool PROG>
__global__ void __launch_bounds__(kTeamThreadsMax, MINB) posterior_kernel(PostArgs g) {
    constexpr int NSA = NSC > 0 ? NSC : kMaxSamples;   // capacity of per-sample register arrays
    constexpr int NSU = NSC > 0 ? NSC : 1;              // unroll factor of per-sample loops
    constexpr int kTeamsPerCta = kTeamThreadsMax / (32 * TEAM);
    __shared__ TeamShared s_team[kTeamsPerCta];
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const int team_in_cta = warp_in_cta / TEAM;
    const int wt = warp_in_cta % TEAM;           // warp index within the team
    const int tt = wt * 32 + lane;               // thread index within the team
    constexpr int TT = TEAM * 32;
    TeamShared &S = s_team[team_in_cta];
    // per-team staging of one sample's observations: OC = {cR, kk, sigma, m_i}, BC = beta*cA per cfg
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int obs_cap = g.obs_cap;
    const size_t per_team = (size_t)obs_cap * sizeof(double4) +
                            (size_t)obs_cap * m_ncfg_cap(g) * sizeof(double) +
                            (size_t)(g.gs_cap + g.tab_cap) * sizeof(double) + (size_t)g.seg_cap * sizeof(int) +
                            (size_t)g.val_cap * 2 * sizeof(double);
    double4 *OC = (double4 *)(s_dyn + (size_t)team_in_cta * per_team);
    double *BC = (double *)(OC + obs_cap);
    double *GS = BC + (size_t)obs_cap * m_ncfg_cap(g);  // blocked contaminated path (may be empty)
    const int team_global = blockIdx.x * kTeamsPerCta + team_in_cta;
    // likelihood table of the site: shared memory when it fits, else per-team global scratch
    double *table = g.tab_cap ? GS + g.gs_cap : g.table + (int64_t)team_global * g.table_stride;
    int *seg_cnt = (int *)(GS + g.gs_cap + g.tab_cap);  // grid combinations per segment (0 = gated out)
    __shared__ double s_part[kTeamsPerCta][TEAM][2][kMaxEvents];  // per-warp (max, sum) of every event
    if (tt == 0) {
        S.val = (double *)(seg_cnt + g.seg_cap);
        S.lnw = S.val + g.val_cap;
    }
    const DevModel &m = g.m;
    const int NS = NSC > 0 ? NSC : m.n_samples;

    #pragma unroll 1
    for (;;) {
        // ---- next site (dynamic; chunks of consecutive sites and a prefetch of the following
        // site's rows were measured and lose 2-8 % on the C2 / C5 workloads)
        __shared__ int64_t s_site[kTeamsPerCta];
        team_sync<TEAM>();
        if (tt == 0) s_site[team_in_cta] = (int64_t)atomicAdd(g.cursor, 1);
        team_sync<TEAM>();
        const int64_t site = s_site[team_in_cta];
        if (site >= g.n_sites) break;

        // ---- 0. reset counters
        #pragma unroll 1
        for (int k = tt; k < (int)(offsetof(TeamShared, spec_off) / 4); k += TT) ((int *)&S)[k] = 0;
        if (tt == 0) S.overflow = 0;
        team_sync<TEAM>();

        // ---- 1. per-sample counters for grid sizes, bias learning and gating
        #pragma unroll 1
        for (int smp = wt; smp < NS; smp += TEAM) {
            const int64_t o0 = g.obs_off[site * NS + smp], o1 = g.obs_off[site * NS + smp + 1];
            const int n = (int)(o1 - o0);
            int c_strong = 0, c_ev[EV_N], a_ev[EV_N], a_ps[PS_N], c_unc = 0, c_so = 0;
            double sm = 0.0;
#pragma unroll
            for (int v = 0; v < EV_N; ++v) { c_ev[v] = 0; a_ev[v] = 0; }
#pragma unroll
            for (int v = 0; v < PS_N; ++v) a_ps[v] = 0;
            #pragma unroll 1
            for (int i = lane; i < n; i += 32) {
                const int64_t row = o0 + i;
                const unsigned cat = g.obs.cat[row];
                const double pa = g.obs.prob_alt[row], pr = g.obs.prob_ref[row];
                const bool strong = strong_obs(g.obs.prob_mapping[row], pa, pr);
                const unsigned st = VLR_CAT_STRAND(cat), orr = VLR_CAT_ORIENT(cat);
                // "bias evidence" = prob(obs) != ln 0 for the biased variant of each component
                const bool ev[EV_N] = {st == 0, st == 1, orr != 1, orr != 0,
                                       VLR_CAT_READPOS(cat) == 0, VLR_CAT_SOFTCLIP(cat) != 0,
                                       VLR_CAT_INDEL(cat) == 1};
#pragma unroll
                for (int v = 0; v < EV_N; ++v) { a_ev[v] |= ev[v]; c_ev[v] += (strong && ev[v]); }
                c_strong += strong;
                c_so += strong && ev[EV_DI];
                c_unc += orr > 1;
                // unbiased components: StrandBias::None prob = double_overlap | ln.5+single_overlap
                const double sn = st == 2 ? g.obs.prob_double_overlap[row]
                                          : -0.6931471805599453 + g.obs.prob_single_overlap[row];
                a_ps[PS_STRAND_NONE] |= sn != -INFINITY;
                a_ps[PS_RP_NONE] |= g.obs.prob_hit_base[row] != -INFINITY;
                a_ps[PS_DI_NONE] |= VLR_CAT_INDEL(cat) != 1;
                a_ps[PS_ANYOBS] = 1;
                const double mi = fmax(fmax(pa, pr), g.obs.prob_missed_allele[row]);
                sm += mi;
            }
            c_strong = __reduce_add_sync(0xffffffffu, c_strong);
            c_unc = __reduce_add_sync(0xffffffffu, c_unc);
            c_so = __reduce_add_sync(0xffffffffu, c_so);
            sm = warp_sum(sm);
#pragma unroll
            for (int v = 0; v < EV_N; ++v) {
                c_ev[v] = __reduce_add_sync(0xffffffffu, c_ev[v]);
                a_ev[v] = __reduce_or_sync(0xffffffffu, a_ev[v]);
            }
#pragma unroll
            for (int v = 0; v < PS_N; ++v) a_ps[v] = __reduce_or_sync(0xffffffffu, a_ps[v]);
            if (lane == 0) {
                S.n_obs[smp] = n;
                S.strong_all[smp] = c_strong;
                S.sum_m[smp] = sm;
                for (int v = 0; v < EV_N; ++v) {
                    S.strong_ev[smp][v] = c_ev[v];
                    if (a_ev[v]) atomicOr(&S.any_ev[v], 1);
                }
                for (int v = 0; v < PS_N; ++v) if (a_ps[v]) atomicOr(&S.any_ps[v], 1);
                atomicAdd(&S.n_uncertain, c_unc);
                atomicAdd(&S.n_total, n);
                atomicAdd(&S.strong_other_total, c_so);
                atomicAdd(&S.strong_all_total, c_strong);
                int gp = max(n + 1, 5);                      // generic.rs:109-122
                gp = min(gp, m.resolution[smp]);
                if ((gp & 1) == 0) gp += 1;
                S.grid[smp] = gp;
                if (n > g.obs_cap) S.overflow = 1;
            }
        }
        team_sync<TEAM>();

        // ---- 2. allele-frequency values of every spectrum (grouped by sample) + Simpson ln-weights
        if (tt == 0) {
            int off = 0;
            #pragma unroll 1
            for (int smp = 0; smp < NS; ++smp) {
                S.sample_voff[smp] = off;
                #pragma unroll 1
                for (int sp = 0; sp < m.n_spectra; ++sp) {
                    if (m.spectra[sp].sample != smp) continue;
                    S.spec_off[sp] = off;
                    off += m.spectra[sp].kind == VLR_NODE_RANGE ? S.grid[smp] : m.spectra[sp].n_vafs;
                }
                S.sample_nv[smp] = off - S.sample_voff[smp];
            }
            S.sample_voff[NS] = off;
            if (off > g.val_cap) S.overflow = 1;
            // table layout: per sample, per cfg, per key
            int64_t toff = 0;
            #pragma unroll 1
            for (int smp = 0; smp < NS; ++smp) {
                const int64_t keys = (CONT && m.contaminated[smp])
                                         ? (int64_t)S.sample_nv[smp] * S.sample_nv[m.by[smp]]
                                         : (int64_t)S.sample_nv[smp];
                S.tab_off[smp] = toff;
                S.tab_cfg_stride[smp] = keys;
                toff += keys * m.n_cfg;
            }
            if (toff > g.table_stride) S.overflow = 1;
        }
        team_sync<TEAM>();
        if (S.overflow) {
            if (tt == 0) {
                atomicAdd(g.status, 1);
                #pragma unroll 1
                for (int e = 0; e < m.n_events; ++e) g.out_event_ln[site * m.n_events + e] = NAN;
                g.out_marginal[site] = NAN;
                #pragma unroll 1
                for (int smp = 0; smp < NS; ++smp) {
                    g.out_map_af[site * NS + smp] = NAN;
                    g.out_map_bias[site * NS + smp] = -1;
                }
            }
            continue;
        }
        #pragma unroll 1
        for (int sp = wt; sp < m.n_spectra; sp += TEAM) {
            const DevSpectrum spc = m.spectra[sp];
            const int base = S.spec_off[sp];
            if (spc.kind == VLR_NODE_RANGE) {
                const double *r = m.vafs + spc.vaf_off;
                const int nobs = S.n_obs[spc.sample], n = S.grid[spc.sample];
                const double a = observable_min_d(r[0], r[1], r[2] != 0.0, r[3] != 0.0, nobs);
                const double b = observable_max_d(r[1], r[3] != 0.0, nobs);
                const double step = (b - a) / (double)(n - 1);  // itertools_num::linspace
                const double lnc = log_ni(b - a) - log_ni((double)(n - 1)) - 1.0986122886681098;
                #pragma unroll 1
                for (int i = lane; i < n; i += 32) {
                    // no FMA contraction: the grid point must be bit-identical to linspace's a + step*i
                    double x = __dadd_rn(a, __dmul_rn(step, (double)i)), w = 0.0;
                    if (i == 0) x = a;
                    else if (i == n - 1) x = b;
                    else w = (i & 1) ? 1.3862943611198906 : 0.6931471805599453;  // ln 4, ln 2 (Appendix A.1)
                    S.val[base + i] = x;
                    S.lnw[base + i] = w + lnc;
                }
            } else {
                #pragma unroll 1
                for (int i = lane; i < spc.n_vafs; i += 32) {
                    S.val[base + i] = m.vafs[spc.vaf_off + i];
                    S.lnw[base + i] = 0.0;
                }
            }
        }
        // ---- 3. bias learning + gating (one thread per configuration)
        #pragma unroll 1
        for (int c = tt; c < m.n_cfg; c += TT) {
            const vlr_biases b = m.cfg[c];
            // DivIndelBias::learn_parameters (divindel_bias.rs:98-132)
            double orate = 0.0;
            if (b.divindel) {
                const double r = S.strong_all_total > 0
                                     ? (double)S.strong_other_total / (double)S.strong_all_total : 0.0;
                orate = fmax(r, b.min_other_rate);
            }
            S.other_rate[c] = orate;
            bool ok = true;
            // is_possible (bias/mod.rs:30-36, divindel_bias.rs:77-84)
            ok &= b.strand == 0 ? S.any_ps[PS_STRAND_NONE] : S.any_ev[b.strand == 1 ? EV_SF : EV_SR];
            ok &= b.orientation == 0 ? S.any_ps[PS_ANYOBS]
                                     : S.any_ev[b.orientation == 1 ? EV_OF1R2 : EV_OF2R1];
            ok &= b.read_position == 0 ? S.any_ps[PS_RP_NONE] : S.any_ev[EV_RP];
            ok &= b.softclip == 0 ? S.any_ps[PS_ANYOBS] : S.any_ev[EV_SC];
            ok &= b.divindel == 0 ? S.any_ps[PS_DI_NONE] : S.any_ev[EV_DI];
            // is_informative (read_orientation_bias.rs:42-68, softclip_bias.rs:35-43, divindel_bias.rs:65-75)
            if (b.orientation) ok &= (double)S.n_uncertain < (double)S.n_total / 2.0;
            if (b.softclip) ok &= S.any_ev[EV_SC];
            if (b.divindel) ok &= S.any_ev[EV_DI];
            // is_likely (bias/mod.rs:46-67): any sample with < 10 strong obs or enough evidence
            const int comp_ev[5] = {b.strand == 0 ? -1 : (b.strand == 1 ? EV_SF : EV_SR),
                                    b.orientation == 0 ? -1 : (b.orientation == 1 ? EV_OF1R2 : EV_OF2R1),
                                    b.read_position == 0 ? -1 : EV_RP, b.softclip == 0 ? -1 : EV_SC,
                                    b.divindel == 0 ? -1 : EV_DI};
            for (int k = 0; k < 5; ++k) {
                if (comp_ev[k] < 0) continue;
                bool likely = false;
                #pragma unroll 1
                for (int smp = 0; smp < NS; ++smp) {
                    const int sa = S.strong_all[smp];
                    if (sa >= 10) {
                        const double ratio = (double)S.strong_ev[smp][comp_ev[k]] / (double)sa;
                        const double need = k == 4 ? 0.66666 * orate : 0.66666;
                        if (ratio >= need) { likely = true; break; }
                    } else { likely = true; break; }
                }
                ok &= likely;
            }
            S.allowed[c] = ok;
        }
        team_sync<TEAM>();
        if (wt == 0) {  // compact the allowed configurations (kMaxCfg <= 32: one lane each)
            const bool al = lane < m.n_cfg && S.allowed[lane];
            const unsigned mask = __ballot_sync(0xffffffffu, al);
            if (al) S.alist[__popc(mask & ((1u << lane) - 1u))] = lane;
            if (lane == 0) S.n_allowed = __popc(mask);
        }

        // ---- 4. likelihood tables.  Per sample: the team stages the observation constants
        // (scaled linear space) and the per-(observation, bias cfg) factor beta*cA in shared memory,
        // then every thread owns (key, cfg) pairs and walks the observations: mantissa product +
        // exponent sum per pair, ONE log per pair, no cross-lane traffic.
        #pragma unroll 1
        for (int smp = 0; smp < NS; ++smp) {
            const int64_t o0 = g.obs_off[site * NS + smp];
            const int n = S.n_obs[smp];
            const bool cont = CONT && m.contaminated[smp] != 0;
            const double rho = m.purity[smp];
            const int nv = S.sample_nv[smp], v0 = S.sample_voff[smp];
            const int byv0 = cont ? S.sample_voff[m.by[smp]] : 0;
            const int bynv = cont ? S.sample_nv[m.by[smp]] : 1;
            const int nkeys = nv * bynv;
            const int ncfg = m.n_cfg;
            team_sync<TEAM>();  // previous sample's shared constants are no longer read
            if (tt == 0) S.slow = 0;
            team_sync<TEAM>();
            bool fast_ok = true;
            #pragma unroll 1
            for (int i = tt; i < n; i += TT) {
                const int64_t row = o0 + i;
                const double pa = g.obs.prob_alt[row], pr = g.obs.prob_ref[row],
                             pmiss = g.obs.prob_missed_allele[row];
                const double mi = fmax(fmax(pa, pr), pmiss);
                const double pm = exp_ni(g.obs.prob_mapping[row]), pmm = exp_ni(g.obs.prob_mismapping[row]);
                const double hb = exp_ni(g.obs.prob_hit_base[row]);
                const double bbar = 0.25 * hb;
                double cA = 0.0, cR = 0.0, kk = 0.0;  // mi == -inf: T == 0, resolved by the exact path
                if (mi != -INFINITY) {
                    cA = pm * exp_ni(pa - mi);
                    cR = pm * exp_ni(pr - mi) * bbar;
                    kk = pmm * exp_ni(pmiss - mi) * bbar;
                }
                const unsigned cat = g.obs.cat[row];
                const double sN = VLR_CAT_STRAND(cat) == 2 ? exp_ni(g.obs.prob_double_overlap[row])
                                                           : 0.5 * exp_ni(g.obs.prob_single_overlap[row]);
                const double sigma = exp_ni(g.obs.prob_sample_alt[row]);
                OC[i] = make_double4(cR, kk, sigma, mi);
                // unchecked products need every term T_i in [2^-240, 4]: k_i bounds it from below
                fast_ok &= kk >= kTermMin && kk <= 2.0 && cR >= 0.0 && cR <= 2.0 && sigma >= 0.0 && sigma <= 1.0;
                #pragma unroll 1
                for (int c = 0; c < ncfg; ++c) {
                    const double b = S.allowed[c] ? bias_lin(m.cfg[c], S.other_rate[c], cat, sN, hb) * cA : 0.0;
                    BC[i * ncfg + c] = b;
                    fast_ok &= b >= 0.0 && b <= 2.0;  // false for NaN (configurations the reference panics on)
                }
            }
            if (!fast_ok) S.slow = 1;
            team_sync<TEAM>();
            const double sum_m = S.sum_m[smp];
            const int na = S.n_allowed;
            const bool fast = S.slow == 0;
            const int VBP = (bynv + 7) & ~7;
            double *tab = table + S.tab_off[smp];
            const int64_t tcs = S.tab_cfg_stride[smp];
            if (fast && cont && (int64_t)na * n * VBP <= g.gs_cap) {
                // ---- 4a. contaminated sample, blocked: thread = (primary value, cfg), eight
                // secondary values in registers.  GS[ca][i][vb] = (1-rho)*F_i(af_vb) + k_i is staged
                // once; the primary part rho*F_i(af_vi) is computed once per observation and shared
                // by the eight products: 2.5 fp64 instructions per evaluation.
                #pragma unroll 1
                for (int idx = tt; idx < na * n * VBP; idx += TT) {
                    const int vb = idx % VBP, r = idx / VBP;
                    const int i = r % n, ca = r / n;
                    double v = 1.0;  // padding columns: harmless factor
                    if (vb < bynv) {
                        const double4 oc = OC[i];
                        const double af = S.val[byv0 + vb];
                        const double z = af == 1.0 ? 1.0 : af * oc.z;
                        const double F = fma(1.0 - z, oc.x, z * BC[i * ncfg + S.alist[ca]]);
                        v = fma(1.0 - rho, F, oc.y);
                    }
                    GS[idx] = v;
                }
                team_sync<TEAM>();
                const int nvp = (nv + 31) & ~31;  // warps are uniform in cfg: GS reads broadcast
                #pragma unroll 1
                for (int q = tt; q < na * nvp; q += TT) {
                    const int ca = q / nvp, vi = q - ca * nvp;
                    if (vi >= nv) continue;
                    const int c = S.alist[ca];
                    const double af = S.val[v0 + vi];
                    const double afm = af == 1.0 ? 0.0 : af, afa = af == 1.0 ? 1.0 : 0.0;  // z = afm*sigma + afa
                    const double *bc = BC + c;
                    #pragma unroll 1
                    for (int vb0 = 0; vb0 < VBP; vb0 += 8) {
                        const double *gs = GS + (int64_t)ca * n * VBP + vb0;
                        ProdBlock<8> P;
                        P.init();
                        int nren = 0;
                        auto body = [&](int i) {
                            const double4 oc = OC[i];
                            const double z = fma(afm, oc.z, afa);
                            const double gp = rho * fma(1.0 - z, oc.x, z * bc[i * ncfg]);
                            const double2 *gv = (const double2 *)(gs + i * VBP);
                            const double2 g0 = gv[0], g1 = gv[1], g2 = gv[2], g3 = gv[3];
                            P.m[0] *= gp + g0.x; P.m[1] *= gp + g0.y; P.m[2] *= gp + g1.x; P.m[3] *= gp + g1.y;
                            P.m[4] *= gp + g2.x; P.m[5] *= gp + g2.y; P.m[6] *= gp + g3.x; P.m[7] *= gp + g3.y;
                        };
                        int i = 0;
                        #pragma unroll 1
                        for (; i + 4 <= n; i += 4) {
                            body(i); body(i + 1); body(i + 2); body(i + 3);
                            P.renorm(); ++nren;
                        }
                        if (i < n) {
                            #pragma unroll 1
                            for (; i < n; ++i) body(i);
                            P.renorm(); ++nren;
                        }
                        #pragma unroll
                        for (int b = 0; b < 8; ++b)
                            if (vb0 + b < bynv)
                                tab[(int64_t)c * tcs + (int64_t)vi * bynv + vb0 + b] = P.finish(b, sum_m, nren);
                    }
                }
            } else if (fast) {
                // ---- 4b. unchecked direct evaluation: thread = (NVB primary values, secondary
                // value, cfg); exponent extraction once per four observations.  NVB = 4 amortises
                // the shared-memory reads; small tables (few values) take NVB = 1 to fill the lanes.
                auto direct = [&](auto nvb_tag) {
                    constexpr int NVB = decltype(nvb_tag)::value;
                    const int ng = (nv + NVB - 1) / NVB;
                    const int nitems = na * bynv * ng;
                    #pragma unroll 1
                    for (int q = tt; q < nitems; q += TT) {
                        const int vg = q % ng, r = q / ng;
                        const int vb = r % bynv, c = S.alist[r / bynv];
                        double afm[NVB], afa[NVB];
                        #pragma unroll
                        for (int k = 0; k < NVB; ++k) {
                            const double af = S.val[v0 + min(vg * NVB + k, nv - 1)];
                            afm[k] = af == 1.0 ? 0.0 : af;
                            afa[k] = af == 1.0 ? 1.0 : 0.0;
                        }
                        const double af_s = cont ? S.val[byv0 + vb] : 0.0;
                        const double asm_ = af_s == 1.0 ? 0.0 : af_s, asa = af_s == 1.0 ? 1.0 : 0.0;
                        const double rho1 = 1.0 - rho;
                        const double *bc = BC + c;
                        ProdBlock<NVB> P;
                        P.init();
                        int nren = 0;
                        auto body = [&](int i) {
                            const double4 oc = OC[i];
                            const double bcA = bc[i * ncfg];
                            const double z2 = cont ? rho1 * fma(asm_, oc.z, asa) : 0.0;
                            #pragma unroll
                            for (int k = 0; k < NVB; ++k) {
                                double z = fma(afm[k], oc.z, afa[k]);
                                if (cont) z = fma(rho, z, z2);
                                P.m[k] *= fma(1.0 - z, oc.x, fma(z, bcA, oc.y));
                            }
                        };
                        int i = 0;
                        #pragma unroll 1
                        for (; i + 4 <= n; i += 4) {
                            body(i); body(i + 1); body(i + 2); body(i + 3);
                            P.renorm(); ++nren;
                        }
                        if (i < n) {
                            #pragma unroll 1
                            for (; i < n; ++i) body(i);
                            P.renorm(); ++nren;
                        }
                        #pragma unroll
                        for (int k = 0; k < NVB; ++k)
                            if (vg * NVB + k < nv)
                                tab[(int64_t)c * tcs + (int64_t)(vg * NVB + k) * bynv + vb] = P.finish(k, sum_m, nren);
                    }
                };
                if (na * bynv * ((nv + 3) >> 2) >= TT) direct(std::integral_constant<int, 4>());
                else direct(std::integral_constant<int, 1>());
            } else {
                // ---- 4c. checked evaluation: every term is range-tested, terms outside the normal
                // fp64 range (zero, subnormal, NaN) are recomputed in log space
                const int npairs = nkeys * ncfg;
                #pragma unroll 1
                for (int q = tt; q < npairs; q += TT) {
                    const int c = q / nkeys, key = q - c * nkeys;
                    if (!S.allowed[c]) continue;  // gated out: never requested (generic.rs:258-262)
                    const int vi = key / bynv, vb = key - vi * bynv;
                    const double af_p = S.val[v0 + vi];
                    const double af_s = cont ? S.val[byv0 + vb] : 0.0;
                    ProdAcc acc;
                    acc.init();
#pragma unroll 2
                    for (int i = 0; i < n; ++i) {
                        const double4 oc = OC[i];            // cR, kk, sigma, m_i
                        const double bcA = BC[i * ncfg + c];
                        double z = s_eff(af_p, oc.z);
                        double zr = 1.0 - z;
                        if (cont) {
                            const double z2 = s_eff(af_s, oc.z);
                            zr = rho * zr + (1.0 - rho) * (1.0 - z2);
                            z = rho * z + (1.0 - rho) * z2;
                        }
                        if (!acc.add(fma(zr, oc.x, fma(z, bcA, oc.y)), i)) {
                            const double lt = ln_T_exact(g.obs, o0 + i, m.cfg[c], S.other_rate[c], af_p,
                                                         af_s, cont, rho);
                            acc.slow(lt, oc.w);
                        }
                    }
                    tab[(int64_t)c * tcs + key] = acc.finish(sum_m, n);
                }
            }
        }
        team_sync<TEAM>();

        // ---- 5. integrate: per event a log-sum-exp over (bias cfg, path, grid combination).
        // The (event, bias cfg, path) segments come flattened from the host in the reference's
        // visiting order, so all terms of the site form ONE sequence ordered by event.  Thread tt
        // takes terms tt, tt+TT, ... and keeps an online (max, sum) of the event it is in; when
        // lanes move on to another event, the pending partials of the warp are combined by one
        // segmented reduction keyed by event (lanes hold non-decreasing events) and merged into the
        // warp's per-event accumulators in shared memory -- the cost follows the number of rounds,
        // not the number of events.
        double bestJ[2] = {-INFINITY, -INFINITY};  // [0] any, [1] non-artifact
        int bestId[2][3] = {{0x7fffffff, 0, 0}, {0x7fffffff, 0, 0}};
        bool have[2] = {false, false};
        {
            const int NE = m.n_events;
            double *acc_mx = s_part[team_in_cta][wt][0], *acc_sum = s_part[team_in_cta][wt][1];
            #pragma unroll 1
            for (int e = lane; e < NE; e += 32) { acc_mx[e] = -INFINITY; acc_sum[e] = 0.0; }
            // 5a. first term of every segment at this site (a gated-out bias cfg has no terms):
            // seg_pre[k] = exclusive prefix of the segments' grid-combination counts
            int *seg_pre = seg_cnt;
            if (wt == 0) {
                unsigned carry = 0;
                #pragma unroll 1
                for (int k0 = 0; k0 < m.n_segs; k0 += 32) {
                    const int k = k0 + lane;
                    unsigned total = 0;
                    if (k < m.n_segs) {
                        const DevSeg *sg = &m.segs[k];
                        total = S.allowed[sg->cfg] ? 1u : 0u;
                        #pragma unroll NSU
                        for (int smp = 0; smp < NS; ++smp) {
                            const int nvs = sg->nv[smp];
                            total *= nvs < 0 ? (unsigned)S.grid[smp] : (unsigned)nvs;
                        }
                    }
                    unsigned incl = total;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += u;
                    }
                    if (k < m.n_segs) seg_pre[k] = (int)(carry + incl - total);
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) seg_pre[m.n_segs] = (int)carry;
            }
            team_sync<TEAM>();
            const unsigned NT = (unsigned)seg_pre[m.n_segs];
            // segmented flush of the warp's pending partials; `pend` marks lanes whose partial
            // (cur_e, mx, sum) is complete for now.  Keys are non-decreasing in the lane index.
            auto flush = [&](bool pend, int key, double mx, double sum) {
                const int kup = __shfl_up_sync(0xffffffffu, key, 1), kdn = __shfl_down_sync(0xffffffffu, key, 1);
                const bool head = lane == 0 || kup != key, tail = lane == 31 || kdn != key;
                const unsigned heads = __ballot_sync(0xffffffffu, head), tails = __ballot_sync(0xffffffffu, tail);
                const int first = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));   // head of my run
                const int last = lane + __ffs(tails >> lane) - 1;                    // tail of my run
                double M = pend ? mx : -INFINITY;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double u = __shfl_up_sync(0xffffffffu, M, o);
                    if (lane - o >= first) M = fmax(M, u);
                }
                M = __shfl_sync(0xffffffffu, M, last);
                double part = (!pend || mx == -INFINITY) ? 0.0 : sum * exp_ni(mx - M);
                if (pend && isnan(sum)) part = NAN;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double u = __shfl_up_sync(0xffffffffu, part, o);
                    if (lane - o >= first) part += u;
                }
                const bool anyp = (__ballot_sync(0xffffffffu, pend) >> first) & (0xffffffffu >> (31 - (last - first)));
                if (tail && anyp && key < NE) {
                    // merge (M, part) into the warp's accumulator of event `key`
                    const double M0 = acc_mx[key], S0 = acc_sum[key];
                    if (M == -INFINITY) {
                        if (isnan(part)) acc_sum[key] = NAN;
                    } else if (M0 == -INFINITY) {
                        acc_mx[key] = M;
                        acc_sum[key] = isnan(S0) ? NAN : part;
                    } else {
                        const double f = exp_ni(-fabs(M0 - M));
                        acc_mx[key] = fmax(M0, M);
                        acc_sum[key] = M0 >= M ? fma(part, f, S0) : fma(S0, f, part);
                    }
                }
                __syncwarp();
            };
            // 5b. the terms
            int sidx = 0, loaded = -1, cur_e = -1;
            double mx = -INFINITY, sum = 0.0;
            unsigned cnt[NSA];
            int base[NSA];
            int c = 0, ebi = 0, p = 0, seg_e = 0;
            bool art = false;
            const double *prior = nullptr;
            const unsigned rounds = (NT + TT - 1) / TT;
            // (one extra round in which every lane is past the last term flushes what is pending)
            #pragma unroll 1
            for (unsigned r = 0; r <= rounds; ++r) {
                const unsigned t = r * TT + tt;
                const bool valid = t < NT;
                int e_t = NE;  // sentinel: beyond the last term
                if (valid) {
                    while (t >= (unsigned)seg_pre[sidx + 1]) ++sidx;   // empty segments are skipped
                    if (sidx != loaded) {
                        const DevSeg *sg = &m.segs[sidx];
                        c = sg->cfg; ebi = sg->ebi; art = sg->art != 0; p = sg->path;
                        #pragma unroll NSU
                        for (int smp = 0; smp < NS; ++smp) {
                            const int nvs = sg->nv[smp];
                            cnt[smp] = nvs < 0 ? (unsigned)S.grid[smp] : (unsigned)nvs;
                            base[smp] = S.spec_off[sg->spec[smp]] - S.sample_voff[smp];  // value index within sample
                        }
                        prior = m.prior ? m.prior + sg->leaf_off : nullptr;
                        seg_e = sg->event;
                        loaded = sidx;
                    }
                    e_t = seg_e;
                }
                const bool moved = cur_e >= 0 && e_t != cur_e;
                if (__any_sync(0xffffffffu, moved)) {
                    flush(moved, cur_e < 0 ? NE : cur_e, mx, sum);  // lanes without a term yet: sentinel key
                    if (moved) { mx = -INFINITY; sum = 0.0; }
                }
                if (!valid) { if (moved) cur_e = NE; continue; }
                cur_e = e_t;
                const unsigned combo = t - (unsigned)seg_pre[sidx];
                unsigned rem = combo;
                int vidx[NSA];
                double lw = 0.0;
                // mixed radix, LAST sample fastest
                #pragma unroll NSU
                for (int smp = NS - 1; smp >= 0; --smp) {
                    const unsigned q = (NSC == 1) ? 0u : rem / cnt[smp];
                    vidx[smp] = base[smp] + (int)(rem - q * cnt[smp]);
                    rem = q;
                    lw += S.lnw[S.sample_voff[smp] + vidx[smp]];
                }
                double J = prior ? prior[combo] : 0.0;  // Prior::compute (prior.rs:788-805) via the host's table
                if (PROG && !prior && m.prog) {          // ... or evaluated here (universes with Ranges)
                    double afv[NSA];
                    #pragma unroll NSU
                    for (int smp = 0; smp < NS; ++smp) afv[smp] = S.val[S.sample_voff[smp] + vidx[smp]];
                    J = prior_prog_eval(m.prog, NS, afv);
                }
                #pragma unroll NSU
                for (int smp = 0; smp < NS; ++smp) {
                    int key = vidx[smp];
                    if (CONT && m.contaminated[smp]) {
                        int vby = 0;
                        if (NSC > 0) {  // pick the source's digit without dynamic register indexing
                            #pragma unroll
                            for (int s2 = 0; s2 < NSA; ++s2) if (s2 == m.by[smp]) vby = vidx[s2];
                        } else {
                            vby = vidx[m.by[smp]];
                        }
                        key = vidx[smp] * S.sample_nv[m.by[smp]] + vby;
                    }
                    J += table[S.tab_off[smp] + (int64_t)c * S.tab_cfg_stride[smp] + key];
                }
                // joint_probs bookkeeping for the MAP (ties: lowest (eb, path, combo))
                if (!isnan(J)) {
                    #pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        if (w == 1 && art) continue;
                        if (map_better(m, S, J, ebi, p, (int)combo, have[w], bestJ[w], bestId[w][0],
                                       bestId[w][1], bestId[w][2])) {
                            have[w] = true; bestJ[w] = J;
                            bestId[w][0] = ebi; bestId[w][1] = p; bestId[w][2] = (int)combo;
                        }
                    }
                }
                const double tv = J + lw;
                if (isnan(tv)) { sum = NAN; }
                else if (tv > mx) { sum = sum * exp_ni(mx - tv) + 1.0; mx = tv; }
                else if (tv != -INFINITY) sum += exp_ni(tv - mx);
            }
            team_sync<TEAM>();
            #pragma unroll 1
            for (int e = tt; e < m.n_events; e += TT) {
                double M = -INFINITY;
#pragma unroll
                for (int k = 0; k < TEAM; ++k) M = fmax(M, s_part[team_in_cta][k][0][e]);
                double tot = 0.0;
#pragma unroll
                for (int k = 0; k < TEAM; ++k) {
                    const double pm_ = s_part[team_in_cta][k][0][e];
                    if (pm_ != -INFINITY) tot += s_part[team_in_cta][k][1][e] * (TEAM > 1 ? exp_ni(pm_ - M) : 1.0);
                    else if (isnan(s_part[team_in_cta][k][1][e])) tot = NAN;
                }
                S.ev_ln[e] = (M == -INFINITY && !isnan(tot)) ? -INFINITY : m.events[e].bias_prior + M + log_ni(tot);
            }
        }
        // ---- MAP reduction (ties -> lowest id)
        for (int w = 0; w < 2; ++w) {
            double j = have[w] ? bestJ[w] : -INFINITY;
            int i0 = have[w] ? bestId[w][0] : 0x7fffffff, i1 = bestId[w][1], i2 = bestId[w][2];
            int hv = have[w];
            for (int o = 16; o > 0; o >>= 1) {
                const double oj = __shfl_xor_sync(0xffffffffu, j, o);
                const int o0 = __shfl_xor_sync(0xffffffffu, i0, o), o1 = __shfl_xor_sync(0xffffffffu, i1, o),
                          o2 = __shfl_xor_sync(0xffffffffu, i2, o), oh = __shfl_xor_sync(0xffffffffu, hv, o);
                const bool take = oh && map_better(m, S, oj, o0, o1, o2, hv != 0, j, i0, i1, i2);
                if (take) { j = oj; i0 = o0; i1 = o1; i2 = o2; hv = 1; }
            }
            if (lane == 0) {
                S.best_j[wt][w] = hv ? j : -INFINITY;
                S.best_id[wt][w][0] = hv ? i0 : 0x7fffffff;
                S.best_id[wt][w][1] = i1; S.best_id[wt][w][2] = i2;
            }
        }
        team_sync<TEAM>();

        // ---- 6. outputs (first warp of the team, lanes over events): marginal, artifact event,
        // then one thread picks the MAP allele frequencies
        bool is_artifact = true;
        if (wt == 0) {
            const int NE = m.n_events;
            double M = -INFINITY;
            bool nan_any = false;
            #pragma unroll 1
            for (int e = lane; e < NE; e += 32) { const double v = S.ev_ln[e]; nan_any |= isnan(v); M = fmax(M, v); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) M = fmax(M, __shfl_xor_sync(0xffffffffu, M, o));
            nan_any = __any_sync(0xffffffffu, nan_any);
            // LogProb::ln_sum_exp: pmax (first maximum) + ln_1p(sum of the others)
            int imax = 0x7fffffff;
            #pragma unroll 1
            for (int e = lane; e < NE; e += 32) if (S.ev_ln[e] == M) imax = min(imax, e);
            imax = __reduce_min_sync(0xffffffffu, imax);
            double acc = 0.0;
            #pragma unroll 1
            for (int e = lane; e < NE; e += 32) {
                const double v = S.ev_ln[e];
                if (e != imax && v != -INFINITY && !nan_any && M != -INFINITY) acc += exp_ni(v - M);
                g.out_event_ln[site * NE + e] = v;
            }
            acc = warp_sum(acc);
            double marg;
            if (nan_any) marg = NAN;
            else if (M == -INFINITY) marg = -INFINITY;
            else marg = M + log1p_ni(acc);
            if (lane == 0) g.out_marginal[site] = marg;
            // artifact event = ln_sum_exp of artifact posteriors (calling.rs:581-598)
            double am = -INFINITY;
            #pragma unroll 1
            for (int e = lane; e < NE; e += 32)
                if (m.events[e].is_artifact) am = fmax(am, S.ev_ln[e] - marg);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
            double prob_art = -INFINITY;
            if (am != -INFINITY && !isnan(am)) {
                double acc2 = 0.0;
                #pragma unroll 1
                for (int e = lane; e < NE; e += 32)
                    if (m.events[e].is_artifact && S.ev_ln[e] - marg != -INFINITY)
                        acc2 += exp_ni(S.ev_ln[e] - marg - am);
                acc2 = warp_sum(acc2);
                prob_art = am + log_ni(acc2);
            }
            bool viol = false;
            #pragma unroll 1
            for (int e = lane; e < NE; e += 32)
                if (!m.events[e].is_artifact && !((S.ev_ln[e] - marg) < prob_art)) viol = true;
            is_artifact = !__any_sync(0xffffffffu, viol);
        }
        if (tt == 0) {
            const int w = is_artifact ? 0 : 1;
            double j = -INFINITY; int i0 = 0x7fffffff, i1 = 0, i2 = 0;
            for (int k = 0; k < TEAM; ++k) {
                const double oj = S.best_j[k][w];
                const int o0 = S.best_id[k][w][0], o1 = S.best_id[k][w][1], o2 = S.best_id[k][w][2];
                if (o0 == 0x7fffffff) continue;
                const bool take = map_better(m, S, oj, o0, o1, o2, i0 != 0x7fffffff, j, i0, i1, i2);
                if (take) { j = oj; i0 = o0; i1 = o1; i2 = o2; }
            }
            if (i0 == 0x7fffffff) {
                #pragma unroll 1
                for (int smp = 0; smp < NS; ++smp) {
                    g.out_map_af[site * NS + smp] = NAN;
                    g.out_map_bias[site * NS + smp] = -1;
                }
            } else {
                const int c = m.eb[i0].cfg;
                const bool art = m.cfg[c].strand | m.cfg[c].orientation | m.cfg[c].read_position |
                                 m.cfg[c].softclip | m.cfg[c].divindel;
                const DevPath &path = m.paths[m.events[m.eb[i0].event].path_off + i1];
                int64_t rem = i2;
                #pragma unroll 1
                for (int smp = NS - 1; smp >= 0; --smp) {
                    const DevSpectrum spc = m.spectra[path.spec[smp]];
                    const int cnt = spc.kind == VLR_NODE_RANGE ? S.grid[smp] : spc.n_vafs;
                    const int d = (int)(rem % cnt);
                    rem /= cnt;
                    // artifact MAP: allele frequency forced to 0 (calling.rs:659-663)
                    g.out_map_af[site * NS + smp] = art ? 0.0 : S.val[S.spec_off[path.spec[smp]] + d];
                    g.out_map_bias[site * NS + smp] = art ? c : -1;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- instantiation (posterior_inst_*.cu)
#define VLR_POSTERIOR_LAUNCHER(TEAM, MINB, NSC, CONT) launch_posterior_t##TEAM##_m##MINB##_n##NSC##_c##CONT
#define VLR_POSTERIOR_DECLARE(TEAM, MINB, NSC, CONT)                                              \
    cudaError_t VLR_POSTERIOR_LAUNCHER(TEAM, MINB, NSC, CONT)(int grid, size_t dyn_smem,          \
                                                              cudaStream_t st, const PostArgs &a);
#define VLR_POSTERIOR_INSTANTIATE(TEAM, MINB, NSC, CONT)                                          \
    cudaError_t VLR_POSTERIOR_LAUNCHER(TEAM, MINB, NSC, CONT)(int grid, size_t dyn_smem,          \
                                                              cudaStream_t st, const PostArgs &a) { \
        if (a.m.prog) {                                                                           \
            cudaError_t e = cudaFuncSetAttribute(posterior_kernel<TEAM, MINB, NSC, CONT != 0, true>, \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                                 (int)dyn_smem);                                  \
            if (e != cudaSuccess) return e;                                                       \
            posterior_kernel<TEAM, MINB, NSC, CONT != 0, true><<<grid, kTeamThreadsMax, dyn_smem, st>>>(a); \
        } else {                                                                                  \
            cudaError_t e = cudaFuncSetAttribute(posterior_kernel<TEAM, MINB, NSC, CONT != 0, false>, \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                                 (int)dyn_smem);                                  \
            if (e != cudaSuccess) return e;                                                       \
            posterior_kernel<TEAM, MINB, NSC, CONT != 0, false><<<grid, kTeamThreadsMax, dyn_smem, st>>>(a); \
        }                                                                                         \
        return cudaGetLastError();                                                                \
    }
// every (team, CTAs per SM, compile-time sample count, contamination) combination that is built
#define VLR_POSTERIOR_FOR_ALL(X)                                                                  \
    X(1, 7, 1, 0) X(1, 7, 2, 0) X(1, 7, 3, 0)                                                     \
    X(1, 6, 1, 0) X(1, 6, 2, 0) X(1, 6, 3, 0)                                                     \
    X(1, 5, 1, 0) X(1, 5, 2, 0) X(1, 5, 3, 0) X(1, 5, 2, 1) X(1, 5, 0, 1)                         \
    X(1, 3, 1, 0) X(1, 3, 2, 0) X(1, 3, 3, 0) X(1, 3, 2, 1) X(1, 3, 0, 1)                         \
    X(4, 3, 2, 1) X(4, 3, 0, 1)
VLR_POSTERIOR_FOR_ALL(VLR_POSTERIOR_DECLARE)

}  // namespace vlr
