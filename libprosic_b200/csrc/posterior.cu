// posterior.cu -- K3/K4: AF-grid pileup likelihood and fused event integration for sm_100a.
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
#include <math.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace vlr {

constexpr int kMaxSamples = 8;
constexpr int kMaxSpectra = 48;
constexpr int kMaxValues = 384;   // distinct allele-frequency values per site (all samples)
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
};
__device__ __forceinline__ int m_ncfg_cap(const PostArgs &g) { return g.m.n_cfg; }

// Out-of-line transcendentals: one copy each instead of one per call site (the fully inlined kernel
// was 343 KB of SASS and stalled on instruction fetch; ncu profile r1 `no_instruction`).
__device__ __noinline__ double exp_ni(double x) { return exp(x); }
__device__ __noinline__ double log_ni(double x) { return log(x); }
__device__ __noinline__ double log1p_ni(double x) { return log1p(x); }
__device__ __noinline__ double ln_one_minus_exp_d(double p) {
    return p < -0.693 ? log1p(-exp(p)) : log(-expm1(p));
}
__device__ __noinline__ double ln_add_exp2(double a, double b) {
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
__device__ __noinline__ double bias_ln(const vlr_biases &b, double other_rate, unsigned cat,
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
__device__ __noinline__ double lh_mapping_ln(double af, double psa_obs, double pb, double pab, double pa,
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
__device__ __noinline__ double ln_T_exact(const vlr_obs_columns &o, int64_t row, const vlr_biases &b,
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

// Bias::is_strong_obs (bias/mod.rs:79-83)
__device__ __forceinline__ bool strong_obs(double prob_mapping, double pa, double pr) {
    return prob_mapping >= -0.05129329438755058 && exp_ni(pa - pr) > 20.0;
}

// VAFRange::observable_{min,max} (grammar/formula.rs:1029-1071)
__device__ __noinline__ double observable_max_d(double end, bool rex, int n_obs) {
    if (n_obs < 10) return end;
    double c = (double)n_obs * end;
    if (rex && fmod(c, 1.0) == 0.0) c -= 1.0;
    return floor(c) / (double)n_obs;
}
__device__ __noinline__ double observable_min_d(double start, double end, bool lex, bool rex, int n_obs) {
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
    double val[kMaxValues];
    double lnw[kMaxValues];
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
__device__ __noinline__ void decode_afs(const DevModel &m, const TeamShared &S, const DevPath &path, int64_t combo,
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
__device__ __noinline__ bool map_tie_better(const DevModel &m, const TeamShared &S, double j, int eb, int p, int combo,
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

template <int TEAM>
__global__ void __launch_bounds__(kTeamThreadsMax, 3) posterior_kernel(PostArgs g) {
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
                            (size_t)(g.gs_cap + g.tab_cap) * sizeof(double) + (size_t)g.seg_cap * sizeof(int);
    double4 *OC = (double4 *)(s_dyn + (size_t)team_in_cta * per_team);
    double *BC = (double *)(OC + obs_cap);
    double *GS = BC + (size_t)obs_cap * m_ncfg_cap(g);  // blocked contaminated path (may be empty)
    const int team_global = blockIdx.x * kTeamsPerCta + team_in_cta;
    // likelihood table of the site: shared memory when it fits, else per-team global scratch
    double *table = g.tab_cap ? GS + g.gs_cap : g.table + (int64_t)team_global * g.table_stride;
    int *seg_cnt = (int *)(GS + g.gs_cap + g.tab_cap);  // grid combinations per segment (0 = gated out)
    __shared__ double s_part[kTeamsPerCta][TEAM][2][kMaxEvents];  // per-warp (max, sum) of every event
    const DevModel &m = g.m;
    const int NS = m.n_samples;

    #pragma unroll 1
    for (;;) {
        // ---- next site (dynamic)
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
            if (off > kMaxValues) S.overflow = 1;
            // table layout: per sample, per cfg, per key
            int64_t toff = 0;
            #pragma unroll 1
            for (int smp = 0; smp < NS; ++smp) {
                const int64_t keys = m.contaminated[smp]
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
        if (tt == 0) {
            int k = 0;
            #pragma unroll 1
            for (int c = 0; c < m.n_cfg; ++c) if (S.allowed[c]) S.alist[k++] = c;
            S.n_allowed = k;
        }

        // ---- 4. likelihood tables.  Per sample: the team stages the observation constants
        // (scaled linear space) and the per-(observation, bias cfg) factor beta*cA in shared memory,
        // then every thread owns (key, cfg) pairs and walks the observations: mantissa product +
        // exponent sum per pair, ONE log per pair, no cross-lane traffic.
        #pragma unroll 1
        for (int smp = 0; smp < NS; ++smp) {
            const int64_t o0 = g.obs_off[site * NS + smp];
            const int n = S.n_obs[smp];
            const bool cont = m.contaminated[smp] != 0;
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
        // visiting order; every thread takes grid combinations tt, tt+TT, ... of each segment and
        // keeps an online (max, sum) of the current event, flushed per warp at event boundaries.
        double bestJ[2] = {-INFINITY, -INFINITY};  // [0] any, [1] non-artifact
        int bestId[2][3] = {{0x7fffffff, 0, 0}, {0x7fffffff, 0, 0}};
        bool have[2] = {false, false};
        {
            double mx = -INFINITY, sum = 0.0;
            auto flush = [&](int e) {
                double M = mx;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) M = fmax(M, __shfl_xor_sync(0xffffffffu, M, o));
                double part = (mx == -INFINITY) ? 0.0 : sum * exp_ni(mx - M);
                if (isnan(sum)) part = NAN;
                part = warp_sum(part);
                if (lane == 0) {
                    s_part[team_in_cta][wt][0][e] = M;
                    s_part[team_in_cta][wt][1][e] = part;
                }
            };
            // 5a. grid combinations of every segment at this site (0 if its bias cfg is gated out)
            #pragma unroll 1
            for (int k = tt; k < m.n_segs; k += TT) {
                const DevSeg *sg = &m.segs[k];
                unsigned total = S.allowed[sg->cfg] ? 1u : 0u;
                #pragma unroll 1
                for (int smp = 0; smp < NS; ++smp) {
                    const int nvs = sg->nv[smp];
                    total *= nvs < 0 ? (unsigned)S.grid[smp] : (unsigned)nvs;
                }
                seg_cnt[k] = (int)total;
            }
            team_sync<TEAM>();
            // 5b. event by event: the lanes stride over ALL terms of the event (across its
            // segments), so that events made of many single-term segments still fill the lanes
            #pragma unroll 1
            for (int e = 0; e < m.n_events; ++e) {
                const int s_lo = m.events[e].seg_off, s_hi = s_lo + m.events[e].n_seg;
                mx = -INFINITY; sum = 0.0;
                int sidx = s_lo;
                unsigned seg_base = 0;          // first term of segment sidx within the event
                int loaded = -1;                // segment whose radix tables are in cnt[] / base[]
                unsigned cnt[kMaxSamples];
                int base[kMaxSamples];
                int c = 0, ebi = 0, p = -1;
                bool art = false;
                const double *prior = nullptr;
                #pragma unroll 1
                for (unsigned t = tt;; t += TT) {
                    while (sidx < s_hi && t >= seg_base + (unsigned)seg_cnt[sidx]) seg_base += (unsigned)seg_cnt[sidx++];
                    if (sidx >= s_hi) break;
                    if (sidx != loaded) {
                        const DevSeg *sg = &m.segs[sidx];
                        c = sg->cfg; ebi = sg->ebi; art = sg->art != 0;
                        if (sg->path != p) {  // radix tables depend on the tree path only
                            p = sg->path;
                            #pragma unroll 1
                            for (int smp = 0; smp < NS; ++smp) {
                                const int nvs = sg->nv[smp];
                                cnt[smp] = nvs < 0 ? (unsigned)S.grid[smp] : (unsigned)nvs;
                                base[smp] = S.spec_off[sg->spec[smp]] - S.sample_voff[smp];  // value index within sample
                            }
                            prior = m.prior ? m.prior + sg->leaf_off : nullptr;
                        }
                        loaded = sidx;
                    }
                    const unsigned combo = t - seg_base;
                    unsigned rem = combo;
                    int vidx[kMaxSamples];
                    double lw = 0.0;
                    // mixed radix, LAST sample fastest
                    #pragma unroll 1
                    for (int smp = NS - 1; smp >= 0; --smp) {
                        const unsigned q = rem / cnt[smp];
                        vidx[smp] = base[smp] + (int)(rem - q * cnt[smp]);
                        rem = q;
                        lw += S.lnw[S.sample_voff[smp] + vidx[smp]];
                    }
                    double J = prior ? prior[combo] : 0.0;  // Prior::compute (prior.rs:788-805) via the host's table
                    #pragma unroll 1
                    for (int smp = 0; smp < NS; ++smp) {
                        const int key = m.contaminated[smp] ? vidx[smp] * S.sample_nv[m.by[smp]] + vidx[m.by[smp]]
                                                            : vidx[smp];
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
                flush(e);
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

// ---------------------------------------------------------------- host: model flattening
struct HostModel {
    std::vector<DevSpectrum> spectra;
    std::vector<DevPath> paths;
    std::vector<DevEvent> events;
    std::vector<DevEventBias> eb;
    std::vector<vlr_biases> cfg;
    std::vector<double> vafs;
    std::vector<DevSeg> segs;
    std::vector<int64_t> path_leaf_off;  // per path: first leaf (prior-table entry), -1 if it binds a Range
    int64_t n_leaves = 0;                // leaves of all-Set paths
    bool any_range_path = false;
};

static bool same_cfg(const vlr_biases &a, const vlr_biases &b) {
    return a.strand == b.strand && a.orientation == b.orientation &&
           a.read_position == b.read_position && a.softclip == b.softclip &&
           a.divindel == b.divindel && a.min_other_rate == b.min_other_rate;
}

static int flatten_model(vlr_ctx *ctx, const vlr_model *mdl, HostModel &H) {
    const int NS = mdl->n_samples;
    if (NS < 1 || NS > kMaxSamples) return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_samples must be 1..8");
    if (mdl->n_events < 1 || mdl->n_events > kMaxEvents)
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_events must be 1..%d", kMaxEvents);
    // copy vafs referenced by nodes lazily: keep the caller's array layout
    int max_vaf = 0;
    for (int k = 0; k < mdl->n_nodes; ++k) {
        const vlr_node &nd = mdl->nodes[k];
        if (nd.kind == VLR_NODE_SET) max_vaf = std::max(max_vaf, nd.vaf_off + nd.n_vafs);
        if (nd.kind == VLR_NODE_RANGE) max_vaf = std::max(max_vaf, nd.vaf_off + 4);
    }
    H.vafs.assign(mdl->vafs, mdl->vafs + max_vaf);
    auto spectrum_of = [&](const vlr_node &nd) -> int {
        for (size_t s = 0; s < H.spectra.size(); ++s) {
            const DevSpectrum &sp = H.spectra[s];
            if (sp.sample != nd.sample || sp.kind != nd.kind) continue;
            const int n = nd.kind == VLR_NODE_RANGE ? 4 : nd.n_vafs;
            if (nd.kind == VLR_NODE_SET && sp.n_vafs != nd.n_vafs) continue;
            bool eq = true;
            for (int i = 0; i < n && eq; ++i) eq = H.vafs[sp.vaf_off + i] == H.vafs[nd.vaf_off + i];
            if (eq) return (int)s;
        }
        H.spectra.push_back(DevSpectrum{nd.sample, nd.kind, nd.vaf_off, nd.kind == VLR_NODE_RANGE ? 4 : nd.n_vafs});
        return (int)H.spectra.size() - 1;
    };
    for (int e = 0; e < mdl->n_events; ++e) {
        const vlr_event &ev = mdl->events[e];
        DevEvent de;
        de.path_off = (int)H.paths.size();
        de.eb_off = (int)H.eb.size();
        bool art = false;
        for (int k = 0; k < ev.n_biases; ++k) {
            const vlr_biases &b = mdl->biases[ev.bias_off + k];
            art |= b.strand || b.orientation || b.read_position || b.softclip || b.divindel;
            int id = -1;
            for (size_t c = 0; c < H.cfg.size(); ++c) if (same_cfg(H.cfg[c], b)) { id = (int)c; break; }
            if (id < 0) { H.cfg.push_back(b); id = (int)H.cfg.size() - 1; }
            H.eb.push_back(DevEventBias{e, id});
        }
        de.n_eb = ev.n_biases;
        de.is_artifact = art;
        de.bias_prior = art ? log(0.5) + log(1.0 / (double)ev.n_biases) : log(0.5);
        // depth-first path extraction (generic.rs:124-235): SET/RANGE bind a sample, SKIP passes,
        // FALSE kills the branch, >1 children fan out, no children = leaf
        struct Frame { int node; DevPath p; };
        std::vector<Frame> stack;
        for (int r = ev.n_roots - 1; r >= 0; --r) {
            Frame f;
            f.node = mdl->roots[ev.root_off + r];
            f.p.event = e;
            for (int s = 0; s < kMaxSamples; ++s) f.p.spec[s] = -1;
            stack.push_back(f);
        }
        // (stack order reversed above so that paths come out in root order)
        std::vector<DevPath> out;
        while (!stack.empty()) {
            Frame f = stack.back();
            stack.pop_back();
            if (f.node < 0 || f.node >= mdl->n_nodes) return set_err(ctx, VLR_ERR_INVALID, "bad node index");
            const vlr_node &nd = mdl->nodes[f.node];
            if (nd.kind == VLR_NODE_FALSE) continue;
            if (nd.kind == VLR_NODE_SET || nd.kind == VLR_NODE_RANGE) {
                if (nd.sample < 0 || nd.sample >= NS) return set_err(ctx, VLR_ERR_INVALID, "bad sample index");
                if (nd.kind == VLR_NODE_SET && nd.n_vafs < 1) return set_err(ctx, VLR_ERR_INVALID, "empty VAF set");
                f.p.spec[nd.sample] = spectrum_of(nd);
            } else if (nd.kind != VLR_NODE_SKIP) {
                return set_err(ctx, VLR_ERR_INVALID, "bad node kind");
            }
            if (nd.n_children == 0) {
                for (int s = 0; s < NS; ++s)
                    if (f.p.spec[s] < 0)
                        return set_err(ctx, VLR_ERR_UNSUPPORTED, "event %d: a tree path does not bind sample %d", e, s);
                out.push_back(f.p);
            } else {
                for (int c = nd.n_children - 1; c >= 0; --c) {
                    Frame g2 = f;
                    g2.node = mdl->children[nd.child_off + c];
                    stack.push_back(g2);
                }
            }
        }
        de.n_paths = (int)out.size();
        H.paths.insert(H.paths.end(), out.begin(), out.end());
        H.events.push_back(de);
    }
    // leaves (event, path, combination) in visiting order: the index space of the prior table
    H.path_leaf_off.assign(H.paths.size(), -1);
    for (size_t pi = 0; pi < H.paths.size(); ++pi) {
        int64_t combos = 1;
        bool all_set = true;
        for (int smp = 0; smp < NS; ++smp) {
            const DevSpectrum &sp = H.spectra[H.paths[pi].spec[smp]];
            if (sp.kind == VLR_NODE_RANGE) all_set = false;
            else combos *= sp.n_vafs;
        }
        if (all_set) { H.path_leaf_off[pi] = H.n_leaves; H.n_leaves += combos; }
        else H.any_range_path = true;
    }
    for (int e = 0; e < mdl->n_events; ++e) {
        DevEvent &de = H.events[e];
        de.seg_off = (int)H.segs.size();
        de.n_seg = de.n_eb * de.n_paths;
        for (int q = 0; q < de.n_eb; ++q)
            for (int pth = 0; pth < de.n_paths; ++pth) {
                DevSeg sg;
                memset(&sg, 0, sizeof(sg));
                sg.event = e; sg.ebi = de.eb_off + q; sg.cfg = H.eb[de.eb_off + q].cfg; sg.path = pth;
                const vlr_biases &b = H.cfg[sg.cfg];
                sg.art = (b.strand || b.orientation || b.read_position || b.softclip || b.divindel) ? 1 : 0;
                sg.leaf_off = H.path_leaf_off[de.path_off + pth];
                for (int smp = 0; smp < NS; ++smp) {
                    const int spi = H.paths[de.path_off + pth].spec[smp];
                    sg.spec[smp] = (int16_t)spi;
                    sg.nv[smp] = H.spectra[spi].kind == VLR_NODE_RANGE ? (int16_t)-1 : (int16_t)H.spectra[spi].n_vafs;
                    if (H.spectra[spi].kind != VLR_NODE_RANGE && H.spectra[spi].n_vafs > 32767)
                        return set_err(ctx, VLR_ERR_UNSUPPORTED, "VAF set with more than 32767 values");
                }
                H.segs.push_back(sg);
            }
    }
    if ((int)H.spectra.size() > kMaxSpectra) return set_err(ctx, VLR_ERR_UNSUPPORTED, "too many distinct VAF spectra");
    if ((int)H.cfg.size() > kMaxCfg) return set_err(ctx, VLR_ERR_UNSUPPORTED, "too many bias configurations");
    return VLR_OK;
}

// upper bound of the likelihood-table size (doubles) for one site
static int64_t table_bound(const vlr_model *mdl, const HostModel &H) {
    int64_t nv[kMaxSamples] = {0};
    for (const DevSpectrum &sp : H.spectra) {
        // grid = min(max(n_obs + 1, 5), resolution) made odd (generic.rs:109-122) <= resolution | 1
        const int64_t c = sp.kind == VLR_NODE_RANGE ? (int64_t)(mdl->resolution[sp.sample] | 1) : sp.n_vafs;
        nv[sp.sample] += c;
    }
    int64_t tot = 0;
    for (int s = 0; s < mdl->n_samples; ++s) {
        int64_t keys = mdl->contaminated[s] ? nv[s] * nv[mdl->by[s]] : nv[s];
        tot += keys * (int64_t)H.cfg.size();
    }
    return tot;
}

}  // namespace vlr

using namespace vlr;

extern "C" {

int vlr_posterior_batch_dev(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                            const vlr_obs_columns *d_obs, const int64_t *d_obs_off,
                            int32_t max_obs_per_pileup, double *d_out_event_ln,
                            double *d_out_marginal, double *d_out_map_af, int32_t *d_out_map_bias) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!model || !d_obs || !d_obs_off || !d_out_event_ln || !d_out_marginal || !d_out_map_af ||
        !d_out_map_bias || n_sites < 0 || n_sites > 0x7fffffff || max_obs_per_pileup < 0)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_sites == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    HostModel H;
    int rc = flatten_model(ctx, model, H);
    if (rc != VLR_OK) return rc;
    for (int s = 0; s < model->n_samples; ++s)
        if (model->contaminated[s] && (model->by[s] < 0 || model->by[s] >= model->n_samples || model->by[s] == s))
            return set_err(ctx, VLR_ERR_INVALID, "bad contamination source for sample %d", s);
    const int64_t tbound = table_bound(model, H);
    // team size: whole CTA for big tables (tumor/normal-like), one warp per site for small ones
    const int obs_cap = std::max(8, (max_obs_per_pileup + 7) & ~7);
    size_t team_smem = (size_t)obs_cap * sizeof(double4) + (size_t)obs_cap * H.cfg.size() * sizeof(double);
    // blocked contaminated path: cfgs x observations x (values of the contaminant, padded to 8)
    int64_t gs_cap = 0;
    {
        int64_t nvb[kMaxSamples] = {0};
        for (const DevSpectrum &sp : H.spectra) {
            const int64_t c = sp.kind == VLR_NODE_RANGE ? (int64_t)(model->resolution[sp.sample] | 1) : sp.n_vafs;
            nvb[sp.sample] += c;
        }
        for (int smp = 0; smp < model->n_samples; ++smp)
            if (model->contaminated[smp])
                gs_cap = std::max(gs_cap, (int64_t)H.cfg.size() * obs_cap * ((nvb[model->by[smp]] + 7) & ~(int64_t)7));
        if (team_smem + (size_t)gs_cap * 8 > 72 * 1024) gs_cap = 0;  // does not fit: direct evaluation
    }
    team_smem += (size_t)gs_cap * 8;
    // likelihood table of one site in shared memory when it is small enough
    const int64_t tab_cap = (tbound <= 4096 && team_smem + (size_t)tbound * 8 <= 72 * 1024)
                                ? ((tbound + 3) & ~(int64_t)3) : 0;  // keeps the teams' carve-outs 32-byte aligned
    team_smem += (size_t)tab_cap * 8;
    const int seg_cap = ((int)H.segs.size() + 7) & ~7;  // ints; keeps the carve-outs 32-byte aligned
    team_smem += (size_t)seg_cap * 4;
    const bool big = tbound > 2048 || 4 * team_smem > 64 * 1024;
    const int teams_per_cta = big ? 1 : 4;
    const size_t dyn_smem = (size_t)teams_per_cta * team_smem;
    if (dyn_smem > 160 * 1024)
        return set_err(ctx, VLR_ERR_UNSUPPORTED,
                       "%d observations x %d bias configurations do not fit the shared-memory staging",
                       max_obs_per_pileup, (int)H.cfg.size());
    // static shared memory: TeamShared (~8 KB per team) + event partials (4 KB)
    const size_t cta_smem = dyn_smem + (size_t)teams_per_cta * 8704 + 4096 + 1024;
    const int grid = ctx->sm_count * (int)std::max<size_t>(1, std::min<size_t>(3, (size_t)(220 * 1024) / cta_smem));
    const int64_t n_teams = (int64_t)grid * teams_per_cta;
    if (tbound > ((int64_t)1 << 24))
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "likelihood table of %lld entries per site is too large", (long long)tbound);
    const bool has_prior = model->prior_ln != nullptr;
    if (has_prior) {
        if (H.any_range_path)
            return set_err(ctx, VLR_ERR_UNSUPPORTED, "a prior table needs every tree path to bind VAF sets only "
                                                     "(grid points of a Range depend on the site's depth)");
        if (model->n_prior != H.n_leaves)
            return set_err(ctx, VLR_ERR_INVALID, "prior table has %lld entries, the model has %lld leaves",
                           (long long)model->n_prior, (long long)H.n_leaves);
    }
    enum { S_MODEL = 16, S_TABLE, S_CNT };
    // model blob
    const size_t b_spec = H.spectra.size() * sizeof(DevSpectrum), b_path = H.paths.size() * sizeof(DevPath),
                 b_ev = H.events.size() * sizeof(DevEvent), b_eb = H.eb.size() * sizeof(DevEventBias),
                 b_cfg = H.cfg.size() * sizeof(vlr_biases), b_vaf = H.vafs.size() * sizeof(double),
                 b_seg = H.segs.size() * sizeof(DevSeg), b_pri = has_prior ? (size_t)H.n_leaves * 8 : 0;
    size_t off[9];
    off[0] = 0;
    off[1] = align_up(off[0] + b_spec, 16);
    off[2] = align_up(off[1] + b_path, 16);
    off[3] = align_up(off[2] + b_ev, 16);
    off[4] = align_up(off[3] + b_eb, 16);
    off[5] = align_up(off[4] + b_cfg, 16);
    off[6] = align_up(off[5] + b_vaf, 16);
    off[7] = align_up(off[6] + b_seg, 16);
    off[8] = align_up(off[7] + b_pri, 16);
    uint8_t *h_blob = (uint8_t *)pin_scratch(ctx, 4, off[8] + 16);
    uint8_t *d_blob = (uint8_t *)dev_scratch(ctx, S_MODEL, off[8] + 16);
    double *d_table = (double *)dev_scratch(ctx, S_TABLE, (size_t)((tab_cap ? 0 : n_teams * tbound) + 8) * sizeof(double));
    int32_t *d_cnt = (int32_t *)dev_scratch(ctx, S_CNT, 256);
    if (!h_blob || !d_blob || !d_table || !d_cnt) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    // the pinned blob may still be in flight from a previous call on this stream
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(h_blob + off[0], H.spectra.data(), b_spec);
    memcpy(h_blob + off[1], H.paths.data(), b_path);
    memcpy(h_blob + off[2], H.events.data(), b_ev);
    memcpy(h_blob + off[3], H.eb.data(), b_eb);
    memcpy(h_blob + off[4], H.cfg.data(), b_cfg);
    memcpy(h_blob + off[5], H.vafs.data(), b_vaf);
    memcpy(h_blob + off[6], H.segs.data(), b_seg);
    if (has_prior) memcpy(h_blob + off[7], model->prior_ln, b_pri);
    VLR_CUDA(ctx, cudaMemcpyAsync(d_blob, h_blob, off[8], cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 256, st));
    PostArgs a;
    memset(&a, 0, sizeof(a));
    a.m.n_samples = model->n_samples;
    a.m.n_events = model->n_events;
    a.m.n_spectra = (int)H.spectra.size();
    a.m.n_paths = (int)H.paths.size();
    a.m.n_cfg = (int)H.cfg.size();
    a.m.n_eb = (int)H.eb.size();
    for (int s = 0; s < model->n_samples; ++s) {
        a.m.resolution[s] = model->resolution[s];
        a.m.contaminated[s] = model->contaminated[s];
        a.m.by[s] = model->by[s];
        a.m.purity[s] = model->purity[s];
        if (model->resolution[s] < 1) return set_err(ctx, VLR_ERR_INVALID, "resolution must be >= 1");
    }
    a.m.spectra = (const DevSpectrum *)(d_blob + off[0]);
    a.m.paths = (const DevPath *)(d_blob + off[1]);
    a.m.events = (const DevEvent *)(d_blob + off[2]);
    a.m.eb = (const DevEventBias *)(d_blob + off[3]);
    a.m.cfg = (const vlr_biases *)(d_blob + off[4]);
    a.m.vafs = (const double *)(d_blob + off[5]);
    a.m.segs = (const DevSeg *)(d_blob + off[6]);
    a.m.n_segs = (int)H.segs.size();
    a.m.prior = has_prior ? (const double *)(d_blob + off[7]) : nullptr;
    a.obs = *d_obs;
    a.obs_off = d_obs_off;
    a.n_sites = n_sites;
    a.cursor = d_cnt;
    a.status = d_cnt + 1;
    a.table = d_table;
    a.table_stride = tbound;
    a.out_event_ln = d_out_event_ln;
    a.out_marginal = d_out_marginal;
    a.out_map_af = d_out_map_af;
    a.out_map_bias = d_out_map_bias;
    a.obs_cap = obs_cap;
    a.gs_cap = gs_cap;
    a.tab_cap = tab_cap;
    a.seg_cap = seg_cap;
    if (big) {
        VLR_CUDA(ctx, cudaFuncSetAttribute(posterior_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        posterior_kernel<4><<<grid, kTeamThreadsMax, dyn_smem, st>>>(a);
    } else {
        VLR_CUDA(ctx, cudaFuncSetAttribute(posterior_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        posterior_kernel<1><<<grid, kTeamThreadsMax, dyn_smem, st>>>(a);
    }
    VLR_LAUNCH_CHECK(ctx);
    return VLR_OK;
}

int vlr_model_leaves(vlr_ctx *ctx, const vlr_model *model, int64_t capacity, int64_t *out_n_leaves,
                     int32_t *out_event, double *out_afs) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!model || !out_n_leaves) return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    HostModel H;
    int rc = flatten_model(ctx, model, H);
    if (rc != VLR_OK) return rc;
    if (H.any_range_path)
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "leaves are enumerable only when every tree path binds VAF sets");
    *out_n_leaves = H.n_leaves;
    if (!out_event && !out_afs) return VLR_OK;
    if (capacity < H.n_leaves) return set_err(ctx, VLR_ERR_INVALID, "capacity %lld < %lld leaves", (long long)capacity, (long long)H.n_leaves);
    const int NS = model->n_samples;
    for (int e = 0; e < model->n_events; ++e)
        for (int pth = 0; pth < H.events[e].n_paths; ++pth) {
            const int pi = H.events[e].path_off + pth;
            int64_t combos = 1;
            for (int smp = 0; smp < NS; ++smp) combos *= H.spectra[H.paths[pi].spec[smp]].n_vafs;
            for (int64_t combo = 0; combo < combos; ++combo) {
                const int64_t leaf = H.path_leaf_off[pi] + combo;
                if (out_event) out_event[leaf] = e;
                int64_t rem = combo;
                for (int smp = NS - 1; smp >= 0; --smp) {  // LAST sample is the fastest digit
                    const DevSpectrum &sp = H.spectra[H.paths[pi].spec[smp]];
                    if (out_afs) out_afs[leaf * NS + smp] = H.vafs[sp.vaf_off + rem % sp.n_vafs];
                    rem /= sp.n_vafs;
                }
            }
        }
    return VLR_OK;
}

// ---- host-buffer pipeline: the batch is cut into chunks of sites; H2D of chunk c+1 (copy
// stream) overlaps the kernels and D2H of chunk c (main stream), two sets of device buffers.
// T = double: the nine columns as given.  T = float: the seven stored columns (de-quantised
// MiniLogProb values are exactly representable in f32), widened on the device, where the two
// derived columns are computed as ObservationBuilder does (observation.rs:218-226).
}  // extern "C"

namespace vlr {

__global__ void widen_obs_kernel(int64_t n, const float *pm, const float *pa, const float *pr, const float *pmiss,
                                 const float *psa, const float *pdo, const float *phb, double *o_pm,
                                 double *o_pmm, double *o_pa, double *o_pr, double *o_pmiss, double *o_psa,
                                 double *o_pdo, double *o_pso, double *o_phb) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double m = (double)pm[k], d = (double)pdo[k];
        o_pm[k] = m;
        o_pmm[k] = ln_one_minus_exp_d(m);
        o_pa[k] = (double)pa[k];
        o_pr[k] = (double)pr[k];
        o_pmiss[k] = (double)pmiss[k];
        o_psa[k] = (double)psa[k];
        o_pdo[k] = d;
        o_pso[k] = ln_one_minus_exp_d(d);
        o_phb[k] = (double)phb[k];
    }
}

template <typename T>
static int posterior_pipeline(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model, const T *const *cols,
                              int n_cols_in, const uint16_t *cat, const int64_t *obs_off,
                              double *out_event_ln, double *out_marginal, double *out_map_af,
                              int32_t *out_map_bias) {
    if (!model || !cols || !cat || !obs_off || !out_event_ln || !out_marginal || !out_map_af || !out_map_bias ||
        n_sites < 0)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_sites == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int NS = model->n_samples, NE = model->n_events;
    if (NS < 1 || NS > kMaxSamples) return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_samples must be 1..8");
    if (NE < 1 || NE > kMaxEvents) return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_events must be 1..%d", kMaxEvents);
    const int64_t n_off = n_sites * NS + 1;
    const int64_t n_rows = obs_off[n_off - 1] - obs_off[0];
    int64_t max_obs = 0;
    for (int64_t k = 0; k + 1 < n_off; ++k) {
        const int64_t l = obs_off[k + 1] - obs_off[k];
        if (l < 0) return set_err(ctx, VLR_ERR_INVALID, "obs_off not monotone");
        if (l > max_obs) max_obs = l;
    }
    if (max_obs > 4096) return set_err(ctx, VLR_ERR_UNSUPPORTED, "pileup of %lld observations exceeds 4096", (long long)max_obs);
    for (int c = 0; c < n_cols_in; ++c)
        if (!cols[c] && n_rows > 0) return set_err(ctx, VLR_ERR_INVALID, "null observation column");
    int rc = ensure_copy_stream(ctx);
    if (rc != VLR_OK) return rc;
    // ---- chunk boundaries (in sites), ~24 MB of input per chunk
    const double in_bytes = (double)n_rows * (n_cols_in * sizeof(T) + 2) + (double)n_off * 8;
    int n_chunks = (int)std::min<double>(64.0, std::max<double>(1.0, ceil(in_bytes / (24.0 * 1048576.0))));
    if ((int64_t)n_chunks > n_sites) n_chunks = (int)n_sites;
    std::vector<int64_t> cs((size_t)n_chunks + 1);
    cs[0] = 0;
    for (int c = 1; c < n_chunks; ++c) {
        const int64_t want = obs_off[0] + n_rows * c / n_chunks;
        int64_t lo = cs[c - 1] + 1, hi = n_sites - (n_chunks - c);  // keep every chunk non-empty
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (obs_off[mid * NS] < want) lo = mid + 1; else hi = mid;
        }
        cs[c] = lo;
    }
    cs[n_chunks] = n_sites;
    int64_t max_rows = 0, max_sites = 0;
    for (int c = 0; c < n_chunks; ++c) {
        max_rows = std::max(max_rows, obs_off[cs[c + 1] * NS] - obs_off[cs[c] * NS]);
        max_sites = std::max(max_sites, cs[c + 1] - cs[c]);
    }
    // ---- two sets of device buffers
    constexpr bool kWiden = sizeof(T) == 4;
    enum { S_BASE = 100, S_PER = 24 };  // per set: 0..8 fp64 columns, 9 cat, 10 off, 11..14 outputs, 15..21 f32 staging
    double *d_cols[2][9];
    T *d_in[2][9];
    uint16_t *d_cat[2];
    int64_t *d_off[2];
    double *d_ev[2], *d_marg[2], *d_af[2];
    int32_t *d_mb[2];
    for (int b = 0; b < 2; ++b) {
        const int base = S_BASE + b * S_PER;
        for (int c = 0; c < 9; ++c) {
            d_cols[b][c] = (double *)dev_scratch(ctx, base + c, (size_t)(max_rows + 1) * 8);
            if (!d_cols[b][c]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        }
        for (int c = 0; c < n_cols_in; ++c) {
            d_in[b][c] = kWiden ? (T *)dev_scratch(ctx, base + 15 + c, (size_t)(max_rows + 1) * sizeof(T))
                                : (T *)d_cols[b][c];
            if (!d_in[b][c]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        }
        d_cat[b] = (uint16_t *)dev_scratch(ctx, base + 9, (size_t)(max_rows + 1) * 2);
        d_off[b] = (int64_t *)dev_scratch(ctx, base + 10, (size_t)(max_sites * NS + 1) * 8);
        d_ev[b] = (double *)dev_scratch(ctx, base + 11, (size_t)max_sites * NE * 8);
        d_marg[b] = (double *)dev_scratch(ctx, base + 12, (size_t)max_sites * 8);
        d_af[b] = (double *)dev_scratch(ctx, base + 13, (size_t)max_sites * NS * 8);
        d_mb[b] = (int32_t *)dev_scratch(ctx, base + 14, (size_t)max_sites * NS * 4);
        if (!d_cat[b] || !d_off[b] || !d_ev[b] || !d_marg[b] || !d_af[b] || !d_mb[b])
            return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    }
    // earlier work of this context may still read the buffers the copy stream is about to overwrite
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    cudaStream_t cp = ctx->copy_stream;
    for (int c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {  // H2D of chunk c
            const int b = c & 1;
            const int64_t s0 = cs[c], s1 = cs[c + 1];
            const int64_t r0 = obs_off[s0 * NS], rows = obs_off[s1 * NS] - r0;
            if (c >= 2) VLR_CUDA(ctx, cudaStreamWaitEvent(cp, ctx->ev_done[b], 0));
            for (int k = 0; k < n_cols_in && rows > 0; ++k)
                VLR_CUDA(ctx, cudaMemcpyAsync(d_in[b][k], cols[k] + r0, (size_t)rows * sizeof(T), cudaMemcpyHostToDevice, cp));
            if (rows > 0) VLR_CUDA(ctx, cudaMemcpyAsync(d_cat[b], cat + r0, (size_t)rows * 2, cudaMemcpyHostToDevice, cp));
            VLR_CUDA(ctx, cudaMemcpyAsync(d_off[b], obs_off + s0 * NS, (size_t)((s1 - s0) * NS + 1) * 8, cudaMemcpyHostToDevice, cp));
            VLR_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[b], cp));
        }
        if (c >= 1) {  // kernels + D2H of chunk c-1
            const int b = (c - 1) & 1;
            const int64_t s0 = cs[c - 1], s1 = cs[c];
            const int64_t r0 = obs_off[s0 * NS], rows = obs_off[s1 * NS] - r0;
            VLR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_h2d[b], 0));
            if (kWiden && rows > 0) {
                const float *const *f = (const float *const *)d_in[b];
                const int grid = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)ctx->sm_count * 8);
                widen_obs_kernel<<<grid, 256, 0, st>>>(rows, f[0], f[1], f[2], f[3], f[4], f[5], f[6], d_cols[b][0],
                                                       d_cols[b][1], d_cols[b][2], d_cols[b][3], d_cols[b][4],
                                                       d_cols[b][5], d_cols[b][6], d_cols[b][7], d_cols[b][8]);
                VLR_LAUNCH_CHECK(ctx);
            }
            vlr_obs_columns dc;  // offsets stay absolute: rebase the device pointers
            dc.prob_mapping = d_cols[b][0] - r0; dc.prob_mismapping = d_cols[b][1] - r0; dc.prob_alt = d_cols[b][2] - r0;
            dc.prob_ref = d_cols[b][3] - r0; dc.prob_missed_allele = d_cols[b][4] - r0; dc.prob_sample_alt = d_cols[b][5] - r0;
            dc.prob_double_overlap = d_cols[b][6] - r0; dc.prob_single_overlap = d_cols[b][7] - r0;
            dc.prob_hit_base = d_cols[b][8] - r0; dc.cat = d_cat[b] - r0;
            rc = vlr_posterior_batch_dev(ctx, s1 - s0, model, &dc, d_off[b], (int32_t)max_obs, d_ev[b], d_marg[b],
                                         d_af[b], d_mb[b]);
            if (rc != VLR_OK) return rc;
            const int64_t ns_ = s1 - s0;
            VLR_CUDA(ctx, cudaMemcpyAsync(out_event_ln + s0 * NE, d_ev[b], (size_t)ns_ * NE * 8, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaMemcpyAsync(out_marginal + s0, d_marg[b], (size_t)ns_ * 8, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaMemcpyAsync(out_map_af + s0 * NS, d_af[b], (size_t)ns_ * NS * 8, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaMemcpyAsync(out_map_bias + s0 * NS, d_mb[b], (size_t)ns_ * NS * 4, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], st));
        }
    }
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    return VLR_OK;
}

}  // namespace vlr

extern "C" {

int vlr_posterior_batch(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                        const vlr_obs_columns *obs, const int64_t *obs_off, double *out_event_ln,
                        double *out_marginal, double *out_map_af, int32_t *out_map_bias) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!obs) return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    const double *cols[9] = {obs->prob_mapping, obs->prob_mismapping, obs->prob_alt, obs->prob_ref,
                             obs->prob_missed_allele, obs->prob_sample_alt, obs->prob_double_overlap,
                             obs->prob_single_overlap, obs->prob_hit_base};
    return posterior_pipeline<double>(ctx, n_sites, model, cols, 9, obs->cat, obs_off, out_event_ln, out_marginal,
                                      out_map_af, out_map_bias);
}

int vlr_posterior_batch_f32(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                            const vlr_obs_columns_f32 *obs, const int64_t *obs_off, double *out_event_ln,
                            double *out_marginal, double *out_map_af, int32_t *out_map_bias) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!obs) return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    const float *cols[7] = {obs->prob_mapping, obs->prob_alt, obs->prob_ref, obs->prob_missed_allele,
                            obs->prob_sample_alt, obs->prob_double_overlap, obs->prob_hit_base};
    return posterior_pipeline<float>(ctx, n_sites, model, cols, 7, obs->cat, obs_off, out_event_ln, out_marginal,
                                     out_map_af, out_map_bias);
}

}  // extern "C"

// ---------------------------------------------------------------- Likelihood::compute work items
// One warp per (site, sample, AF[, secondary AF], bias configuration) item: the literal log-space
// formulas of likelihood.rs (this entry point is the per-call seam; the batched table path above
// is the fast one).
namespace vlr {
__global__ void likelihood_items_kernel(vlr_obs_columns obs, const int64_t *obs_off, int32_t n_samples,
                                        const vlr_biases *biases, const vlr_lh_item *items,
                                        int64_t n_items, double *out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = w; it < n_items; it += nw) {
        const vlr_lh_item item = items[it];
        const int64_t o0 = obs_off[(int64_t)item.site * n_samples + item.sample];
        const int64_t o1 = obs_off[(int64_t)item.site * n_samples + item.sample + 1];
        const vlr_biases b = biases[item.bias];
        double acc = 0.0;
        for (int64_t row = o0 + lane; row < o1; row += 32)
            acc += ln_T_exact(obs, row, b, item.other_rate, item.af, item.af_secondary,
                              item.contaminated != 0, item.purity);
        acc = warp_sum(acc);
        if (lane == 0) out[it] = acc;
    }
}
}  // namespace vlr

extern "C" int vlr_likelihood_batch(vlr_ctx *ctx, int64_t n_sites, int32_t n_samples,
                                    const vlr_obs_columns *obs, const int64_t *obs_off,
                                    const vlr_biases *biases, int32_t n_biases,
                                    const vlr_lh_item *items, int64_t n_items, double *out_ln_lh) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!obs || !obs_off || !biases || !items || !out_ln_lh || n_sites <= 0 || n_samples <= 0 ||
        n_biases <= 0 || n_items < 0)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_items == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    for (int64_t k = 0; k < n_items; ++k) {
        const vlr_lh_item &it = items[k];
        if (it.site < 0 || it.site >= n_sites || it.sample < 0 || it.sample >= n_samples ||
            it.bias < 0 || it.bias >= n_biases)
            return set_err(ctx, VLR_ERR_INVALID, "item %lld out of range", (long long)k);
        if (it.contaminated && !(it.purity > 0.0 && it.purity <= 1.0))
            return set_err(ctx, VLR_ERR_INVALID, "item %lld: purity must be in (0,1]", (long long)k);
    }
    const int64_t n_off = n_sites * n_samples + 1;
    const int64_t r0 = obs_off[0], n_rows = obs_off[n_off - 1] - r0;
    enum { S_COL0 = 20, S_CAT = 29, S_OFF, S_B = 36, S_IT, S_O };
    const double *cols[9] = {obs->prob_mapping, obs->prob_mismapping, obs->prob_alt, obs->prob_ref,
                             obs->prob_missed_allele, obs->prob_sample_alt, obs->prob_double_overlap,
                             obs->prob_single_overlap, obs->prob_hit_base};
    double *d_cols[9];
    for (int c = 0; c < 9; ++c) {
        d_cols[c] = (double *)dev_scratch(ctx, S_COL0 + c, (size_t)(n_rows + 1) * 8);
        if (!d_cols[c]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        if (n_rows > 0)
            VLR_CUDA(ctx, cudaMemcpyAsync(d_cols[c], cols[c] + r0, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
    }
    uint16_t *d_cat = (uint16_t *)dev_scratch(ctx, S_CAT, (size_t)(n_rows + 1) * 2);
    int64_t *d_off = (int64_t *)dev_scratch(ctx, S_OFF, (size_t)n_off * 8);
    vlr_biases *d_b = (vlr_biases *)dev_scratch(ctx, S_B, (size_t)n_biases * sizeof(vlr_biases));
    vlr_lh_item *d_it = (vlr_lh_item *)dev_scratch(ctx, S_IT, (size_t)n_items * sizeof(vlr_lh_item));
    double *d_o = (double *)dev_scratch(ctx, S_O, (size_t)n_items * 8);
    if (!d_cat || !d_off || !d_b || !d_it || !d_o) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    if (n_rows > 0)
        VLR_CUDA(ctx, cudaMemcpyAsync(d_cat, obs->cat + r0, (size_t)n_rows * 2, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_off, obs_off, (size_t)n_off * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_b, biases, (size_t)n_biases * sizeof(vlr_biases), cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_it, items, (size_t)n_items * sizeof(vlr_lh_item), cudaMemcpyHostToDevice, st));
    vlr_obs_columns dc;
    dc.prob_mapping = d_cols[0] - r0; dc.prob_mismapping = d_cols[1] - r0; dc.prob_alt = d_cols[2] - r0;
    dc.prob_ref = d_cols[3] - r0; dc.prob_missed_allele = d_cols[4] - r0; dc.prob_sample_alt = d_cols[5] - r0;
    dc.prob_double_overlap = d_cols[6] - r0; dc.prob_single_overlap = d_cols[7] - r0;
    dc.prob_hit_base = d_cols[8] - r0; dc.cat = d_cat - r0;
    const int64_t want = (n_items * 32 + 255) / 256;
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 16);
    vlr::likelihood_items_kernel<<<grid, 256, 0, st>>>(dc, d_off, n_samples, d_b, d_it, n_items, d_o);
    VLR_LAUNCH_CHECK(ctx);
    VLR_CUDA(ctx, cudaMemcpyAsync(out_ln_lh, d_o, (size_t)n_items * 8, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    return VLR_OK;
}
