// call_records.cu -- values of the final call record, computed from the posterior outputs on the
// device.  Replaces the numeric part of the BCF writer of `varlociraptor call`
// (reference src/calling/variants/mod.rs:99-399 and calling.rs:573-602):
//   PROB_<EVENT>  = | PHREDProb(posterior) | as f32, posterior = event_ln - marginal (mod.rs:326-333);
//                   the "artifact" event is ln_sum_exp over the artifact events (calling.rs:583-598);
//                   missing (BCF float missing) when every sample's pileup is empty (mod.rs:309-321)
//   AF            = MAP allele frequency as f32 (mod.rs:185)
//   DP            = round(exp(ln_sum_exp(prob_mapping))) per sample (observation.rs:31-36, mod.rs:187)
#include <math.h>
#include <string.h>

#include <algorithm>
#include <utility>

#include "common.cuh"

namespace vlr {

constexpr uint32_t kBcfFloatMissing = 0x7F800001u;  // htslib bcf_float_missing

struct CallArgs {
    int64_t n_sites;
    int32_t n_events, n_samples, n_named;
    const int32_t *named_of_event;  // event -> output column, -1 for artifact events (device)
    const double *event_ln, *marginal, *map_af, *prob_mapping;
    const int64_t *obs_off;
    float *out_prob, *out_af;
    int32_t *out_dp;
};

__global__ void call_records_kernel(CallArgs g) {
    const double kPhred = -4.342944819032518;  // -10 / ln 10 (PHREDProb::from(LogProb))
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < g.n_sites;
         s += (int64_t)gridDim.x * blockDim.x) {
        // ---- DP per sample, and whether the site has any observation at all
        bool any_obs = false;
        for (int m = 0; m < g.n_samples; ++m) {
            const int64_t o0 = g.obs_off[s * g.n_samples + m], o1 = g.obs_off[s * g.n_samples + m + 1];
            any_obs |= o1 > o0;
            // LogProb::ln_sum_exp: pmax + ln_1p(sum of the others), then exp and round
            double mx = -INFINITY;
            int64_t imax = o0;
            for (int64_t r = o0; r < o1; ++r) {
                const double p = g.prob_mapping[r];
                if (p > mx) { mx = p; imax = r; }
            }
            double dp = 0.0;
            if (mx != -INFINITY) {
                double acc = 0.0;
                for (int64_t r = o0; r < o1; ++r) {
                    const double p = g.prob_mapping[r];
                    if (r != imax && p != -INFINITY) acc += exp(p - mx);
                }
                dp = exp(mx + log1p(acc));
            }
            g.out_dp[s * g.n_samples + m] = (int32_t)(uint32_t)round(dp);
            g.out_af[s * g.n_samples + m] = (float)g.map_af[s * g.n_samples + m];
        }
        // ---- event probabilities
        const double marg = g.marginal[s];
        float *out = g.out_prob + s * (int64_t)(g.n_named + 1);
        if (!any_obs) {
            for (int k = 0; k <= g.n_named; ++k) out[k] = __uint_as_float(kBcfFloatMissing);
            continue;
        }
        double amax = -INFINITY;
        int n_art = 0;
        for (int e = 0; e < g.n_events; ++e) {
            const double post = g.event_ln[s * g.n_events + e] - marg;  // ModelInstance::posterior
            const int col = g.named_of_event[e];
            if (col >= 0) out[col] = (float)fabs(post * kPhred);
            else { n_art++; if (post > amax || isnan(post)) amax = isnan(post) ? post : fmax(amax, post); }
        }
        double part = -INFINITY;
        if (n_art > 0 && !(amax == -INFINITY)) {
            // ln_sum_exp over the artifact posteriors: first maximum, then the others in order
            int imax = -1;
            double acc = 0.0;
            for (int e = 0; e < g.n_events; ++e)
                if (g.named_of_event[e] < 0 && imax < 0 && g.event_ln[s * g.n_events + e] - marg == amax) imax = e;
            for (int e = 0; e < g.n_events; ++e) {
                if (g.named_of_event[e] >= 0 || e == imax) continue;
                const double post = g.event_ln[s * g.n_events + e] - marg;
                if (post != -INFINITY) acc += exp(post - amax);
            }
            part = amax + log1p(acc);
        }
        out[g.n_named] = (float)fabs(part * kPhred);
    }
}

}  // namespace vlr

using namespace vlr;

static int vlr_call_records_impl(vlr_ctx *ctx, int64_t n_sites, int32_t n_events, int32_t n_samples,
                                const uint8_t *event_is_artifact, const double *event_ln,
                                const double *marginal, const double *map_af,
                                const double *prob_mapping, const int64_t *obs_off, float *out_prob,
                                float *out_af, int32_t *out_dp) {
    if (!ctx) return VLR_ERR_INVALID;
    if (n_sites < 0 || n_events < 1 || n_events > 4096 || n_samples < 1 || !event_is_artifact || !event_ln ||
        !marginal || !map_af || !prob_mapping || !obs_off || !out_prob || !out_af || !out_dp)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_sites == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> named((size_t)n_events);
    int n_named = 0;
    for (int e = 0; e < n_events; ++e) named[e] = event_is_artifact[e] ? -1 : n_named++;
    const int64_t n_off = n_sites * n_samples + 1;
    const int64_t r0 = obs_off[0], n_rows = obs_off[n_off - 1] - r0;
    if (n_rows < 0) return set_err(ctx, VLR_ERR_INVALID, "obs_off not monotone");
    enum { S_NAMED = 80, S_EV, S_MARG, S_AF, S_PM, S_OFF, S_OP, S_OAF, S_ODP };
    int32_t *d_named = (int32_t *)dev_scratch(ctx, S_NAMED, (size_t)n_events * 4);
    double *d_ev = (double *)dev_scratch(ctx, S_EV, (size_t)n_sites * n_events * 8);
    double *d_marg = (double *)dev_scratch(ctx, S_MARG, (size_t)n_sites * 8);
    double *d_af = (double *)dev_scratch(ctx, S_AF, (size_t)n_sites * n_samples * 8);
    double *d_pm = (double *)dev_scratch(ctx, S_PM, (size_t)(n_rows + 1) * 8);
    int64_t *d_off = (int64_t *)dev_scratch(ctx, S_OFF, (size_t)n_off * 8);
    float *d_op = (float *)dev_scratch(ctx, S_OP, (size_t)n_sites * (n_named + 1) * 4);
    float *d_oaf = (float *)dev_scratch(ctx, S_OAF, (size_t)n_sites * n_samples * 4);
    int32_t *d_odp = (int32_t *)dev_scratch(ctx, S_ODP, (size_t)n_sites * n_samples * 4);
    if (!d_named || !d_ev || !d_marg || !d_af || !d_pm || !d_off || !d_op || !d_oaf || !d_odp)
        return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    auto H2D = [&](void *d, const void *h, size_t b) { return cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, st); };
    VLR_CUDA(ctx, H2D(d_named, named.data(), (size_t)n_events * 4));
    VLR_CUDA(ctx, H2D(d_ev, event_ln, (size_t)n_sites * n_events * 8));
    VLR_CUDA(ctx, H2D(d_marg, marginal, (size_t)n_sites * 8));
    VLR_CUDA(ctx, H2D(d_af, map_af, (size_t)n_sites * n_samples * 8));
    if (n_rows > 0) VLR_CUDA(ctx, H2D(d_pm, prob_mapping + r0, (size_t)n_rows * 8));
    VLR_CUDA(ctx, H2D(d_off, obs_off, (size_t)n_off * 8));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));  // `named` is a stack-owned staging vector
    CallArgs a;
    a.n_sites = n_sites; a.n_events = n_events; a.n_samples = n_samples; a.n_named = n_named;
    a.named_of_event = d_named; a.event_ln = d_ev; a.marginal = d_marg; a.map_af = d_af;
    a.prob_mapping = d_pm - r0; a.obs_off = d_off; a.out_prob = d_op; a.out_af = d_oaf; a.out_dp = d_odp;
    const int grid = (int)std::min<int64_t>((n_sites + 127) / 128, (int64_t)ctx->sm_count * 16);
    call_records_kernel<<<grid, 128, 0, st>>>(a);
    VLR_LAUNCH_CHECK(ctx);
    VLR_CUDA(ctx, cudaMemcpyAsync(out_prob, d_op, (size_t)n_sites * (n_named + 1) * 4, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(out_af, d_oaf, (size_t)n_sites * n_samples * 4, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(out_dp, d_odp, (size_t)n_sites * n_samples * 4, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    return VLR_OK;
}

extern "C" int vlr_call_records(vlr_ctx *ctx, int64_t n_sites, int32_t n_events, int32_t n_samples,
                                const uint8_t *event_is_artifact, const double *event_ln,
                                const double *marginal, const double *map_af,
                                const double *prob_mapping, const int64_t *obs_off, float *out_prob,
                                float *out_af, int32_t *out_dp) {
    return vlr::drain_on_error(ctx, vlr_call_records_impl(ctx, n_sites, n_events, n_samples, event_is_artifact, event_ln, marginal, map_af, prob_mapping, obs_off, out_prob, out_af, out_dp));
}


// ---------------------------------------------------------------- OBS / SOBS strings, bias flags
// The per-sample OBS and SOBS FORMAT strings of the call record (src/calling/variants/mod.rs:186-260)
// are "generalized CIGAR" strings over a 7-character code per observation (utils/mod.rs:101-146):
//   Kass-Raftery letter of the alt-vs-ref Bayes factor (N E B P S V; lower case if
//   prob_mapping < ln 0.95), p|s (paired), strand * - +, read orientation > < * !, read position ^ *,
//   softclip $ ., indel operations * # .
// The device derives the packed code of every observation (one exp per row, fused with nothing else
// worth a kernel: it is one pass over three columns); counting and formatting is host work.
namespace vlr {

// packed code: letter (0 N, 1 E, 2 B, 3 P, 4 S, 5 V) | lower << 3 | paired << 4 | strand (0 *, 1 -, 2 +,
// 3 invalid) << 5 | orientation (0 >, 1 <, 2 *, 3 !) << 7 | read position (0 ^, 1 *) << 9 |
// softclip << 10 | indel (0 *, 1 #, 2 .) << 11
__host__ __device__ static inline uint16_t obs_code_of(double prob_alt, double prob_ref, double prob_mapping,
                                                        unsigned cat, double bf) {
    // BayesFactor::evidence_kass_raftery (bio 0.37) + bayes_factor_to_letter
    int letter;
    if (bf <= 1.0) {
        const double d = fabs(bf - 1.0);  // approx::relative_eq!(bf, 1.0)
        letter = (bf == 1.0 || d <= 2.220446049250313e-16 || d <= fmax(fabs(bf), 1.0) * 2.220446049250313e-16) ? 1 : 0;
    } else if (bf <= 3.0) letter = 2;
    else if (bf <= 20.0) letter = 3;
    else if (bf <= 150.0) letter = 4;
    else letter = 5;
    (void)prob_alt; (void)prob_ref;
    const unsigned lower = prob_mapping < -0.05129329438755058 ? 1u : 0u;  // ln 0.95
    const unsigned st = VLR_CAT_STRAND(cat), orr = VLR_CAT_ORIENT(cat);
    const unsigned strand = st == 2 ? 0u : (st == 1 ? 1u : (st == 0 ? 2u : 3u));
    const unsigned orient = orr == 0 ? 0u : (orr == 1 ? 1u : (orr == 8 ? 2u : 3u));
    return (uint16_t)(letter | (lower << 3) | (VLR_CAT_PAIRED(cat) << 4) | (strand << 5) | (orient << 7) |
                      (VLR_CAT_READPOS(cat) << 9) | (VLR_CAT_SOFTCLIP(cat) << 10) | (VLR_CAT_INDEL(cat) << 11));
}

__global__ void obs_codes_kernel(int64_t n, const double *prob_alt, const double *prob_ref, const double *prob_mapping,
                                 const uint16_t *cat, uint16_t *out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = obs_code_of(prob_alt[k], prob_ref[k], prob_mapping[k], cat[k], exp(prob_alt[k] - prob_ref[k]));
}

static int code_text(uint16_t c, int simple, char *o) {
    const char L[6] = {'N', 'E', 'B', 'P', 'S', 'V'};
    char l = L[(c & 7) < 6 ? (c & 7) : 0];
    if ((c >> 3) & 1) l = (char)(l + 32);
    o[0] = l;
    if (simple) return 1;
    o[1] = ((c >> 4) & 1) ? 'p' : 's';
    o[2] = "*-+?"[(c >> 5) & 3];
    o[3] = "><*!"[(c >> 7) & 3];
    o[4] = ((c >> 9) & 1) ? '*' : '^';
    o[5] = ((c >> 10) & 1) ? '$' : '.';
    o[6] = "*#.?"[(c >> 11) & 3];
    return 7;
}

}  // namespace vlr

extern "C" int vlr_obs_codes(vlr_ctx *ctx, int64_t n_rows, const double *prob_alt, const double *prob_ref,
                             const double *prob_mapping, const uint16_t *cat, uint16_t *out_code) {
    if (!ctx) return VLR_ERR_INVALID;
    if (n_rows < 0 || (n_rows > 0 && (!prob_alt || !prob_ref || !prob_mapping || !cat || !out_code)))
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_rows == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    enum { S_A = 94, S_R, S_M, S_C, S_O };
    double *d_a = (double *)dev_scratch(ctx, S_A, (size_t)n_rows * 8), *d_r = (double *)dev_scratch(ctx, S_R, (size_t)n_rows * 8);
    double *d_m = (double *)dev_scratch(ctx, S_M, (size_t)n_rows * 8);
    uint16_t *d_c = (uint16_t *)dev_scratch(ctx, S_C, (size_t)n_rows * 2), *d_o = (uint16_t *)dev_scratch(ctx, S_O, (size_t)n_rows * 2);
    if (!d_a || !d_r || !d_m || !d_c || !d_o) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    VLR_CUDA(ctx, cudaMemcpyAsync(d_a, prob_alt, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_r, prob_ref, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_m, prob_mapping, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_c, cat, (size_t)n_rows * 2, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<int64_t>((n_rows + 255) / 256, (int64_t)ctx->sm_count * 8);
    obs_codes_kernel<<<grid, 256, 0, st>>>(n_rows, d_a, d_r, d_m, d_c, d_o);
    VLR_LAUNCH_CHECK(ctx);
    VLR_CUDA(ctx, cudaMemcpyAsync(out_code, d_o, (size_t)n_rows * 2, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    return VLR_OK;
}

// generalized_cigar(items, keep_order = false, aux_sort): "<count><code>" per distinct code, most
// common first, codes starting with E after the others and with N last (stable).  The reference's
// order among equal counts is a HashMap's; this library breaks ties by the packed code.
extern "C" int64_t vlr_obs_cigar(const uint16_t *codes, int64_t n, int simple, char *out, int64_t capacity) {
    if (n < 0 || (n > 0 && !codes) || !out || capacity < 2) return VLR_ERR_INVALID;
    std::vector<std::pair<uint16_t, int64_t>> cnt;  // (code or letter class, count)
    {
        std::vector<int64_t> hist(1 << 13, 0);
        for (int64_t k = 0; k < n; ++k) hist[(simple ? (codes[k] & 15) : codes[k]) & 0x1fff]++;
        for (int c = 0; c < (1 << 13); ++c) if (hist[c]) cnt.emplace_back((uint16_t)c, hist[c]);
    }
    auto klass = [](uint16_t c) { return (c & 7) == 0 && !((c >> 3) & 1) ? 2 : (((c & 7) == 1 && !((c >> 3) & 1)) ? 1 : 0); };
    std::stable_sort(cnt.begin(), cnt.end(), [](const std::pair<uint16_t, int64_t> &a, const std::pair<uint16_t, int64_t> &b) {
        return a.second != b.second ? a.second > b.second : a.first < b.first;
    });
    std::stable_sort(cnt.begin(), cnt.end(), [&](const std::pair<uint16_t, int64_t> &a, const std::pair<uint16_t, int64_t> &b) {
        return klass(a.first) < klass(b.first);
    });
    int64_t len = 0;
    if (cnt.empty()) {  // an empty observation list is written as "."
        out[0] = '.'; out[1] = 0;
        return 1;
    }
    for (const auto &e : cnt) {
        char buf[40];
        int m = snprintf(buf, sizeof(buf), "%lld", (long long)e.second);
        m += vlr::code_text(e.first, simple, buf + m);
        if (len + m + 1 > capacity) return VLR_ERR_INVALID;
        memcpy(out + len, buf, (size_t)m);
        len += m;
    }
    out[len] = 0;
    return len;
}

// SB ROB RPB SCB DIB flag characters of a bias configuration (calling/variants/mod.rs:147-183)
extern "C" int vlr_bias_chars(const vlr_biases *b, char out[5]) {
    if (!b || !out) return VLR_ERR_INVALID;
    out[0] = b->strand == 1 ? '+' : (b->strand == 2 ? '-' : '.');
    out[1] = b->orientation == 1 ? '>' : (b->orientation == 2 ? '<' : '.');
    out[2] = b->read_position ? '^' : '.';
    out[3] = b->softclip ? '$' : '.';
    out[4] = b->divindel ? '#' : '.';
    return VLR_OK;
}
