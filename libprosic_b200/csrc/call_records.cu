// call_records.cu -- values of the final call record, computed from the posterior outputs on the
// device.  Replaces the numeric part of the BCF writer of `varlociraptor call`
// (reference src/calling/variants/mod.rs:99-399 and calling.rs:573-602):
//   PROB_<EVENT>  = | PHREDProb(posterior) | as f32, posterior = event_ln - marginal (mod.rs:326-333);
//                   the "artifact" event is ln_sum_exp over the artifact events (calling.rs:583-598);
//                   missing (BCF float missing) when every sample's pileup is empty (mod.rs:309-321)
//   AF            = MAP allele frequency as f32 (mod.rs:185)
//   DP            = round(exp(ln_sum_exp(prob_mapping))) per sample (observation.rs:31-36, mod.rs:187)
#include <math.h>

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

extern "C" int vlr_call_records(vlr_ctx *ctx, int64_t n_sites, int32_t n_events, int32_t n_samples,
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
