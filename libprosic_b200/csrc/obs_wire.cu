// obs_wire.cu -- the observation wire format ON THE DEVICE: the records `varlociraptor preprocess`
// writes into the BCF (bincode bytes as u16 words, one INFO vector per field and sample; reference
// src/calling/variants/preprocessing/mod.rs:452-598, MiniLogProb src/utils/mod.rs:381-407) are uploaded
// as they are and decoded by a kernel into the struct-of-arrays observation columns the posterior
// kernels read -- `call` no longer decodes on the host.  The host codec of obs_codec.cu stays as the
// checker (tests compare the two bit for bit on the records written by the reference itself).
//
// A MiniLogProb element is 6 bytes (u32 variant 0 + f16) or 8 bytes (u32 variant 1 + f32), so element k
// of a vector can only be found by walking elements 0..k-1: one thread per (record, field) walks its
// vector (a pileup has tens to hundreds of observations; hundreds of thousands of records per batch
// keep the device busy; a warp takes the same field of 32 consecutive records), one more thread per
// record assembles the packed categoricals from the four
// enum vectors (fixed 4-byte elements) and the two BitVec<u8>.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace vlr {

struct WireDev {
    const uint16_t *w[VLR_OBS_N_FIELDS];   // rebased: w[f] + off[f][r] is the first word of record r
    const int64_t *off[VLR_OBS_N_FIELDS];  // rebased: off[f][r] for chunk-local r
};
struct ColsDev {
    double *c[9];  // vlr_obs_columns order
    uint16_t *cat;
};

__device__ __forceinline__ double ln_one_minus_exp_w(double p) {  // observation.rs:218-226 (LogProb::ln_one_minus_exp)
    return p < -0.693 ? log1p(-exp(p)) : log(-expm1(p));
}
__device__ __forceinline__ uint32_t u32_at(const uint16_t *p, int64_t w) { return (uint32_t)p[w] | ((uint32_t)p[w + 1] << 16); }
__device__ __forceinline__ uint64_t u64_at(const uint16_t *p, int64_t w) {
    return (uint64_t)u32_at(p, w) | ((uint64_t)u32_at(p, w + 2) << 32);
}
__device__ __forceinline__ uint32_t byte_at(const uint16_t *p, int64_t b) { return (p[b >> 1] >> (8 * (b & 1))) & 0xffu; }
__device__ __forceinline__ uint64_t u64_at_byte(const uint16_t *p, int64_t b) {
    uint64_t v = 0;
    for (int k = 0; k < 8; ++k) v |= (uint64_t)byte_at(p, b + k) << (8 * k);
    return v;
}

// err[0]: number of malformed (record, field) vectors; err[1]: first offending record * 16 + field + 1
__device__ __forceinline__ void wire_fail(int32_t *err, int64_t r, int f) {
    if (atomicAdd(&err[0], 1) == 0) err[1] = (int32_t)(r * 16 + f + 1);
}

__global__ void __launch_bounds__(256) obs_wire_decode_kernel(WireDev W, int64_t n_records, const int64_t *obs_off,
                                                              int64_t row_base, ColsDev out, int32_t *err) {
    // a CTA of 256 threads takes 32 records; warp w of the CTA handles task w of all of them, so the
    // lanes of a warp run the same code (the transcendentals of the two derived columns fill whole
    // warps instead of 2 lanes in 8)
    const int64_t r = blockIdx.x * (int64_t)32 + (threadIdx.x & 31);
    const int task = (int)(threadIdx.x >> 5);
    if (r >= n_records) return;
    const int64_t row0 = obs_off[r] - row_base, n = obs_off[r + 1] - obs_off[r];
    if (task < 7) {
        // ---- Vec<MiniLogProb>: u64 length, then (u32 variant, f16 | f32) per element
        const int col_of_field[7] = {0, 3, 2, 4, 5, 6, 8};  // field order -> vlr_obs_columns order
        const uint16_t *p = W.w[task] + W.off[task][r];
        const int64_t nw = W.off[task][r + 1] - W.off[task][r];
        if (nw < 4 || (int64_t)u64_at(p, 0) != n) { wire_fail(err, r, task); return; }
        double *dst = out.c[col_of_field[task]] + row0;
        double *der = task == 0 ? out.c[1] + row0 : (task == 5 ? out.c[7] + row0 : nullptr);
        int64_t pos = 4;
        for (int64_t k = 0; k < n; ++k) {
            if (pos + 3 > nw) { wire_fail(err, r, task); return; }
            const uint32_t var = u32_at(p, pos);
            double v;
            if (var == 0) {
                v = (double)__half2float(__ushort_as_half(p[pos + 2]));
                pos += 3;
            } else if (var == 1 && pos + 4 <= nw) {
                v = (double)__uint_as_float(u32_at(p, pos + 2));
                pos += 4;
            } else {
                wire_fail(err, r, task);
                return;
            }
            dst[k] = v;
            if (der) der[k] = ln_one_minus_exp_w(v);  // prob_mismapping / prob_single_overlap
        }
        if (pos != nw) wire_fail(err, r, task);
        return;
    }
    // ---- categoricals: four enum vectors (u64 length, u32 variant index each) and two BitVec<u8>
    const int enum_field[4] = {VLR_OBS_STRAND, VLR_OBS_READ_ORIENTATION, VLR_OBS_READ_POSITION, VLR_OBS_INDEL_OPERATIONS};
    const int enum_shift[4] = {0, 2, 6, 8};
    const uint32_t enum_n[4] = {4, 9, 2, 3};
    const uint16_t *ep[4];
    bool have[4];
    for (int e = 0; e < 4; ++e) {
        const int f = enum_field[e];
        have[e] = W.w[f] != nullptr && W.off[f][r + 1] > W.off[f][r];
        ep[e] = nullptr;
        if (!have[e]) {
            if (f != VLR_OBS_INDEL_OPERATIONS) { wire_fail(err, r, f); return; }
            continue;  // format 7: IndelOperations::None
        }
        ep[e] = W.w[f] + W.off[f][r];
        const int64_t nw = W.off[f][r + 1] - W.off[f][r];
        if (nw != 4 + 2 * n || (int64_t)u64_at(ep[e], 0) != n) { wire_fail(err, r, f); return; }
    }
    const int bit_field[2] = {VLR_OBS_SOFTCLIPPED, VLR_OBS_PAIRED};
    const int bit_shift[2] = {7, 10};
    const uint16_t *bp[2];
    for (int e = 0; e < 2; ++e) {
        const int f = bit_field[e];
        if (!W.w[f]) { wire_fail(err, r, f); return; }
        bp[e] = W.w[f] + W.off[f][r];
        const int64_t nw = W.off[f][r + 1] - W.off[f][r];
        if (nw < 5) { wire_fail(err, r, f); return; }
        const uint32_t tag = byte_at(bp[e], 0);
        bool ok;
        if (tag == 0) {
            ok = n == 0 && u64_at_byte(bp[e], 1) == 0 && nw == 5 && byte_at(bp[e], 9) == 0;
        } else if (tag == 1) {
            const int64_t nb = (int64_t)u64_at_byte(bp[e], 1), bytes = 17 + nb;
            ok = nb == (n + 7) / 8 && nw == (bytes + 1) / 2 && (int64_t)u64_at_byte(bp[e], 9 + nb) == n &&
                 ((bytes & 1) == 0 || byte_at(bp[e], bytes) == 0);
        } else {
            ok = false;
        }
        if (!ok) { wire_fail(err, r, f); return; }
    }
    for (int64_t k = 0; k < n; ++k) {
        uint32_t c = 0;
        for (int e = 0; e < 4; ++e) {
            if (!have[e]) { c |= 2u << 8; continue; }
            const uint32_t v = u32_at(ep[e], 4 + 2 * k);
            if (v >= enum_n[e]) { wire_fail(err, r, enum_field[e]); return; }
            c |= v << enum_shift[e];
        }
        for (int e = 0; e < 2; ++e)
            if ((byte_at(bp[e], 9 + (k >> 3)) >> (k & 7)) & 1) c |= 1u << bit_shift[e];
        out.cat[row0 + k] = (uint16_t)c;
    }
}

// Observations of record r from the STRAND vector (fixed 4-byte elements): words = 4 + 2 n.
static int wire_row_offsets(vlr_ctx *ctx, int64_t n_records, const vlr_obs_wire *wire, std::vector<int64_t> &obs_off,
                            int64_t &max_obs) {
    for (int f = 0; f < VLR_OBS_N_FIELDS; ++f) {
        if (f == VLR_OBS_INDEL_OPERATIONS && !wire->words[f]) continue;  // format 7
        if (!wire->words[f] || !wire->word_off[f]) return set_err(ctx, VLR_ERR_INVALID, "wire field %d missing", f);
        for (int64_t r = 0; r < n_records; ++r)
            if (wire->word_off[f][r + 1] < wire->word_off[f][r])
                return set_err(ctx, VLR_ERR_INVALID, "word_off of field %d not monotone", f);
    }
    obs_off.resize((size_t)n_records + 1);
    obs_off[0] = 0;
    max_obs = 0;
    const int64_t *so = wire->word_off[VLR_OBS_STRAND];
    for (int64_t r = 0; r < n_records; ++r) {
        const int64_t nw = so[r + 1] - so[r];
        if (nw < 4 || ((nw - 4) & 1)) return set_err(ctx, VLR_ERR_INVALID, "record %lld: malformed STRAND vector", (long long)r);
        const int64_t n = (nw - 4) / 2;
        obs_off[r + 1] = obs_off[r] + n;
        max_obs = std::max(max_obs, n);
    }
    return VLR_OK;
}

// Streams the records [rec0, rec1) of every field to the device buffers of set b (copy stream) and
// returns the rebased device view.  obs_off values stay absolute.
struct WireSet {
    uint16_t *words[VLR_OBS_N_FIELDS];
    int64_t *off[VLR_OBS_N_FIELDS];
    int64_t *obs_off;
};

static int wire_upload(vlr_ctx *ctx, cudaStream_t cp, const vlr_obs_wire *wire, const int64_t *h_obs_off, int64_t rec0,
                       int64_t rec1, const WireSet &S, WireDev &W) {
    for (int f = 0; f < VLR_OBS_N_FIELDS; ++f) {
        W.w[f] = nullptr;
        W.off[f] = nullptr;
        if (!wire->words[f]) continue;
        const int64_t w0 = wire->word_off[f][rec0], nw = wire->word_off[f][rec1] - w0;
        if (nw > 0)
            VLR_CUDA(ctx, cudaMemcpyAsync(S.words[f], wire->words[f] + w0, (size_t)nw * 2, cudaMemcpyHostToDevice, cp));
        VLR_CUDA(ctx, cudaMemcpyAsync(S.off[f], wire->word_off[f] + rec0, (size_t)(rec1 - rec0 + 1) * 8,
                                      cudaMemcpyHostToDevice, cp));
        W.w[f] = S.words[f] - w0;   // offsets stay absolute: rebase the pointer
        W.off[f] = S.off[f];
    }
    VLR_CUDA(ctx, cudaMemcpyAsync(S.obs_off, h_obs_off + rec0, (size_t)(rec1 - rec0 + 1) * 8, cudaMemcpyHostToDevice, cp));
    return VLR_OK;
}

// chunk boundaries (in units of `unit` records) with about `target_bytes` of wire data each
static std::vector<int64_t> wire_chunks(const vlr_obs_wire *wire, int64_t n_units, int64_t unit, double target_bytes) {
    auto words_before = [&](int64_t u) {
        int64_t s = 0;
        for (int f = 0; f < VLR_OBS_N_FIELDS; ++f)
            if (wire->words[f]) s += wire->word_off[f][u * unit] - wire->word_off[f][0];
        return s;
    };
    const double total = 2.0 * (double)words_before(n_units);
    int n_chunks = (int)std::min<double>(64.0, std::max<double>(1.0, ceil(total / target_bytes)));
    if ((int64_t)n_chunks > n_units) n_chunks = (int)std::max<int64_t>(1, n_units);
    std::vector<int64_t> cs((size_t)n_chunks + 1);
    cs[0] = 0;
    for (int c = 1; c < n_chunks; ++c) {
        const double want = total * c / n_chunks;
        int64_t lo = cs[c - 1] + 1, hi = n_units - (n_chunks - c);
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (2.0 * (double)words_before(mid) < want) lo = mid + 1; else hi = mid;
        }
        cs[c] = lo;
    }
    cs[n_chunks] = n_units;
    return cs;
}

enum { W_BASE = 200, W_PER = 48 };  // per set: 0..12 words, 13..25 offsets, 26 obs_off, 27..35 columns, 36 cat, 37..40 outputs

static int wire_alloc(vlr_ctx *ctx, int b, const vlr_obs_wire *wire, const std::vector<int64_t> &cs, int64_t unit,
                      const int64_t *h_obs_off, WireSet &S, ColsDev &C, int64_t &max_rows, int64_t &max_recs) {
    const int base = W_BASE + b * W_PER;
    max_rows = 0;
    max_recs = 0;
    int64_t max_words[VLR_OBS_N_FIELDS] = {0};
    for (size_t c = 0; c + 1 < cs.size(); ++c) {
        const int64_t r0 = cs[c] * unit, r1 = cs[c + 1] * unit;
        max_rows = std::max(max_rows, h_obs_off[r1] - h_obs_off[r0]);
        max_recs = std::max(max_recs, r1 - r0);
        for (int f = 0; f < VLR_OBS_N_FIELDS; ++f)
            if (wire->words[f]) max_words[f] = std::max(max_words[f], wire->word_off[f][r1] - wire->word_off[f][r0]);
    }
    for (int f = 0; f < VLR_OBS_N_FIELDS; ++f) {
        S.words[f] = (uint16_t *)dev_scratch(ctx, base + f, (size_t)(max_words[f] + 8) * 2);
        S.off[f] = (int64_t *)dev_scratch(ctx, base + 13 + f, (size_t)(max_recs + 1) * 8);
        if (!S.words[f] || !S.off[f]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    }
    S.obs_off = (int64_t *)dev_scratch(ctx, base + 26, (size_t)(max_recs + 1) * 8);
    for (int c = 0; c < 9; ++c) C.c[c] = (double *)dev_scratch(ctx, base + 27 + c, (size_t)(max_rows + 1) * 8);
    C.cat = (uint16_t *)dev_scratch(ctx, base + 36, (size_t)(max_rows + 1) * 2);
    if (!S.obs_off || !C.cat) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    for (int c = 0; c < 9; ++c)
        if (!C.c[c]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    return VLR_OK;
}

static int wire_launch_decode(vlr_ctx *ctx, cudaStream_t st, const WireDev &W, int64_t n_records, const int64_t *d_obs_off,
                              int64_t row_base, const ColsDev &C, int32_t *d_err) {
    if (n_records <= 0) return VLR_OK;
    obs_wire_decode_kernel<<<(unsigned)((n_records + 31) / 32), 256, 0, st>>>(W, n_records, d_obs_off, row_base, C, d_err);
    VLR_LAUNCH_CHECK(ctx);
    return VLR_OK;
}

}  // namespace vlr

using namespace vlr;

extern "C" {

int vlr_obs_decode_batch(vlr_ctx *ctx, int64_t n_records, const vlr_obs_wire *wire, int64_t capacity,
                         int64_t *out_obs_off, double *const *out_cols, uint16_t *out_cat) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!wire || !out_obs_off || n_records < 0) return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    out_obs_off[0] = 0;
    if (n_records == 0) return VLR_OK;
    std::vector<int64_t> obs_off;
    int64_t max_obs = 0;
    int rc = wire_row_offsets(ctx, n_records, wire, obs_off, max_obs);
    if (rc != VLR_OK) return rc;
    memcpy(out_obs_off, obs_off.data(), (size_t)(n_records + 1) * 8);
    if (!out_cols && !out_cat) return VLR_OK;  // size query
    const int64_t n_rows = obs_off[n_records];
    if (!out_cols || !out_cat || n_rows > capacity) return set_err(ctx, VLR_ERR_INVALID, "output capacity too small");
    for (int c = 0; c < 9; ++c)
        if (!out_cols[c]) return set_err(ctx, VLR_ERR_INVALID, "null output column");
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const std::vector<int64_t> cs = {0, n_records};
    WireSet S;
    ColsDev C;
    int64_t max_rows, max_recs;
    rc = wire_alloc(ctx, 0, wire, cs, 1, obs_off.data(), S, C, max_rows, max_recs);
    if (rc != VLR_OK) return rc;
    int32_t *d_err = (int32_t *)dev_scratch(ctx, W_BASE + 2 * W_PER, 64);
    if (!d_err) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    VLR_CUDA(ctx, cudaMemsetAsync(d_err, 0, 64, st));
    WireDev W;
    rc = wire_upload(ctx, st, wire, obs_off.data(), 0, n_records, S, W);
    if (rc != VLR_OK) { cudaStreamSynchronize(st); return rc; }
    rc = wire_launch_decode(ctx, st, W, n_records, S.obs_off, 0, C, d_err);
    if (rc != VLR_OK) { cudaStreamSynchronize(st); return rc; }
    int32_t h_err[2] = {0, 0};
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < 9 && e == cudaSuccess && n_rows > 0; ++c)
        e = cudaMemcpyAsync(out_cols[c], C.c[c], (size_t)n_rows * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && n_rows > 0) e = cudaMemcpyAsync(out_cat, C.cat, (size_t)n_rows * 2, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_err, d_err, 8, cudaMemcpyDeviceToHost, st);
    const cudaError_t es = cudaStreamSynchronize(st);  // drains the copies from the caller's buffers on every path
    if (e != cudaSuccess || es != cudaSuccess)
        return set_err(ctx, VLR_ERR_CUDA, "wire decode failed: %s", cudaGetErrorString(e != cudaSuccess ? e : es));
    if (h_err[0])
        return set_err(ctx, VLR_ERR_INVALID, "%d malformed observation vector(s); first: record %d, field %d", h_err[0],
                       (h_err[1] - 1) / 16, (h_err[1] - 1) % 16);
    return VLR_OK;
}

int vlr_posterior_batch_wire(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model, const vlr_obs_wire *wire,
                             double *out_event_ln, double *out_marginal, double *out_map_af, int32_t *out_map_bias) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!model || !wire || !out_event_ln || !out_marginal || !out_map_af || !out_map_bias || n_sites < 0)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_sites == 0) return VLR_OK;
    const int NS = model->n_samples, NE = model->n_events;
    if (NS < 1 || NS > 8) return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_samples must be 1..8");
    if (NE < 1) return set_err(ctx, VLR_ERR_INVALID, "no events");
    const int64_t n_records = n_sites * NS;
    std::vector<int64_t> obs_off;
    int64_t max_obs = 0;
    int rc = wire_row_offsets(ctx, n_records, wire, obs_off, max_obs);
    if (rc != VLR_OK) return rc;
    if (max_obs > 4096) return set_err(ctx, VLR_ERR_UNSUPPORTED, "pileup of %lld observations exceeds 4096", (long long)max_obs);
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    rc = ensure_copy_stream(ctx);
    if (rc != VLR_OK) return rc;
    cudaStream_t cp = ctx->copy_stream;
    const std::vector<int64_t> cs = wire_chunks(wire, n_sites, NS, 32.0 * 1048576.0);
    const int n_chunks = (int)cs.size() - 1;
    WireSet S[2];
    ColsDev C[2];
    double *d_ev[2], *d_marg[2], *d_af[2];
    int32_t *d_mb[2];
    for (int b = 0; b < 2; ++b) {
        int64_t max_rows, max_recs;
        rc = wire_alloc(ctx, b, wire, cs, NS, obs_off.data(), S[b], C[b], max_rows, max_recs);
        if (rc != VLR_OK) return rc;
        const int base = W_BASE + b * W_PER;
        const int64_t max_sites = max_recs / NS;
        d_ev[b] = (double *)dev_scratch(ctx, base + 37, (size_t)max_sites * NE * 8);
        d_marg[b] = (double *)dev_scratch(ctx, base + 38, (size_t)max_sites * 8);
        d_af[b] = (double *)dev_scratch(ctx, base + 39, (size_t)max_sites * NS * 8);
        d_mb[b] = (int32_t *)dev_scratch(ctx, base + 40, (size_t)max_sites * NS * 4);
        if (!d_ev[b] || !d_marg[b] || !d_af[b] || !d_mb[b]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    }
    int32_t *d_err = (int32_t *)dev_scratch(ctx, W_BASE + 2 * W_PER, 64);
    int32_t *h_status = (int32_t *)pin_scratch(ctx, 6, (size_t)n_chunks * 4 + 32);
    if (!d_err || !h_status) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    memset(h_status, 0, (size_t)n_chunks * 4 + 32);
    int32_t *h_err = h_status + n_chunks;
    // earlier work of this context may still read the buffers the copy stream is about to overwrite
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    rc = VLR_OK;
    cudaError_t e = cudaMemsetAsync(d_err, 0, 64, st);
    for (int c = 0; c <= n_chunks && rc == VLR_OK && e == cudaSuccess; ++c) {
        if (c < n_chunks) {  // H2D of chunk c
            const int b = c & 1;
            if (c >= 2) e = cudaStreamWaitEvent(cp, ctx->ev_done[b], 0);
            WireDev W;
            if (e == cudaSuccess) rc = wire_upload(ctx, cp, wire, obs_off.data(), cs[c] * NS, cs[c + 1] * NS, S[b], W);
            if (rc == VLR_OK && e == cudaSuccess) e = cudaEventRecord(ctx->ev_h2d[b], cp);
        }
        if (c >= 1 && rc == VLR_OK && e == cudaSuccess) {  // decode + posterior + D2H of chunk c-1
            const int b = (c - 1) & 1;
            const int64_t s0 = cs[c - 1], s1 = cs[c], r0 = s0 * NS, r1 = s1 * NS;
            const int64_t row_base = obs_off[r0];
            e = cudaStreamWaitEvent(st, ctx->ev_h2d[b], 0);
            if (e != cudaSuccess) break;
            WireDev W;
            for (int f = 0; f < VLR_OBS_N_FIELDS; ++f) {
                W.w[f] = wire->words[f] ? S[b].words[f] - wire->word_off[f][r0] : nullptr;
                W.off[f] = wire->words[f] ? S[b].off[f] : nullptr;
            }
            rc = wire_launch_decode(ctx, st, W, r1 - r0, S[b].obs_off, row_base, C[b], d_err);
            if (rc != VLR_OK) break;
            vlr_obs_columns dc;  // obs_off stays absolute: rebase the column pointers
            dc.prob_mapping = C[b].c[0] - row_base; dc.prob_mismapping = C[b].c[1] - row_base;
            dc.prob_alt = C[b].c[2] - row_base; dc.prob_ref = C[b].c[3] - row_base;
            dc.prob_missed_allele = C[b].c[4] - row_base; dc.prob_sample_alt = C[b].c[5] - row_base;
            dc.prob_double_overlap = C[b].c[6] - row_base; dc.prob_single_overlap = C[b].c[7] - row_base;
            dc.prob_hit_base = C[b].c[8] - row_base; dc.cat = C[b].cat - row_base;
            rc = vlr_posterior_batch_dev(ctx, s1 - s0, model, &dc, S[b].obs_off, (int32_t)max_obs, d_ev[b], d_marg[b],
                                         d_af[b], d_mb[b]);
            if (rc != VLR_OK) break;
            const int64_t ns_ = s1 - s0;
            e = cudaMemcpyAsync(&h_status[c - 1], (const int32_t *)dev_scratch(ctx, 18, 256) + 1, 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_event_ln + s0 * NE, d_ev[b], (size_t)ns_ * NE * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_marginal + s0, d_marg[b], (size_t)ns_ * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_map_af + s0 * NS, d_af[b], (size_t)ns_ * NS * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_map_bias + s0 * NS, d_mb[b], (size_t)ns_ * NS * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_done[b], st);
        }
    }
    if (rc == VLR_OK && e == cudaSuccess) e = cudaMemcpyAsync(h_err, d_err, 8, cudaMemcpyDeviceToHost, st);
    // every exit drains both streams: asynchronous copies may still read the caller's buffers
    const cudaError_t e1 = cudaStreamSynchronize(cp), e2 = cudaStreamSynchronize(st);
    if (rc != VLR_OK) return rc;
    if (e != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess)
        return set_err(ctx, VLR_ERR_CUDA, "wire pipeline failed: %s",
                       cudaGetErrorString(e != cudaSuccess ? e : (e1 != cudaSuccess ? e1 : e2)));
    if (h_err[0])
        return set_err(ctx, VLR_ERR_INVALID, "%d malformed observation vector(s); first: record %d, field %d", h_err[0],
                       (h_err[1] - 1) / 16, (h_err[1] - 1) % 16);
    int64_t n_over = 0;
    for (int c = 0; c < n_chunks; ++c) n_over += h_status[c];
    if (n_over)
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "%lld site(s) exceeded a kernel limit (observations per pileup, "
                       "allele-frequency values or likelihood-table entries per site); their outputs are NaN",
                       (long long)n_over);
    return VLR_OK;
}

}  // extern "C"
