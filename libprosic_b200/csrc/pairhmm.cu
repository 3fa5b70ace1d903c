// pairhmm.cu -- host side of K1 (pair classification, launches, C-ABI entry points).
// The kernel itself is in pairhmm_kernel.cuh, instantiated in pairhmm_inst_*.cu.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pairhmm_host.cuh"
#include "pairhmm_kernel.cuh"
#include "pairhmm_band.cuh"
#include "pairhmm_litband.cuh"
#include "pairhmm_diag.cuh"

namespace vlr {

// ---------------------------------------------------------------- classification of pairs by row count
__host__ __device__ inline int class_of_len(int ly) {
    // R rows per lane over 31 lanes (lane 31 stays a zero padding lane, see hmm_step)
    if (ly <= 31) return 0;
    if (ly <= 62) return 1;
    if (ly <= 93) return 2;
    if (ly <= 124) return 3;
    if (ly <= 155) return 4;
    if (ly <= 186) return 5;
    if (ly <= 248) return 6;
    return 7;
}

// counters layout (int32): [0..7] class counts, [8..15] class offsets, [16..23] scatter cursors,
// [24..31] work cursors (one per class launch), [32] fallback count, [33] fallback work cursor
constexpr int kCntCount = 0, kCntOff = 8, kCntScatter = 16, kCntWork = 24, kCntFb = 32,
              kCntFbWork = 33, kCntFbLong = 34, kCntFbLongWork = 35, kCntGen = 36, kCntGenWork = 37,
              kCntBand = 40 /* four work cursors of the band-only kernels */, kCntTotal = 64;
constexpr int64_t kLongWarpsCap = 160 * kLongBlocksPerSm * kWarps;  // boundary-row slots (<= 160 SMs)

__global__ void classify_kernel(int64_t n_pairs, const int64_t *read_off, const int32_t *pair_read,
                                const uint8_t *mode, int32_t *counters) {
    __shared__ int s_cnt[kNumClasses];
    if (threadIdx.x < kNumClasses) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs;
         p += (int64_t)gridDim.x * blockDim.x) {
        if (mode && mode[p] != 2) continue;
        const int64_t rd = pair_read ? (int64_t)pair_read[p] : p;
        const int ly = (int)(read_off[rd + 1] - read_off[rd]);
        atomicAdd(&s_cnt[class_of_len(ly)], 1);
    }
    __syncthreads();
    if (threadIdx.x < kNumClasses && s_cnt[threadIdx.x])
        atomicAdd(&counters[kCntCount + threadIdx.x], s_cnt[threadIdx.x]);
}

__global__ void class_scan_kernel(int32_t *counters) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int off = 0;
        for (int k = 0; k < kNumClasses; ++k) {
            counters[kCntOff + k] = off;
            off += counters[kCntCount + k];
        }
    }
}

__global__ void scatter_kernel(int64_t n_pairs, const int64_t *read_off, const int32_t *pair_read,
                               const uint8_t *mode, int32_t *counters, int32_t *list) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs;
         p += (int64_t)gridDim.x * blockDim.x) {
        if (mode && mode[p] != 2) continue;
        const int64_t rd = pair_read ? (int64_t)pair_read[p] : p;
        const int ly = (int)(read_off[rd + 1] - read_off[rd]);
        const int k = class_of_len(ly);
        // warp-aggregated cursor bump
        const unsigned peers = __match_any_sync(__activemask(), k);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if ((threadIdx.x & 31) == leader) base = atomicAdd(&counters[kCntScatter + k], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        const int rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
        list[counters[kCntOff + k] + base + rank] = (int32_t)p;
    }
}

// pairs longer than the kernel's row capacity are reported as NaN (host turns this into an error)
__global__ void mark_unsupported_kernel(const int32_t *list, const int32_t *counters, double *out) {
    const int n = counters[kCntCount + 7];
    const int32_t *l = list + counters[kCntOff + 7];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
        out[l[p]] = NAN;
}

// approx: 0 = exact sum, 1 = two-way, 2 = three-way ln_sum3_exp_approx
#define VLR_APPROX_SWITCH(approx, CALL)                                                           \
    switch (approx) {                                                                             \
    case 0: return CALL(0);                                                                       \
    case 1: return CALL(1);                                                                       \
    default: return CALL(2);                                                                      \
    }
static cudaError_t launch_lin(bool banded, int approx, int cls, bool ext, const HmmArgs &a,
                              int sm_count, cudaStream_t st) {
#define VLR_CALL_B(A) launch_pairhmm_lin<true, A>(cls, ext, a, sm_count, st)
#define VLR_CALL_U(A) launch_pairhmm_lin<false, A>(cls, ext, a, sm_count, st)
    if (banded) { VLR_APPROX_SWITCH(approx, VLR_CALL_B) }
    VLR_APPROX_SWITCH(approx, VLR_CALL_U)
#undef VLR_CALL_B
#undef VLR_CALL_U
}
static cudaError_t launch_long(int approx, bool ext, bool log_space, const HmmArgs &a, int sm_count,
                               cudaStream_t st) {
#define VLR_CALL(A) launch_pairhmm_long<A>(ext, log_space, a, sm_count, st)
    VLR_APPROX_SWITCH(approx, VLR_CALL)
#undef VLR_CALL
}
static cudaError_t launch_generic(bool banded, int approx, const HmmArgs &a, int sm_count,
                                  cudaStream_t st) {
#define VLR_CALL_B(A) launch_pairhmm_generic<true, A>(a, sm_count, st)
#define VLR_CALL_U(A) launch_pairhmm_generic<false, A>(a, sm_count, st)
    if (banded) { VLR_APPROX_SWITCH(approx, VLR_CALL_B) }
    VLR_APPROX_SWITCH(approx, VLR_CALL_U)
#undef VLR_CALL_B
#undef VLR_CALL_U
}
static cudaError_t launch_log(bool banded, int approx, const HmmArgs &a, int sm_count,
                              cudaStream_t st) {
#define VLR_CALL_B(A) launch_pairhmm_log<true, A>(a, sm_count, st)
#define VLR_CALL_U(A) launch_pairhmm_log<false, A>(a, sm_count, st)
    if (banded) { VLR_APPROX_SWITCH(approx, VLR_CALL_B) }
    VLR_APPROX_SWITCH(approx, VLR_CALL_U)
#undef VLR_CALL_B
#undef VLR_CALL_U
}
static cudaError_t launch_band(int approx, const HmmArgs &a, int32_t *list_rw, const int32_t *n_total,
                               int32_t *cursors, int sm_count, cudaStream_t st) {
#define VLR_CALL(A) launch_pairhmm_band<A>(a, list_rw, n_total, cursors, sm_count, st)
    VLR_APPROX_SWITCH(approx, VLR_CALL)
#undef VLR_CALL
}
static cudaError_t launch_literal(int approx, const HmmArgs &a, int sm_count, cudaStream_t st) {
#define VLR_CALL(A) launch_pairhmm_literal<A>(a, sm_count, st)
    VLR_APPROX_SWITCH(approx, VLR_CALL)
#undef VLR_CALL
}
static cudaError_t launch_diag(int approx, bool literal, const HmmArgs &a, const DiagArgs &da, int n_ctas,
                               cudaStream_t st) {
#define VLR_CALL(A) launch_pairhmm_diag<A>(literal, a, da, n_ctas, st)
    VLR_APPROX_SWITCH(approx, VLR_CALL)
#undef VLR_CALL
}
// per-warp scratch of the anti-diagonal kernel: three diagonals of (M, X, Y, distance) per read row
constexpr int kDiagBlocksPerSm = 2;
static size_t diag_warp_bytes(int64_t max_ly) { return align_up((size_t)(max_ly + 1) * 84 + 64, 256); }
static int approx_of(uint32_t flags) {
    return (flags & VLR_HMM_SUM3_APPROX3) ? 2 : ((flags & VLR_HMM_SUM3_APPROX) ? 1 : 0);
}

__global__ void set_int_kernel(int32_t *p, int32_t v) { *p = v; }

// PairHMM::new (Appendix A.2): transition constants from the four gap parameters
static void make_consts(const vlr_gap_params *gp, uint32_t flags, HmmConsts *lin, HmmConsts *ln) {
    // the library's deterministic libm (detmath.cuh): the literal log-space kernel must see the
    // same constants as a host evaluation of the reference's formulas
    const double gx = gp->prob_gap_x, gy = gp->prob_gap_y, gxe = gp->prob_gap_x_extend,
                 gye = gp->prob_gap_y_extend;
    const double no_gap = dm::ln_one_minus_exp(dm::ln_add_exp(gx, gy));
    const double no_gxe = dm::ln_one_minus_exp(gxe), no_gye = dm::ln_one_minus_exp(gye);
    ln->t_mm = no_gap;
    if (flags & VLR_HMM_PATH_CLOSE) { ln->t_xm = no_gye; ln->t_ym = no_gxe; }
    else { ln->t_xm = no_gxe; ln->t_ym = no_gye; }
    ln->t_gy = gy; ln->t_gye = gye; ln->t_gx = gx; ln->t_gxe = gxe;
    lin->t_mm = dm::exp(ln->t_mm); lin->t_xm = dm::exp(ln->t_xm); lin->t_ym = dm::exp(ln->t_ym);
    lin->t_gy = dm::exp(gy); lin->t_gye = dm::exp(gye); lin->t_gx = dm::exp(gx); lin->t_gxe = dm::exp(gxe);
    // scaled-state kernels keep M' = t_mm * M (pairhmm_kernel.cuh, hmm_step<..., LUT = true>)
    lin->t_gy_s = lin->t_gy / lin->t_mm;
    lin->t_gx_s = lin->t_gx / lin->t_mm;
    lin->inv_tmm = 1.0 / lin->t_mm;
    lin->one_s = VLR_SCALE * lin->t_mm;
    ln->t_gy_s = ln->t_gx_s = ln->inv_tmm = ln->one_s = NAN;  // unused in log space
}

struct PairWs {
    int32_t *counters;      // kCntTotal
    int32_t *list;          // n_pairs
    int32_t *fb_list;       // n_pairs
    int32_t *fb_long_list;  // n_pairs
    int32_t *gen_list;      // n_pairs
    double *long_ws;        // boundary rows of the long-read variant
};
static size_t pair_ws_bytes(int64_t n_pairs) {
    return align_up(kCntTotal * sizeof(int32_t), 256) + 4 * align_up((size_t)n_pairs * 4 + 4, 256);
}
static PairWs carve_ws(void *ws, int64_t n_pairs) {
    PairWs w;
    uint8_t *p = (uint8_t *)ws;
    w.counters = (int32_t *)p;
    p += align_up(kCntTotal * sizeof(int32_t), 256);
    w.list = (int32_t *)p;
    p += align_up((size_t)n_pairs * 4 + 4, 256);
    w.fb_list = (int32_t *)p;
    p += align_up((size_t)n_pairs * 4 + 4, 256);
    w.fb_long_list = (int32_t *)p;
    p += align_up((size_t)n_pairs * 4 + 4, 256);
    w.gen_list = (int32_t *)p;
    p += align_up((size_t)n_pairs * 4 + 4, 256);
    w.long_ws = (double *)p;
    return w;
}

}  // namespace vlr

using namespace vlr;

static int pairhmm_core_dev(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *d_hap_bytes,
                            const int64_t *d_hap_off, int64_t n_haps, const uint8_t *d_read_bases,
                            const uint8_t *d_read_quals, const int64_t *d_read_off, int64_t n_reads,
                            const int32_t *d_pair_hap, const int32_t *d_pair_read,
                            const int32_t *d_max_edit_dist, const int64_t *d_win_start,
                            const int32_t *d_win_len, const uint8_t *d_mode,
                            const vlr_gap_params *gap, uint32_t flags, void *d_workspace,
                            size_t workspace_bytes, double *d_out_lnprob) {
    if (!ctx) return VLR_ERR_INVALID;
    if (n_pairs < 0 || n_pairs > 0x7fffffff)
        return set_err(ctx, VLR_ERR_INVALID, "n_pairs out of range");
    if (n_pairs == 0) return VLR_OK;
    if (!d_hap_bytes || (!d_hap_off && !d_win_start) || !d_read_bases || !d_read_quals || !d_read_off || !gap ||
        !d_workspace || !d_out_lnprob)
        return set_err(ctx, VLR_ERR_INVALID, "null pointer argument");
    if ((!d_win_start && !d_pair_hap && n_haps < n_pairs) || (!d_pair_read && n_reads < n_pairs))
        return set_err(ctx, VLR_ERR_INVALID, "identity pair mapping needs n_haps/n_reads >= n_pairs");
    if (workspace_bytes < pair_ws_bytes(n_pairs))
        return set_err(ctx, VLR_ERR_INVALID, "workspace too small");
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PairWs w = carve_ws(d_workspace, n_pairs);
    // whatever the caller allocated beyond the bookkeeping is boundary-row space for long reads
    const int64_t long_stride =
        (int64_t)((workspace_bytes - pair_ws_bytes(n_pairs)) / (size_t)(kLongWarpsCap * 6 * sizeof(double)));
    // the same space, cut per warp, is the scratch of the anti-diagonal kernel (banded long windows)
    const int64_t diag_bytes =
        (int64_t)((workspace_bytes - pair_ws_bytes(n_pairs)) / (size_t)kLongWarpsCap) & ~(int64_t)255;
    VLR_CUDA(ctx, cudaMemsetAsync(w.counters, 0, kCntTotal * sizeof(int32_t), st));
    const int64_t want = (n_pairs + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    const int cgrid = (int)(want < cap ? want : cap);
    classify_kernel<<<cgrid, 256, 0, st>>>(n_pairs, d_read_off, d_pair_read, d_mode, w.counters);
    VLR_LAUNCH_CHECK(ctx);
    class_scan_kernel<<<1, 32, 0, st>>>(w.counters);
    VLR_LAUNCH_CHECK(ctx);
    scatter_kernel<<<cgrid, 256, 0, st>>>(n_pairs, d_read_off, d_pair_read, d_mode, w.counters, w.list);
    VLR_LAUNCH_CHECK(ctx);

    HmmArgs a;
    a.hap = d_hap_bytes; a.hap_off = d_hap_off;
    a.rbase = d_read_bases; a.rqual = d_read_quals; a.read_off = d_read_off;
    a.pair_hap = d_pair_hap; a.pair_read = d_pair_read; a.med = d_max_edit_dist;
    a.win_start = d_win_start; a.win_len = d_win_len;
    a.fb_list = w.fb_list; a.fb_count = w.counters + kCntFb;
    a.gen_list = w.gen_list; a.gen_count = w.counters + kCntGen;
    a.out = d_out_lnprob; a.qlut = ctx->d_qlut;
    a.long_ws = w.long_ws; a.long_stride = long_stride;
    make_consts(gap, flags, &a.lin, &a.ln);
    const bool banded = d_max_edit_dist != nullptr;
    const int approx = approx_of(flags);
    // extension probabilities of exactly 0 (CLI default) select the 5-op specialisation
    const bool ext = !(gap->prob_gap_x_extend == -INFINITY && gap->prob_gap_y_extend == -INFINITY);

    // Banded pairs on narrow windows (the allele-support path): the band-only kernel takes what it
    // can and marks those list entries; the class kernels below skip them.  Needs gap-extension
    // probabilities of 0 and small gap-open probabilities (its error bound, pairhmm_band.cuh).
    {
        const char *e = getenv("VLR_K1_BAND");
        const double small = -9.210340371976182;  // ln 1e-4
        if (banded && !ext && gap->prob_gap_x <= small && gap->prob_gap_y <= small && !(e && atoi(e) == 0)) {
            HmmArgs ab = a;
            ab.list = w.list;
            cudaError_t ce = launch_band(approx, ab, w.list, w.counters + kCntOff + 7, w.counters + kCntBand,
                                         ctx->sm_count, st);
            if (ce != cudaSuccess)
                return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(ce));
            ctx->launches += 4;
        }
    }
    // Class sizes live on the device: every class kernel is launched as a persistent grid with a
    // dynamic work counter and exits at once when its list is empty, so the device entry point
    // never synchronises with the host.
    for (int cls = 0; cls < 7; ++cls) {
        HmmArgs ac = a;
        ac.list = w.list;
        ac.list_off = w.counters + kCntOff + cls;
        ac.count = w.counters + kCntCount + cls;
        ac.cursor = w.counters + kCntWork + cls;
        cudaError_t e = launch_lin(banded, approx, cls, ext, ac, ctx->sm_count, st);
        if (e != cudaSuccess)
            return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    // reads with a base outside ACGT (N, IUPAC codes): same arithmetic, emission by byte compare
    {
        HmmArgs ac = a;
        ac.list = w.gen_list;
        ac.list_off = nullptr;
        ac.count = w.counters + kCntGen;
        ac.cursor = w.counters + kCntGenWork;
        cudaError_t e = launch_generic(banded, approx, ac, ctx->sm_count, st);
        if (e != cudaSuccess)
            return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    // log-space pass over the pairs whose scaled-linear result left the fp64 range
    {
        HmmArgs ac = a;
        ac.list = w.fb_list;
        ac.list_off = nullptr;
        ac.count = w.counters + kCntFb;
        ac.cursor = w.counters + kCntFbWork;
        cudaError_t e = launch_log(banded, approx, ac, ctx->sm_count, st);
        if (e != cudaSuccess)
            return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    if (!banded && !d_win_start && long_stride > 0 && ctx->sm_count <= 160) {
        // read windows longer than 248 rows: strip-wise variant (+ its log-space pass)
        HmmArgs al = a;
        al.list = w.list;
        al.list_off = w.counters + kCntOff + 7;
        al.count = w.counters + kCntCount + 7;
        al.cursor = w.counters + kCntWork + 7;
        al.fb_list = w.fb_long_list;
        al.fb_count = w.counters + kCntFbLong;
        cudaError_t e = launch_long(approx, ext, false, al, ctx->sm_count, st);
        if (e != cudaSuccess) return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
        al.list = w.fb_long_list;
        al.list_off = nullptr;
        al.count = w.counters + kCntFbLong;
        al.cursor = w.counters + kCntFbLongWork;
        e = launch_long(approx, ext, true, al, ctx->sm_count, st);
        if (e != cudaSuccess) return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    } else if (diag_bytes >= (int64_t)diag_warp_bytes(249) && ctx->sm_count <= 160) {
        // banded pairs / explicit windows with read windows above 248 rows: anti-diagonal log-space sweep
        HmmArgs ad = a;
        ad.list = w.list;
        ad.list_off = w.counters + kCntOff + 7;
        ad.count = w.counters + kCntCount + 7;
        ad.cursor = w.counters + kCntWork + 7;
        DiagArgs da;
        da.ws = (uint8_t *)w.long_ws; da.ws_bytes = diag_bytes; da.min_rows = 0;
        cudaError_t e = launch_diag(approx, false, ad, da, ctx->sm_count * kDiagBlocksPerSm, st);
        if (e != cudaSuccess) return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    } else {
        mark_unsupported_kernel<<<32, 256, 0, st>>>(w.list, w.counters, d_out_lnprob);
        VLR_LAUNCH_CHECK(ctx);
    }
    return VLR_OK;
}


namespace vlr {
// Literal log-space evaluation (LogLit) of the pairs `d_list[0 .. *d_count)` (d_list == nullptr:
// pairs 0 .. n_pairs_cap-1); windows as in pairhmm_core_dev.  max_lx sizes the per-warp prob_cols.
int pairhmm_literal_dev(vlr_ctx *ctx, int64_t n_pairs_cap, const uint8_t *d_hap_bytes, const int64_t *d_hap_off,
                        const int64_t *d_win_start, const int32_t *d_win_len, const uint8_t *d_read_bases,
                        const uint8_t *d_read_quals, const int64_t *d_read_off, const int32_t *d_pair_hap,
                        const int32_t *d_pair_read, const int32_t *d_med, const int32_t *d_list,
                        const int32_t *d_count, const vlr_gap_params *gap, uint32_t flags, int64_t max_lx,
                        int64_t max_ly, double *d_out) {
    enum { S_LIT = 90, S_LITCNT = 91, S_LITDIAG = 95 };
    cudaStream_t st = ctx->stream;
    const int64_t stride = (max_lx + 31) / 32 * 32;
    const size_t warps = (size_t)ctx->sm_count * kLitBlocksPerSm * kWarps;
    double *ws = (double *)dev_scratch(ctx, S_LIT, warps * 3 * (size_t)stride * sizeof(double));
    int32_t *cnt = (int32_t *)dev_scratch(ctx, S_LITCNT, 64);
    if (!ws || !cnt) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    VLR_CUDA(ctx, cudaMemsetAsync(cnt, 0, 64, st));
    if (!d_count) {
        set_int_kernel<<<1, 1, 0, st>>>(cnt + 1, (int32_t)n_pairs_cap);
        VLR_LAUNCH_CHECK(ctx);
    }
    HmmArgs a;
    memset(&a, 0, sizeof(a));
    a.hap = d_hap_bytes; a.hap_off = d_hap_off;
    a.rbase = d_read_bases; a.rqual = d_read_quals; a.read_off = d_read_off;
    a.pair_hap = d_pair_hap; a.pair_read = d_pair_read; a.med = d_med;
    a.win_start = d_win_start; a.win_len = d_win_len;
    a.list = d_list; a.list_off = nullptr;
    a.count = d_count ? d_count : cnt + 1;
    a.cursor = cnt;
    a.out = d_out; a.qlut = ctx->d_qlut;
    a.lit_ws = ws; a.lit_stride = stride;
    make_consts(gap, flags, &a.lin, &a.ln);
    // Banded pairs go through the row-sweep kernel (pairhmm_litband.cuh); what it does not take (no band,
    // gap extension on the haplotype side, very long windows) is left on `rest` for the wavefront kernel
    const char *lbenv = getenv("VLR_LIT_BAND");
    if (d_med && !(lbenv && atoi(lbenv) == 0)) {
        enum { S_LITREST = 94 };
        int32_t *rest = (int32_t *)dev_scratch(ctx, S_LITREST, (size_t)n_pairs_cap * 4 + 64);
        if (!rest) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        int32_t *rest_count = cnt + 2;  // cnt[0]: cursor of this launch, cnt[3]: cursor of the next
        const int n_ctas = (int)std::min<size_t>(warps, (size_t)ctx->sm_count * 6);
        cudaError_t e = cudaSuccess;
        switch (approx_of(flags)) {
        case 0: e = launch_pairhmm_litband<0>(a, rest, rest_count, n_ctas, st); break;
        case 1: e = launch_pairhmm_litband<1>(a, rest, rest_count, n_ctas, st); break;
        default: e = launch_pairhmm_litband<2>(a, rest, rest_count, n_ctas, st); break;
        }
        if (e != cudaSuccess) return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
        a.list = rest; a.list_off = nullptr;
        a.count = rest_count;
        a.cursor = cnt + 3;
    }
    cudaError_t e = launch_literal(approx_of(flags), a, ctx->sm_count, st);
    if (e != cudaSuccess) return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    if (max_ly > 248) {
        // read windows above the wavefront kernel's 248 rows (it skips them): anti-diagonal sweep over the
        // same list, same arithmetic.  kLitBlocksPerSm == kDiagBlocksPerSm: the prob_cols slots match.
        static_assert(kLitBlocksPerSm == kDiagBlocksPerSm, "per-warp prob_cols are indexed by the same warp id");
        DiagArgs da;
        da.ws_bytes = (int64_t)diag_warp_bytes(max_ly);
        da.ws = (uint8_t *)dev_scratch(ctx, S_LITDIAG, warps * (size_t)da.ws_bytes);
        da.min_rows = 249;
        if (!da.ws) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        a.cursor = cnt + 4;
        e = launch_diag(approx_of(flags), true, a, da, ctx->sm_count * kDiagBlocksPerSm, st);
        if (e != cudaSuccess) return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    return VLR_OK;
}

int pairhmm_windows_dev(vlr_ctx *ctx, int64_t n_slots, const uint8_t *d_hap_base,
                        const int64_t *d_win_start, const int32_t *d_win_len,
                        const uint8_t *d_read_bases, const uint8_t *d_read_quals,
                        const int64_t *d_read_off, const int32_t *d_pair_read, const int32_t *d_med,
                        const uint8_t *d_mode, const vlr_gap_params *gap, uint32_t flags,
                        void *d_workspace, size_t workspace_bytes, double *d_out) {
    return pairhmm_core_dev(ctx, n_slots, d_hap_base, nullptr, 0, d_read_bases, d_read_quals, d_read_off,
                            (int64_t)1 << 40, nullptr, d_pair_read, d_med, d_win_start, d_win_len, d_mode,
                            gap, flags, d_workspace, workspace_bytes, d_out);
}
}  // namespace vlr

extern "C" {

size_t vlr_pairhmm_workspace_bytes(int64_t n_pairs, int64_t max_len_x, int64_t max_len_y) {
    size_t b = pair_ws_bytes(n_pairs < 1 ? 1 : n_pairs);
    if (max_len_y > 248 && max_len_x > 0)  // boundary rows of the strip-wise long-read variant or the
                                           // three anti-diagonals of the banded one, whichever is larger
        b += (size_t)kLongWarpsCap * std::max<size_t>(6 * sizeof(double) * (size_t)(max_len_x + 32),
                                                      diag_warp_bytes(max_len_y));
    return b;
}

int vlr_pairhmm_batch_dev(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *d_hap_bytes,
                          const int64_t *d_hap_off, int64_t n_haps, const uint8_t *d_read_bases,
                          const uint8_t *d_read_quals, const int64_t *d_read_off, int64_t n_reads,
                          const int32_t *d_pair_hap, const int32_t *d_pair_read,
                          const int32_t *d_max_edit_dist, const vlr_gap_params *gap, uint32_t flags,
                          void *d_workspace, size_t workspace_bytes, double *d_out_lnprob) {
    if (ctx && (flags & VLR_HMM_LITERAL))
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "VLR_HMM_LITERAL needs the host-buffer entry (window lengths size its scratch)");
    return pairhmm_core_dev(ctx, n_pairs, d_hap_bytes, d_hap_off, n_haps, d_read_bases, d_read_quals,
                            d_read_off, n_reads, d_pair_hap, d_pair_read, d_max_edit_dist, nullptr,
                            nullptr, nullptr, gap, flags, d_workspace, workspace_bytes, d_out_lnprob);
}

static int vlr_pairhmm_batch_impl(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes,
                      const int64_t *hap_off, int64_t n_haps, const uint8_t *read_bases,
                      const uint8_t *read_quals, const int64_t *read_off, int64_t n_reads,
                      const int32_t *pair_hap, const int32_t *pair_read,
                      const int32_t *max_edit_dist, const vlr_gap_params *gap, uint32_t flags,
                      double *out_lnprob) {
    if (!ctx) return VLR_ERR_INVALID;
    if (n_pairs < 0 || n_pairs > 0x7fffffff) return set_err(ctx, VLR_ERR_INVALID, "n_pairs out of range");
    if (n_pairs == 0) return VLR_OK;
    if (!hap_bytes || !hap_off || !read_bases || !read_quals || !read_off || !gap || !out_lnprob ||
        n_haps <= 0 || n_reads <= 0)
        return set_err(ctx, VLR_ERR_INVALID, "null/empty argument");
    if ((!pair_hap && n_haps < n_pairs) || (!pair_read && n_reads < n_pairs))
        return set_err(ctx, VLR_ERR_INVALID, "identity pair mapping needs n_haps/n_reads >= n_pairs");
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // validate indices and lengths on the host (errors must not reach the kernels); a few threads, each
    // on a contiguous range: one core's memory bandwidth made this 5 ms of a 144 ms call at 2 M pairs
    struct Part { int64_t max_ly = 0, bad_hap = -1, bad_read = -1; bool bad_off = false, unordered = false; };
    const int vt = validation_threads(n_pairs + n_reads);
    std::vector<Part> parts((size_t)vt);
    parallel_ranges(n_reads, vt, [&](int t, int64_t r0, int64_t r1) {
        Part &q = parts[(size_t)t];
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t l = read_off[r + 1] - read_off[r];
            if (l < 0) q.bad_off = true;
            if (l > q.max_ly) q.max_ly = l;
        }
    });
    if (pair_hap || pair_read)
        parallel_ranges(n_pairs, vt, [&](int t, int64_t k0, int64_t k1) {
            Part &q = parts[(size_t)t];
            for (int64_t k = k0; k < k1; ++k) {
                if (pair_hap && (pair_hap[k] < 0 || pair_hap[k] >= n_haps) && q.bad_hap < 0) q.bad_hap = k;
                if (pair_read) {
                    if ((pair_read[k] < 0 || pair_read[k] >= n_reads) && q.bad_read < 0) q.bad_read = k;
                    if (k > 0 && pair_read[k] < pair_read[k - 1]) q.unordered = true;
                }
            }
        });
    int64_t max_ly = 0;
    bool reads_in_pair_order = pair_hap && pair_read;  // pair_read non-decreasing: reads can stream in chunks
    for (const Part &q : parts) {
        if (q.bad_off) return set_err(ctx, VLR_ERR_INVALID, "read_off not monotone");
        if (q.bad_hap >= 0) return set_err(ctx, VLR_ERR_INVALID, "pair_hap[%lld] out of range", (long long)q.bad_hap);
        if (q.bad_read >= 0) return set_err(ctx, VLR_ERR_INVALID, "pair_read[%lld] out of range", (long long)q.bad_read);
        if (q.unordered) reads_in_pair_order = false;
        max_ly = std::max(max_ly, q.max_ly);
    }
    const int64_t hap_lo = hap_off[0], hap_hi = hap_off[n_haps];
    const int64_t rd_lo = read_off[0], rd_hi = read_off[n_reads];
    int64_t max_lx = 0;
    for (int64_t h = 0; h < n_haps; ++h) {
        if (hap_off[h + 1] < hap_off[h]) return set_err(ctx, VLR_ERR_INVALID, "hap_off not monotone");
        if (hap_off[h + 1] - hap_off[h] > max_lx) max_lx = hap_off[h + 1] - hap_off[h];
    }
    enum { S_HAP = 0, S_HOFF, S_RB, S_RQ, S_ROFF, S_PH, S_PR, S_MED, S_WS, S_OUT };
    const size_t hap_bytes_n = (size_t)(hap_hi - hap_lo), rd_bytes_n = (size_t)(rd_hi - rd_lo);
    uint8_t *d_hap = (uint8_t *)dev_scratch(ctx, S_HAP, hap_bytes_n + 64);
    int64_t *d_hoff = (int64_t *)dev_scratch(ctx, S_HOFF, (size_t)(n_haps + 1) * 8);
    uint8_t *d_rb = (uint8_t *)dev_scratch(ctx, S_RB, rd_bytes_n + 16);
    uint8_t *d_rq = (uint8_t *)dev_scratch(ctx, S_RQ, rd_bytes_n + 16);
    int64_t *d_roff = (int64_t *)dev_scratch(ctx, S_ROFF, (size_t)(n_reads + 1) * 8);
    int32_t *d_ph = pair_hap ? (int32_t *)dev_scratch(ctx, S_PH, (size_t)n_pairs * 4) : nullptr;
    int32_t *d_pr = pair_read ? (int32_t *)dev_scratch(ctx, S_PR, (size_t)n_pairs * 4) : nullptr;
    int32_t *d_med = max_edit_dist ? (int32_t *)dev_scratch(ctx, S_MED, (size_t)n_pairs * 4) : nullptr;
    const size_t ws_bytes = vlr_pairhmm_workspace_bytes(n_pairs, max_lx, max_ly);
    void *d_ws = dev_scratch(ctx, S_WS, ws_bytes);
    double *d_out = (double *)dev_scratch(ctx, S_OUT, (size_t)n_pairs * 8);
    if (!d_hap || !d_hoff || !d_rb || !d_rq || !d_roff || !d_ws || !d_out || (pair_hap && !d_ph) ||
        (pair_read && !d_pr) || (max_edit_dist && !d_med))
        return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    // H2D straight from the caller's buffers (full speed when they are pinned)
    VLR_CUDA(ctx, cudaMemcpyAsync(d_hap, hap_bytes + hap_lo, hap_bytes_n, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_hoff, hap_off, (size_t)(n_haps + 1) * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_roff, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    const double stream_bytes = 2.0 * (double)rd_bytes_n + 12.0 * (double)n_pairs;
    // Chunk sizes grow geometrically (x3 from ~8 MB): the first copy, which nothing hides, is short, the
    // copy of chunk c+1 still hides under the kernels of chunk c as long as they take a third of its
    // transfer time per byte (the C3 shape computes 20x longer than it copies), and the batch needs 4-5
    // kernel chains instead of the 10 of equal 32 MB chunks -- each chain ends in the tails of its
    // persistent kernels and of the small re-evaluation launches (~0.4 ms per chunk on the C3 shape).
    std::vector<int64_t> chunk_end;   // pair index where each chunk ends
    if (reads_in_pair_order && stream_bytes >= 64.0 * 1048576.0) {
        const char *ce = getenv("VLR_K1_CHUNK_MB");   // tuning aid: size of the first chunk
        double want = (ce && atof(ce) > 0 ? atof(ce) : 8.0) * 1048576.0, done = 0.0;
        while (done + want * 1.34 < stream_bytes && chunk_end.size() < 31) {  // no sliver at the end
            done += want;
            chunk_end.push_back((int64_t)((double)n_pairs * (done / stream_bytes)));
            want *= 3.0;
        }
    }
    chunk_end.push_back(n_pairs);
    int n_chunks = (int)chunk_end.size();
    const bool literal = (flags & VLR_HMM_LITERAL) != 0;
    if (literal) n_chunks = 1;
    if (n_chunks > 1 && ensure_copy_stream(ctx) != VLR_OK) n_chunks = 1;
    if (n_chunks <= 1) {
        VLR_CUDA(ctx, cudaMemcpyAsync(d_rb, read_bases + rd_lo, rd_bytes_n, cudaMemcpyHostToDevice, st));
        VLR_CUDA(ctx, cudaMemcpyAsync(d_rq, read_quals + rd_lo, rd_bytes_n, cudaMemcpyHostToDevice, st));
        if (pair_hap) VLR_CUDA(ctx, cudaMemcpyAsync(d_ph, pair_hap, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
        if (pair_read) VLR_CUDA(ctx, cudaMemcpyAsync(d_pr, pair_read, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
        if (max_edit_dist)
            VLR_CUDA(ctx, cudaMemcpyAsync(d_med, max_edit_dist, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
        // offsets are absolute into the caller's arrays: rebase the device pointers instead of the offsets
        int rc = literal
                     ? pairhmm_literal_dev(ctx, n_pairs, d_hap - hap_lo, d_hoff, nullptr, nullptr, d_rb - rd_lo,
                                           d_rq - rd_lo, d_roff, d_ph, d_pr, d_med, nullptr, nullptr, gap, flags,
                                           max_lx, max_ly, d_out)
                     : vlr_pairhmm_batch_dev(ctx, n_pairs, d_hap - hap_lo, d_hoff, n_haps, d_rb - rd_lo,
                                             d_rq - rd_lo, d_roff, n_reads, d_ph, d_pr, d_med, gap, flags, d_ws,
                                             ws_bytes, d_out);
        if (rc != VLR_OK) return rc;
    } else {
        // Pairs reference their reads in non-decreasing order: stream reads and pair lists in
        // chunks on the copy stream while the main stream computes the previous chunk.  All device
        // buffers are full size, so chunks never alias; one event per chunk orders copy -> kernel.
        cudaStream_t cp = ctx->copy_stream;
        std::vector<cudaEvent_t> evs((size_t)n_chunks, nullptr);
        cudaEvent_t ev_start = nullptr;
        auto cleanup = [&]() {
            for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
            if (ev_start) cudaEventDestroy(ev_start);
        };
        // the copy stream must not overtake earlier work on the main stream that still reads the buffers
        if (cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(ev_start, st) != cudaSuccess || cudaStreamWaitEvent(cp, ev_start, 0) != cudaSuccess) {
            cleanup();
            return set_err(ctx, VLR_ERR_CUDA, "event setup failed");
        }
        int64_t p0 = 0;
        int rc = VLR_OK;
        for (int c = 0; c < n_chunks && rc == VLR_OK; ++c) {
            const int64_t p1 = chunk_end[(size_t)c];
            if (p1 <= p0) continue;
            // reads [ra, rb) are first needed by this chunk (earlier chunks brought everything below ra)
            const int64_t ra = c == 0 ? 0 : (int64_t)pair_read[p0 - 1] + 1;
            const int64_t rb = c + 1 == n_chunks ? n_reads : (int64_t)pair_read[p1 - 1] + 1;
            bool ok = cudaEventCreateWithFlags(&evs[c], cudaEventDisableTiming) == cudaSuccess;
            if (ok && rb > ra) {
                const int64_t b0 = read_off[ra] - rd_lo, nb = read_off[rb] - read_off[ra];
                ok = ok && cudaMemcpyAsync(d_rb + b0, read_bases + read_off[ra], (size_t)nb, cudaMemcpyHostToDevice, cp) == cudaSuccess;
                ok = ok && cudaMemcpyAsync(d_rq + b0, read_quals + read_off[ra], (size_t)nb, cudaMemcpyHostToDevice, cp) == cudaSuccess;
            }
            ok = ok && cudaMemcpyAsync(d_ph + p0, pair_hap + p0, (size_t)(p1 - p0) * 4, cudaMemcpyHostToDevice, cp) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(d_pr + p0, pair_read + p0, (size_t)(p1 - p0) * 4, cudaMemcpyHostToDevice, cp) == cudaSuccess;
            if (max_edit_dist)
                ok = ok && cudaMemcpyAsync(d_med + p0, max_edit_dist + p0, (size_t)(p1 - p0) * 4, cudaMemcpyHostToDevice, cp) == cudaSuccess;
            ok = ok && cudaEventRecord(evs[c], cp) == cudaSuccess && cudaStreamWaitEvent(st, evs[c], 0) == cudaSuccess;
            if (!ok) { rc = set_err(ctx, VLR_ERR_CUDA, "chunk copy failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
            rc = vlr_pairhmm_batch_dev(ctx, p1 - p0, d_hap - hap_lo, d_hoff, n_haps, d_rb - rd_lo, d_rq - rd_lo,
                                       d_roff, n_reads, d_ph + p0, d_pr + p0, d_med ? d_med + p0 : nullptr, gap,
                                       flags, d_ws, ws_bytes, d_out + p0);
            if (rc == VLR_OK &&
                cudaMemcpyAsync(out_lnprob + p0, d_out + p0, (size_t)(p1 - p0) * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                rc = set_err(ctx, VLR_ERR_CUDA, "D2H failed");
            p0 = p1;
        }
        cudaStreamSynchronize(cp);
        cudaError_t es = cudaStreamSynchronize(st);
        cleanup();
        if (rc != VLR_OK) return rc;
        if (es != cudaSuccess) return set_err(ctx, VLR_ERR_CUDA, "pairhmm pipeline failed: %s", cudaGetErrorString(es));
        return VLR_OK;
    }
    VLR_CUDA(ctx, cudaMemcpyAsync(out_lnprob, d_out, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    if (getenv("VLR_TRACE") && !literal) {  // how many pairs took the slower kernels
        int32_t cnt[kCntTotal];
        if (cudaMemcpy(cnt, d_ws, sizeof(cnt), cudaMemcpyDeviceToHost) == cudaSuccess)
            fprintf(stderr, "[vlr trace] pairhmm %lld pairs: exact-decision / generic kernel %d, log-space pass %d, long %d\n",
                    (long long)n_pairs, cnt[kCntGen], cnt[kCntFb], cnt[kCntCount + 7]);
    }
    return VLR_OK;
}

int vlr_pairhmm_batch(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes,
                      const int64_t *hap_off, int64_t n_haps, const uint8_t *read_bases,
                      const uint8_t *read_quals, const int64_t *read_off, int64_t n_reads,
                      const int32_t *pair_hap, const int32_t *pair_read,
                      const int32_t *max_edit_dist, const vlr_gap_params *gap, uint32_t flags,
                      double *out_lnprob) {
    return vlr::drain_on_error(ctx, vlr_pairhmm_batch_impl(ctx, n_pairs, hap_bytes, hap_off, n_haps, read_bases, read_quals, read_off, n_reads, pair_hap, pair_read, max_edit_dist, gap, flags, out_lnprob));
}


}  // extern "C"
