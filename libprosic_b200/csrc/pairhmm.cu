// pairhmm.cu -- K1: anti-diagonal wavefront pair-HMM forward kernel for sm_100a.
//
// Replaces bio::stats::pairhmm::PairHMM::prob_related as called from
// PairHMMRealigner::calculate_prob_allele (reference src/variants/evidence/realignment/mod.rs:430-444)
// with the emission model of realignment/pairhmm.rs:123-211 and the semiglobal start/end rules of
// pairhmm.rs:103-121.  Recurrence (x = haplotype column i, y = read row j):
//   M[i][j] = e_xy(i,j) * sum3( t_mm*M[i-1][j-1], t_xm*X[i-1][j-1], t_ym*Y[i-1][j-1] )
//   X[i][j] =             t_gy*M[i-1][j] + t_gye*X[i-1][j]            (hap-only state, emit 1)
//   Y[i][j] = eps_j *   ( t_gx*M[i][j-1] + t_gxe*Y[i][j-1] )          (read-only state)
//   M[i][0] = 1 for every column (free start), result = sum_i (M+X+Y)[i][len_y] (free end), capped at 1.
//
// Mapping: one warp per (read, haplotype) pair.  Lane l owns R consecutive read rows
// [l*R, l*R+R); at step t it computes column i = t - l for its rows, so the warp sweeps
// anti-diagonals.  The bottom row of lane l-1 travels to lane l by __shfl_up_sync (M, X, Y, and the
// band's min-edit-distance); the haplotype window is staged into a per-warp shared-memory ring by
// 1-D TMA bulk copies (cp.async.bulk + mbarrier), read bases/qualities become per-row register
// constants from a 256-entry LUT.  Arithmetic is fp64 in SCALED LINEAR space (start mass 2^960,
// no per-cell transcendental, one log per pair); pairs whose result leaves the representable
// range are re-run by the same kernel instantiated with log-space arithmetic (still on the GPU).
#include <math.h>

#include <string.h>

#include <vector>

#include "common.cuh"

namespace vlr {

constexpr int kWarps = 4;          // warps per CTA; each warp owns one pair at a time
constexpr int kChunk = 256;        // haplotype bytes per TMA bulk copy
constexpr int kRing = 2 * kChunk;  // per-warp ring buffer (two chunks)
constexpr int kMedMax = 1 << 20;   // "usize::MAX" of the band bookkeeping
constexpr int kNumClasses = 8;     // R = 1,2,3,4,5,6,8 and "too long"

// 2^960 and its log; linear-space values are scaled by this (see header comment)
#define VLR_SCALE 9.7453140114e288
__device__ __constant__ double c_scale_ln = 665.4212933375474;  // 960 * ln 2

struct HmmConsts {  // linear or ln, depending on the arithmetic policy
    double t_mm, t_xm, t_ym, t_gy, t_gye, t_gx, t_gxe;
};

struct HmmArgs {
    const uint8_t *hap;
    const int64_t *hap_off;
    const uint8_t *rbase;
    const uint8_t *rqual;
    const int64_t *read_off;
    const int32_t *pair_hap;
    const int32_t *pair_read;
    const int32_t *med;
    const int32_t *list;      // pair ids handled by this launch (device)
    const int32_t *list_off;  // device offset into `list` (nullable)
    const int32_t *count;  // number of entries of `list` (device)
    int32_t *cursor;       // dynamic work counter (device, zero at launch)
    int32_t *fb_list;      // pairs that need the log-space pass
    int32_t *fb_count;
    double *out;
    const double *qlut;
    HmmConsts lin, ln;
};

// ---------------------------------------------------------------- PTX helpers (TMA + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VLR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra VLR_DONE;\n"
        "bra VLR_WAIT;\n"
        "VLR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------- arithmetic policies
__device__ __forceinline__ double ln_add_exp_d(double a, double b) {
    double p0 = fmax(a, b), p1 = fmin(a, b);
    if (p0 == -INFINITY) return -INFINITY;
    if (p1 == -INFINITY) return p0;
    return p0 + log1p(exp(p1 - p0));
}

struct Lin {  // scaled linear space
    static constexpr bool kLog = false;
    __device__ static double zero() { return 0.0; }
    __device__ static double one() { return VLR_SCALE; }
    __device__ static double mul(double a, double b) { return a * b; }
    __device__ static double add(double a, double b) { return a + b; }
    // a*b + c*d
    __device__ static double dot2(double a, double b, double c, double d) { return fma(a, b, c * d); }
    template <bool APPROX>
    __device__ static double sum3(double a, double b, double c) {
        if (APPROX) {
            // bio pairhmm ln_sum3_exp_approx: keep only the maximum if the second largest is more
            // than e^-10 below it, else the exact sum
            double hi = a > b ? a : b, lo = a > b ? b : a;
            double mx = hi > c ? hi : c;
            double hc = hi > c ? c : hi;
            double second = lo > hc ? lo : hc;
            return (second < mx * 4.5399929762484854e-05) ? mx : (a + b + c);
        }
        return a + b + c;
    }
};

struct Log {  // natural-log space, the literal arithmetic of the reference
    static constexpr bool kLog = true;
    __device__ static double zero() { return -INFINITY; }
    __device__ static double one() { return 0.0; }
    __device__ static double mul(double a, double b) { return a + b; }
    __device__ static double add(double a, double b) { return ln_add_exp_d(a, b); }
    __device__ static double dot2(double a, double b, double c, double d) {
        return ln_add_exp_d(a + b, c + d);
    }
    template <bool APPROX>
    __device__ static double sum3(double a, double b, double c) {
        double p0 = a, p1 = b, p2 = c, t;
        if (p1 < p2) { t = p1; p1 = p2; p2 = t; }
        if (p1 > p0) { t = p1; p1 = p0; p0 = t; }
        if (p2 > p0) { t = p2; p2 = p0; p0 = t; }
        double second = fmax(p1, p2);
        if (APPROX && (p0 - second > 10.0)) return p0;
        if (p0 == -INFINITY) return -INFINITY;
        double s = 0.0;
        if (p1 != -INFINITY) s += exp(p1 - p0);
        if (p2 != -INFINITY) s += exp(p2 - p0);
        return p0 + log1p(s);
    }
};

// ---------------------------------------------------------------- the kernel
template <int R>
struct Occ {
    static constexpr int kMinBlocks = R <= 3 ? 5 : (R <= 4 ? 4 : (R <= 6 ? 3 : 2));
};

template <int R, bool BANDED, bool APPROX, typename A>
__global__ void __launch_bounds__(kWarps * 32, Occ<R>::kMinBlocks) pairhmm_kernel(HmmArgs g) {
    __shared__ __align__(128) uint8_t s_hap[kWarps][kRing];
    __shared__ __align__(8) uint64_t s_bar[kWarps][2];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *ring = s_hap[warp];
    uint64_t *bar = s_bar[warp];
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    uint32_t parity0 = 0, parity1 = 0;  // phase parity of the two ring slots

    const HmmConsts c = A::kLog ? g.ln : g.lin;
    const int n = *g.count;
    const int32_t *list = g.list ? g.list + (g.list_off ? *g.list_off : 0) : nullptr;

    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(g.cursor, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= n) break;
        const int pair = list ? list[p] : p;
        const int64_t h = g.pair_hap ? (int64_t)g.pair_hap[pair] : (int64_t)pair;
        const int64_t rd = g.pair_read ? (int64_t)g.pair_read[pair] : (int64_t)pair;
        const int64_t ho = g.hap_off[h];
        const int lx = (int)(g.hap_off[h + 1] - ho);
        const int64_t ro = g.read_off[rd];
        const int ly = (int)(g.read_off[rd + 1] - ro);
        int band = kMedMax;
        if (BANDED) {
            int k = g.med[pair];
            band = k < 0 ? kMedMax : k;
        }
        if (lx <= 0 || ly <= 0) {
            // len_x == 0: ln_sum_exp of no columns = ln 0; len_y == 0 cannot occur in the reference
            if (lane == 0) g.out[pair] = -INFINITY;
            continue;
        }

        // ---- stage the first haplotype chunk (TMA) while the row constants are loaded.
        // Bulk copies need 16-byte aligned sources: copy from the aligned-down address and shift
        // the ring index by the misalignment.
        const int shift = (int)((uintptr_t)(g.hap + ho) & 15);
        const uint8_t *hsrc = g.hap + ho - shift;
        const int lxs = lx + shift;
        const int n_chunks = (lxs + kChunk - 1) / kChunk;
        __syncwarp();  // every lane is done with the ring contents of the previous pair
        if (lane == 0) {
            uint32_t bytes = (uint32_t)min(kChunk, (lxs + 15) & ~15);
            mbar_expect_tx(&bar[0], bytes);
            tma_load_1d(ring, hsrc, bytes, &bar[0]);
        }

        // ---- per-row constants (ReadEmission, pairhmm.rs:160-200)
        double pm[R], pmm[R], cyo[R], cye[R];
        int rb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = lane * R + r;
            if (j < ly) {
                const int q = g.rqual[ro + j];
                rb[r] = g.rbase[ro + j];
                if (A::kLog) {
                    const double ln_eps = (double)q * -0.23025850929940456;
                    pm[r] = g.qlut[q * kQlutStride + 3];
                    pmm[r] = ln_eps + -1.098712293668443;  // ln(0.3333)
                    cyo[r] = ln_eps + c.t_gx;
                    cye[r] = ln_eps + c.t_gxe;
                } else {
                    const double eps = g.qlut[q * kQlutStride + 0];
                    pm[r] = g.qlut[q * kQlutStride + 1];
                    pmm[r] = g.qlut[q * kQlutStride + 2];
                    cyo[r] = eps * c.t_gx;
                    cye[r] = eps * c.t_gxe;
                }
            } else {  // padding row below the read: stays at zero mass
                rb[r] = -1;
                pm[r] = pmm[r] = cyo[r] = cye[r] = A::zero();
            }
        }
        const int last_lane = (ly - 1) / R;
        const int last_r = (ly - 1) - last_lane * R;

        // ---- state: previous column of my rows, values in flight between lanes
        double M[R], X[R], Y[R];
        int med[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            M[r] = X[r] = Y[r] = A::zero();
            med[r] = kMedMax;
        }
        double bM = A::zero(), bX = A::zero(), bY = A::zero();  // my bottom row, last column done
        double dM = lane == 0 ? A::one() : A::zero();           // diagonal inputs of my row 0
        double dX = A::zero(), dY = A::zero();
        int bMed = kMedMax, dMed = lane == 0 ? 0 : kMedMax;
        double acc = A::zero();

        mbar_wait(&bar[0], parity0);
        parity0 ^= 1;

        const int T = lx + last_lane;
        for (int t = 0; t < T; ++t) {
            // ---- haplotype ring maintenance (warp-uniform)
            const int ts = t + shift;
            if ((ts & (kChunk - 1)) == 32) {
                const int cnext = (ts >> 8) + 1;  // kChunk == 256
                if (cnext < n_chunks && lane == 0) {
                    const int off = cnext * kChunk;
                    uint32_t bytes = (uint32_t)min(kChunk, ((lxs - off) + 15) & ~15);
                    uint64_t *b = &bar[cnext & 1];
                    mbar_expect_tx(b, bytes);
                    tma_load_1d(ring + (cnext & 1) * kChunk, hsrc + off, bytes, b);
                }
            } else if ((ts & (kChunk - 1)) == 0 && t > 0) {
                const int cidx = ts >> 8;
                if (cidx < n_chunks) {
                    if (cidx & 1) { mbar_wait(&bar[1], parity1); parity1 ^= 1; }
                    else { mbar_wait(&bar[0], parity0); parity0 ^= 1; }
                }
            }
            const int i = t - lane;
            int xb = ring[(i + shift) & (kRing - 1)];
            if ((unsigned)(xb - 'a') < 26u) xb -= 32;  // to_ascii_uppercase (pairhmm.rs:190)

            // ---- receive the bottom row of the lane above (its column i)
            double uM = __shfl_up_sync(0xffffffffu, bM, 1);
            double uX = __shfl_up_sync(0xffffffffu, bX, 1);
            double uY = __shfl_up_sync(0xffffffffu, bY, 1);
            int uMed = kMedMax;
            if (BANDED) uMed = __shfl_up_sync(0xffffffffu, bMed, 1);
            if (lane == 0) {  // row 0 boundary: fm[curr][0] = 0, fx = fy = 0, med[curr][0] = MAX
                uM = A::zero(); uX = A::zero(); uY = A::zero();
                uMed = kMedMax;
            }

            double diagM = dM, diagX = dX, diagY = dY, upM = uM, upY = uY;
            int diagMed = dMed, upMed = uMed;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool match = xb == rb[r];
                const double e = match ? pm[r] : pmm[r];
                double s = A::template sum3<APPROX>(A::mul(c.t_mm, diagM), A::mul(c.t_xm, diagX),
                                                    A::mul(c.t_ym, diagY));
                double nM = A::mul(e, s);
                double nX = A::dot2(c.t_gy, M[r], c.t_gye, X[r]);
                double nY = A::dot2(cyo[r], upM, cye[r], upY);
                int nMed = kMedMax;
                if (BANDED) {
                    // PairHMM band: skip the cell if all three predecessors exceed max_edit_dist
                    const int leftMed = med[r];
                    const int mn = min(diagMed, min(upMed, leftMed));
                    const bool skip = mn > band;
                    nMed = min(min(diagMed + (match ? 0 : 1), leftMed + 1), upMed + 1);
                    nMed = min(nMed, kMedMax);
                    if (skip) {
                        nM = A::zero(); nX = A::zero(); nY = A::zero();
                        nMed = kMedMax;
                    }
                    diagMed = leftMed;
                    med[r] = nMed;
                    upMed = nMed;
                }
                diagM = M[r]; diagX = X[r]; diagY = Y[r];
                M[r] = nM; X[r] = nX; Y[r] = nY;
                upM = nM; upY = nY;
            }
            // next column's diagonal inputs of my row 0 are this column's "up" values
            dM = lane == 0 ? A::one() : uM;
            dX = uX;
            dY = uY;
            bM = M[R - 1]; bX = X[R - 1]; bY = Y[R - 1];
            if (BANDED) { dMed = lane == 0 ? 0 : uMed; bMed = med[R - 1]; }

            // ---- free end: add the last read row of column i (prob_cols)
            if (lane == last_lane && (unsigned)i < (unsigned)lx) {
                double tsum = A::zero();
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r == last_r) tsum = A::add(A::add(M[r], X[r]), Y[r]);
                acc = A::add(acc, tsum);
            }
        }

        // ---- result lives in lane last_lane
        acc = __shfl_sync(0xffffffffu, acc, last_lane);
        if (lane == 0) {
            double res;
            if (A::kLog) {
                res = acc;
            } else {
                // representable? (2^-940 of scaled mass; also catches inf/nan)
                if (!(acc >= 4.3e-283 && acc < 1.0e307)) {
                    const int slot = atomicAdd(g.fb_count, 1);
                    g.fb_list[slot] = pair;
                    res = NAN;
                } else {
                    res = log(acc) - c_scale_ln;
                }
            }
            if (res > 0.0) res = 0.0;  // "sum of paths can exceed 1" cap
            g.out[pair] = res;
        }
    }
}

// ---------------------------------------------------------------- classification of pairs by row count
__host__ __device__ inline int class_of_len(int ly) {
    if (ly <= 32) return 0;
    if (ly <= 64) return 1;
    if (ly <= 96) return 2;
    if (ly <= 128) return 3;
    if (ly <= 160) return 4;
    if (ly <= 192) return 5;
    if (ly <= 256) return 6;
    return 7;
}
static const int kClassR[kNumClasses] = {1, 2, 3, 4, 5, 6, 8, 0};

// counters layout (int32): [0..7] class counts, [8..15] class offsets, [16..23] scatter cursors,
// [24..31] work cursors (one per class launch), [32] fallback count, [33] fallback work cursor
constexpr int kCntCount = 0, kCntOff = 8, kCntScatter = 16, kCntWork = 24, kCntFb = 32,
              kCntFbWork = 33, kCntTotal = 64;

__global__ void classify_kernel(int64_t n_pairs, const int64_t *read_off, const int32_t *pair_read,
                                int32_t *counters) {
    __shared__ int s_cnt[kNumClasses];
    if (threadIdx.x < kNumClasses) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs;
         p += (int64_t)gridDim.x * blockDim.x) {
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
                               int32_t *counters, int32_t *list) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs;
         p += (int64_t)gridDim.x * blockDim.x) {
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

template <int R, bool BANDED, bool APPROX, typename A>
static cudaError_t launch_one(const HmmArgs &a, int grid, cudaStream_t st) {
    pairhmm_kernel<R, BANDED, APPROX, A><<<grid, kWarps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

template <bool BANDED, bool APPROX, typename A>
static cudaError_t launch_class(int cls, const HmmArgs &a, int sm_count, cudaStream_t st) {
    switch (cls) {
    case 0: return launch_one<1, BANDED, APPROX, A>(a, sm_count * Occ<1>::kMinBlocks, st);
    case 1: return launch_one<2, BANDED, APPROX, A>(a, sm_count * Occ<2>::kMinBlocks, st);
    case 2: return launch_one<3, BANDED, APPROX, A>(a, sm_count * Occ<3>::kMinBlocks, st);
    case 3: return launch_one<4, BANDED, APPROX, A>(a, sm_count * Occ<4>::kMinBlocks, st);
    case 4: return launch_one<5, BANDED, APPROX, A>(a, sm_count * Occ<5>::kMinBlocks, st);
    case 5: return launch_one<6, BANDED, APPROX, A>(a, sm_count * Occ<6>::kMinBlocks, st);
    case 6: return launch_one<8, BANDED, APPROX, A>(a, sm_count * Occ<8>::kMinBlocks, st);
    default: return cudaErrorInvalidValue;
    }
}

template <typename A>
static cudaError_t launch_dispatch(bool banded, bool approx, int cls, const HmmArgs &a, int sm_count,
                                   cudaStream_t st) {
    if (banded) {
        return approx ? launch_class<true, true, A>(cls, a, sm_count, st)
                      : launch_class<true, false, A>(cls, a, sm_count, st);
    }
    return approx ? launch_class<false, true, A>(cls, a, sm_count, st)
                  : launch_class<false, false, A>(cls, a, sm_count, st);
}

// PairHMM::new (Appendix A.2): transition constants from the four gap parameters
static void make_consts(const vlr_gap_params *gp, uint32_t flags, HmmConsts *lin, HmmConsts *ln) {
    auto l1me = [](double p) -> double { return p < -0.693 ? log1p(-exp(p)) : log(-expm1(p)); };
    auto lae = [](double a, double b) -> double {
        double p0 = a > b ? a : b, p1 = a > b ? b : a;
        if (p0 == -INFINITY) return -INFINITY;
        if (p1 == -INFINITY) return p0;
        return p0 + log1p(exp(p1 - p0));
    };
    const double gx = gp->prob_gap_x, gy = gp->prob_gap_y, gxe = gp->prob_gap_x_extend,
                 gye = gp->prob_gap_y_extend;
    const double no_gap = l1me(lae(gx, gy));
    const double no_gxe = l1me(gxe), no_gye = l1me(gye);
    ln->t_mm = no_gap;
    if (flags & VLR_HMM_PATH_CLOSE) { ln->t_xm = no_gye; ln->t_ym = no_gxe; }
    else { ln->t_xm = no_gxe; ln->t_ym = no_gye; }
    ln->t_gy = gy; ln->t_gye = gye; ln->t_gx = gx; ln->t_gxe = gxe;
    lin->t_mm = exp(ln->t_mm); lin->t_xm = exp(ln->t_xm); lin->t_ym = exp(ln->t_ym);
    lin->t_gy = exp(gy); lin->t_gye = exp(gye); lin->t_gx = exp(gx); lin->t_gxe = exp(gxe);
}

struct PairWs {
    int32_t *counters;  // kCntTotal
    int32_t *list;      // n_pairs
    int32_t *fb_list;   // n_pairs
};
static size_t pair_ws_bytes(int64_t n_pairs) {
    return align_up(kCntTotal * sizeof(int32_t), 256) + 2 * align_up((size_t)n_pairs * 4 + 4, 256);
}
static PairWs carve_ws(void *ws, int64_t n_pairs) {
    PairWs w;
    uint8_t *p = (uint8_t *)ws;
    w.counters = (int32_t *)p;
    p += align_up(kCntTotal * sizeof(int32_t), 256);
    w.list = (int32_t *)p;
    p += align_up((size_t)n_pairs * 4 + 4, 256);
    w.fb_list = (int32_t *)p;
    return w;
}

}  // namespace vlr

using namespace vlr;

extern "C" {

size_t vlr_pairhmm_workspace_bytes(int64_t n_pairs, int64_t max_len_x, int64_t max_len_y) {
    (void)max_len_x; (void)max_len_y;
    return pair_ws_bytes(n_pairs < 1 ? 1 : n_pairs);
}

int vlr_pairhmm_batch_dev(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *d_hap_bytes,
                          const int64_t *d_hap_off, int64_t n_haps, const uint8_t *d_read_bases,
                          const uint8_t *d_read_quals, const int64_t *d_read_off, int64_t n_reads,
                          const int32_t *d_pair_hap, const int32_t *d_pair_read,
                          const int32_t *d_max_edit_dist, const vlr_gap_params *gap, uint32_t flags,
                          void *d_workspace, size_t workspace_bytes, double *d_out_lnprob) {
    if (!ctx) return VLR_ERR_INVALID;
    if (n_pairs < 0 || n_pairs > 0x7fffffff)
        return set_err(ctx, VLR_ERR_INVALID, "n_pairs out of range");
    if (n_pairs == 0) return VLR_OK;
    if (!d_hap_bytes || !d_hap_off || !d_read_bases || !d_read_quals || !d_read_off || !gap ||
        !d_workspace || !d_out_lnprob)
        return set_err(ctx, VLR_ERR_INVALID, "null pointer argument");
    if ((!d_pair_hap && n_haps < n_pairs) || (!d_pair_read && n_reads < n_pairs))
        return set_err(ctx, VLR_ERR_INVALID, "identity pair mapping needs n_haps/n_reads >= n_pairs");
    if (workspace_bytes < pair_ws_bytes(n_pairs))
        return set_err(ctx, VLR_ERR_INVALID, "workspace too small");
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PairWs w = carve_ws(d_workspace, n_pairs);
    VLR_CUDA(ctx, cudaMemsetAsync(w.counters, 0, kCntTotal * sizeof(int32_t), st));
    const int64_t want = (n_pairs + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    const int cgrid = (int)(want < cap ? want : cap);
    classify_kernel<<<cgrid, 256, 0, st>>>(n_pairs, d_read_off, d_pair_read, w.counters);
    VLR_LAUNCH_CHECK(ctx);
    class_scan_kernel<<<1, 32, 0, st>>>(w.counters);
    VLR_LAUNCH_CHECK(ctx);
    scatter_kernel<<<cgrid, 256, 0, st>>>(n_pairs, d_read_off, d_pair_read, w.counters, w.list);
    VLR_LAUNCH_CHECK(ctx);

    HmmArgs a;
    a.hap = d_hap_bytes; a.hap_off = d_hap_off;
    a.rbase = d_read_bases; a.rqual = d_read_quals; a.read_off = d_read_off;
    a.pair_hap = d_pair_hap; a.pair_read = d_pair_read; a.med = d_max_edit_dist;
    a.fb_list = w.fb_list; a.fb_count = w.counters + kCntFb;
    a.out = d_out_lnprob; a.qlut = ctx->d_qlut;
    make_consts(gap, flags, &a.lin, &a.ln);
    const bool banded = d_max_edit_dist != nullptr;
    const bool approx = (flags & VLR_HMM_SUM3_APPROX) != 0;

    // Class sizes live on the device: every class kernel is launched as a persistent grid with a
    // dynamic work counter and exits at once when its list is empty, so the device entry point
    // never synchronises with the host.
    for (int cls = 0; cls < 7; ++cls) {
        HmmArgs ac = a;
        ac.list = w.list;
        ac.list_off = w.counters + kCntOff + cls;
        ac.count = w.counters + kCntCount + cls;
        ac.cursor = w.counters + kCntWork + cls;
        cudaError_t e = launch_dispatch<Lin>(banded, approx, cls, ac, ctx->sm_count, st);
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
        cudaError_t e = launch_dispatch<Log>(banded, approx, 6, ac, ctx->sm_count, st);
        if (e != cudaSuccess)
            return set_err(ctx, VLR_ERR_CUDA, "pairhmm launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    mark_unsupported_kernel<<<32, 256, 0, st>>>(w.list, w.counters, d_out_lnprob);
    VLR_LAUNCH_CHECK(ctx);
    return VLR_OK;
}

int vlr_pairhmm_batch(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes,
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
    // validate indices and lengths on the host (cheap, and errors must not reach the kernels)
    int64_t max_ly = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        int64_t l = read_off[r + 1] - read_off[r];
        if (l < 0) return set_err(ctx, VLR_ERR_INVALID, "read_off not monotone");
        if (l > max_ly) max_ly = l;
    }
    if (max_ly > 256)
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "read window of %lld rows exceeds 256",
                       (long long)max_ly);
    if (pair_hap || pair_read)
        for (int64_t k = 0; k < n_pairs; ++k) {
            if (pair_hap && (pair_hap[k] < 0 || pair_hap[k] >= n_haps))
                return set_err(ctx, VLR_ERR_INVALID, "pair_hap[%lld] out of range", (long long)k);
            if (pair_read && (pair_read[k] < 0 || pair_read[k] >= n_reads))
                return set_err(ctx, VLR_ERR_INVALID, "pair_read[%lld] out of range", (long long)k);
        }
    const int64_t hap_lo = hap_off[0], hap_hi = hap_off[n_haps];
    const int64_t rd_lo = read_off[0], rd_hi = read_off[n_reads];
    for (int64_t h = 0; h < n_haps; ++h)
        if (hap_off[h + 1] < hap_off[h]) return set_err(ctx, VLR_ERR_INVALID, "hap_off not monotone");
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
    const size_t ws_bytes = pair_ws_bytes(n_pairs);
    void *d_ws = dev_scratch(ctx, S_WS, ws_bytes);
    double *d_out = (double *)dev_scratch(ctx, S_OUT, (size_t)n_pairs * 8);
    if (!d_hap || !d_hoff || !d_rb || !d_rq || !d_roff || !d_ws || !d_out || (pair_hap && !d_ph) ||
        (pair_read && !d_pr) || (max_edit_dist && !d_med))
        return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    // H2D straight from the caller's buffers (full speed when they are pinned)
    VLR_CUDA(ctx, cudaMemcpyAsync(d_hap, hap_bytes + hap_lo, hap_bytes_n, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_hoff, hap_off, (size_t)(n_haps + 1) * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_rb, read_bases + rd_lo, rd_bytes_n, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_rq, read_quals + rd_lo, rd_bytes_n, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_roff, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (pair_hap) VLR_CUDA(ctx, cudaMemcpyAsync(d_ph, pair_hap, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
    if (pair_read) VLR_CUDA(ctx, cudaMemcpyAsync(d_pr, pair_read, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
    if (max_edit_dist)
        VLR_CUDA(ctx, cudaMemcpyAsync(d_med, max_edit_dist, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
    // offsets are absolute into the caller's arrays: rebase the device pointers instead of the offsets
    int rc = vlr_pairhmm_batch_dev(ctx, n_pairs, d_hap - hap_lo, d_hoff, n_haps, d_rb - rd_lo,
                                   d_rq - rd_lo, d_roff, n_reads, d_ph, d_pr, d_med, gap, flags, d_ws,
                                   ws_bytes, d_out);
    if (rc != VLR_OK) return rc;
    VLR_CUDA(ctx, cudaMemcpyAsync(out_lnprob, d_out, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    return VLR_OK;
}

}  // extern "C"
